// rb2_tip_math.cuh -- hyperboloid-tip pair arithmetic shared by the pair / field kernels (rb2_pair.cu) and the
// device-resident tip sampler (rb2_mh.cu).
#pragma once

#include "rb2_internal.cuh"

// ---- hyperboloid tip math ---------------------------------------------------------------------
// Two forms of every pair term.  The literal one (IEEE sqrt / divide, operation by operation like the .inc files) is the
// slow path and the per-particle set-up; the fast one rewrites the same sums on the MUFU-seeded inverse cubes of
// rb2_internal.cuh (~36 FP64 instructions per pair instead of ~155):
//   Coulomb + Sphere_IC_field of source b on target a (a imaged):  f = d (w_c + pre w_a) - pre k_a w_b (im_a - r_b)
//   with d = r_a - r_b, w_c = 1/(|d| + eps)^3, w_a = 1/|d|^3, w_b = 1/|r_b - im_a|^3, k_a = r_tip / dis_a, pre = q_0/(4 pi eps_0).
// Pairs closer than 1e-11 m raise the `close` flag (rb2_is_close on |d|^2) and are redone by the caller on the slow path.
struct TipImage {
    double dis_a, x_im, y_im, z_im;
};
// src/acc_tip_image_point.inc:14-18
__device__ __forceinline__ TipImage tip_image_point(const TipParams &T, double x_a, double y_a, double z_a)
{
    TipImage im;
    const double zr = z_a - T.z_0;
    const double zz = zr * zr;
    im.dis_a = sqrt(x_a * x_a + y_a * y_a + zz);
    im.z_im = T.z_0 + (T.r_tip * T.r_tip) / (sqrt(1.0 + (x_a * x_a) / zz + (y_a * y_a) / zz) * im.dis_a);
    im.x_im = (im.z_im - T.z_0) * x_a / zr;
    im.y_im = (im.z_im - T.z_0) * y_a / zr;
    return im;
}
// src/acc_tip_ic_force.inc:17-22 -- carries q_0/(4 pi eps0) itself, like Sphere_IC_field
__device__ __forceinline__ void tip_ic_force(const TipParams &T, const TipImage &im, double x_a, double y_a, double z_a,
                                             double x_b, double y_b, double z_b, double &ic_x, double &ic_y, double &ic_z)
{
    const double pre = 1.0 * rb2k::q_0 / (4.0 * RB2_PI * rb2k::epsilon_0);
    const double sa = (x_b - x_a) * (x_b - x_a) + (y_b - y_a) * (y_b - y_a) + (z_b - z_a) * (z_b - z_a);
    const double sb = (x_b - im.x_im) * (x_b - im.x_im) + (y_b - im.y_im) * (y_b - im.y_im) + (z_b - im.z_im) * (z_b - im.z_im);
    const double tmp_dis_a = sa * sqrt(sa);  // (..)**(3/2)
    const double tmp_dis_b = sb * sqrt(sb);
    // two IEEE divides instead of the six of the source line by line (same value to rounding: 1e-16, the parity bar is 1e-11)
    const double wa = 1.0 / tmp_dis_a, wb = T.r_tip / (im.dis_a * tmp_dis_b);
    ic_x = pre * ((x_a - x_b) * wa - (im.x_im - x_b) * wb);
    ic_y = pre * ((y_a - y_b) * wa - (im.y_im - y_b) * wb);
    ic_z = pre * ((z_a - z_b) * wa - (im.z_im - z_b) * wb);
}


// Field at the point (xi, yi, zi) -- whose sphere image is im_i -- of ONE source pj, without the source's charge:
// Coulomb (src/mod_verlet.F90:1511-1518) + Sphere_IC_field(p, r_j) (:1520: the field point is imaged).
__device__ __forceinline__ void tip_point_field(const TipParams &T, const TipImage &im_i, bool do_ic, double xi, double yi, double zi,
                                                const double4 pj, double &fx, double &fy, double &fz)
{
    const double dx = xi - pj.x, dy = yi - pj.y, dz = zi - pj.z;
    const double r = sqrt(dx * dx + dy * dy + dz * dz) + rb2k::soft;
    const double inv_r3 = 1.0 / (r * r * r);
    fx = inv_r3 * dx; fy = inv_r3 * dy; fz = inv_r3 * dz;
    if (do_ic) {
        double ic_x, ic_y, ic_z;
        tip_ic_force(T, im_i, xi, yi, zi, pj.x, pj.y, pj.z, ic_x, ic_y, ic_z);
        fx += ic_x; fy += ic_y; fz += ic_z;
    }
}

// Per-particle constants of the fast forms: the sphere image and kk = pre * r_tip / dis_a (exact arithmetic, once per
// particle / field point).
__device__ __forceinline__ double4 tip_image_packed(const TipParams &T, double x, double y, double z)
{
    const TipImage im = tip_image_point(T, x, y, z);
    const double pre = 1.0 * rb2k::q_0 / (4.0 * RB2_PI * rb2k::epsilon_0);
    return make_double4(im.x_im, im.y_im, im.z_im, pre * T.r_tip / im.dis_a);
}
// Target a = (xi, yi, zi) with image ia, source b = pj: the field-point form and the acceleration of a particle from a
// HIGHER-indexed source (src/mod_verlet.F90:1511-1520, :1380-1400 with a = i).
__device__ __forceinline__ void tip_pair_fast_upper(const double4 ia, bool do_ic, double xi, double yi, double zi, const double4 pj,
                                                    double &fx, double &fy, double &fz, bool &close)
{
    const double dx = xi - pj.x, dy = yi - pj.y, dz = zi - pj.z;
    const double s = fma(dz, dz, fma(dy, dy, fma(dx, dx, RB2_S_FLOOR)));
    close = close || rb2_is_close(s);
    double wc, wa;
    rb2_inv_r3_both(s, wc, wa);
    if (!do_ic) { fx = wc * dx; fy = wc * dy; fz = wc * dz; return; }
    const double pre = 1.0 * rb2k::q_0 / (4.0 * RB2_PI * rb2k::epsilon_0);
    const double ex = ia.x - pj.x, ey = ia.y - pj.y, ez = ia.z - pj.z;
    const double V = ia.w * rb2_inv_r3_far(fma(ez, ez, fma(ey, ey, ex * ex)));
    const double U = fma(pre, wa, wc);
    fx = fma(dx, U, -(V * ex));
    fy = fma(dy, U, -(V * ey));
    fz = fma(dz, U, -(V * ez));
}
// Acceleration of particle i = (xi, yi, zi) from a LOWER-indexed source j = pj with image ij: the reference evaluates the
// pair with a = j, b = i and mirrors x, y (sgn_xy = -1, src/mod_verlet.F90:1380-1400):
//   f_xy = d (w_c + pre w_a) + pre k_j w_b (im_j - r_i),   f_z = d_z (w_c - pre w_a) - pre k_j w_b (im_j - r_i)_z
__device__ __forceinline__ void tip_pair_fast_lower(const double4 ij, bool do_ic, double xi, double yi, double zi, const double4 pj,
                                                    double &fx, double &fy, double &fz, bool &close)
{
    const double dx = xi - pj.x, dy = yi - pj.y, dz = zi - pj.z;
    const double s = fma(dz, dz, fma(dy, dy, fma(dx, dx, RB2_S_FLOOR)));
    close = close || rb2_is_close(s);
    double wc, wa;
    rb2_inv_r3_both(s, wc, wa);
    if (!do_ic) { fx = wc * dx; fy = wc * dy; fz = wc * dz; return; }
    const double pre = 1.0 * rb2k::q_0 / (4.0 * RB2_PI * rb2k::epsilon_0);
    const double ex = ij.x - xi, ey = ij.y - yi, ez = ij.z - zi;
    const double V = ij.w * rb2_inv_r3_far(fma(ez, ez, fma(ey, ey, ex * ex)));
    const double U1 = fma(pre, wa, wc), U2 = fma(-pre, wa, wc);
    fx = fma(dx, U1, V * ex);
    fy = fma(dy, U1, V * ey);
    fz = fma(dz, U2, -(V * ez));
}

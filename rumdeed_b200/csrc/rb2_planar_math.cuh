// rb2_planar_math.cuh -- planar pair arithmetic shared by the gather kernel (rb2_pair.cu) and the
// pair-symmetric kernel (rb2_pair_sym.cu).
//
// Coulomb (reference src/mod_verlet.F90:1297-1306) + image series
// (src/acc_ic_planar_series.inc:20-63) of a source at (xj, yj, zj) evaluated at (xi, yi, zi),
// WITHOUT any charge prefactor.  Partner heights relative to the evaluation height, with
// S = z_i + z_j and D = z_i - z_j:  opposite charge  S, S-2nd, S+2nd ;  same charge  D-2nd, D+2nd.
// With the roles swapped (reference j < i) S is unchanged and D -> -D, which maps the two
// same-charge partners onto each other with dz negated: same weights, opposite z-sum.  Hence
//   field_x = dx * U,  field_y = dy * U,  field_z = dz * wc - Zopp + sgn * Zsame
// with U = wc + W (Coulomb + signed lateral image weights) and sgn = +1 when the evaluation
// particle has the lower index, -1 otherwise (F8-ii).
#pragma once

#include "rb2_internal.cuh"

struct Acc4 {
    double x, y, z, t;  // t: same-charge partner z-sum, signed by the caller (image roles)
};

struct PairW {
    double dx, dy, dz;
    double U;      // wc + W
    double wc;     // softened 1/r^3 of the direct pair
    double Zopp;   // opposite-charge partner z-sum (enters with a minus)
    double Zsame;  // same-charge partner z-sum (enters with the role sign)
};

// NIC: -1 image charge off, 0, 1, >= 2 (runtime N_ic_max loop).  `close` collects "lateral offset below 1e-11 m" over
// the calls (rb2_is_close).
// The symmetric rounds (i < j throughout, no role sign) only need the image z-sum as a whole: icz = Zsame - Zopp in one
// FMA chain; FAR: the three partners that are at least d away skip the softening term (PlanarParams.far_ok).  N_ic_max = 1.
struct PairS {
    double dx, dy, dz, U, wc, icz;
};
template <bool FAR>
__device__ __forceinline__ PairS planar_weights_sym1(double xi, double yi, double zi, double xj, double yj, double zj,
                                                     const PlanarParams &P, bool &close)
{
    PairS w;
    w.dx = xi - xj;
    w.dy = yi - yj;
    w.dz = zi - zj;
    const double dxy2 = fma(w.dy, w.dy, fma(w.dx, w.dx, RB2_S_FLOOR));
    close = close || rb2_is_close(dxy2);
    w.wc = rb2_inv_r3_soft(fma(w.dz, w.dz, dxy2));
    const double S = zi + zj;
    const double a1 = S - P.two_d, a2 = S + P.two_d, b1 = w.dz - P.two_d, b2 = w.dz + P.two_d;
    const double w0 = rb2_inv_r3_soft(fma(S, S, dxy2));
    const double w1 = rb2_inv_r3_soft(fma(a1, a1, dxy2));
    const double w2 = FAR ? rb2_inv_r3_far(fma(a2, a2, dxy2)) : rb2_inv_r3_soft(fma(a2, a2, dxy2));
    const double w3 = FAR ? rb2_inv_r3_far(fma(b1, b1, dxy2)) : rb2_inv_r3_soft(fma(b1, b1, dxy2));
    const double w4 = FAR ? rb2_inv_r3_far(fma(b2, b2, dxy2)) : rb2_inv_r3_soft(fma(b2, b2, dxy2));
    w.U = w.wc + ((w3 + w4) - ((w0 + w1) + w2));
    w.icz = fma(b2, w4, fma(b1, w3, fma(-a2, w2, fma(-a1, w1, -(S * w0)))));
    return w;
}

template <int NIC, bool FAR = false>
__device__ __forceinline__ PairW planar_weights(double xi, double yi, double zi, double xj, double yj, double zj,
                                                const PlanarParams &P, bool &close)
{
    PairW w;
    w.dx = xi - xj;
    w.dy = yi - yj;
    w.dz = zi - zj;
    const double dxy2 = fma(w.dy, w.dy, fma(w.dx, w.dx, RB2_S_FLOOR));
    close = close || rb2_is_close(dxy2);
    w.wc = rb2_inv_r3_soft(fma(w.dz, w.dz, dxy2));
    if (NIC < 0) {
        w.U = w.wc;
        w.Zopp = 0.0;
        w.Zsame = 0.0;
        return w;
    }
    const double S = zi + zj;
    const double w0 = rb2_inv_r3_soft(fma(S, S, dxy2));
    double W = -w0;
    w.Zopp = S * w0;
    w.Zsame = 0.0;
    if (NIC == 1) {
        const double a1 = S - P.two_d, a2 = S + P.two_d, b1 = w.dz - P.two_d, b2 = w.dz + P.two_d;
        const double w1 = rb2_inv_r3_soft(fma(a1, a1, dxy2));
        const double w2 = FAR ? rb2_inv_r3_far(fma(a2, a2, dxy2)) : rb2_inv_r3_soft(fma(a2, a2, dxy2));
        const double w3 = FAR ? rb2_inv_r3_far(fma(b1, b1, dxy2)) : rb2_inv_r3_soft(fma(b1, b1, dxy2));
        const double w4 = FAR ? rb2_inv_r3_far(fma(b2, b2, dxy2)) : rb2_inv_r3_soft(fma(b2, b2, dxy2));
        W = (w3 + w4) - ((w0 + w1) + w2);
        w.Zopp = fma(a2, w2, fma(a1, w1, w.Zopp));
        w.Zsame = fma(b2, w4, b1 * w3);
    } else if (NIC >= 2) {
        for (int n = 1; n <= P.nic; ++n) {
            const double h = P.two_d * (double)n;
            const double a1 = S - h, a2 = S + h, b1 = w.dz - h, b2 = w.dz + h;
            const double w1 = rb2_inv_r3_soft(fma(a1, a1, dxy2));
            const double w2 = rb2_inv_r3_soft(fma(a2, a2, dxy2));
            const double w3 = rb2_inv_r3_soft(fma(b1, b1, dxy2));
            const double w4 = rb2_inv_r3_soft(fma(b2, b2, dxy2));
            W += (w3 + w4) - (w1 + w2);
            w.Zopp = fma(a2, w2, fma(a1, w1, w.Zopp));
            w.Zsame = fma(b2, w4, fma(b1, w3, w.Zsame));
        }
    }
    w.U = w.wc + W;
    return w;
}

// The same quantities for a laterally close pair (slow path), with the reference's own operations: every partner
// height z_ic and offset diff_z = z_a - z_ic formed like src/acc_ic_planar_series.inc:20-63 does -- the shared-S / shared-D
// shortcuts above round differently by ulp(2d) ~ 4e-22 m, which matters once a partner is within ~1e-10 m -- and the
// inverse cubes by IEEE sqrt / divide (src/mod_verlet.F90:1302-1303).  `swapped`: the reference's roles for j < i
// (z_a = z_j, z_b = z_i, src/mod_verlet.F90:1316-1323); the result is returned in the convention of planar_weights
// (the caller's role sign on Zsame then reproduces the reference's value).
template <int NIC>
__device__ __forceinline__ PairW planar_weights_exact(double xi, double yi, double zi, double xj, double yj, double zj,
                                                      const PlanarParams &P, bool swapped)
{
    PairW w;
    w.dx = xi - xj;
    w.dy = yi - yj;
    w.dz = zi - zj;
    const double dxy2 = fma(w.dy, w.dy, fma(w.dx, w.dx, RB2_S_FLOOR));
    w.wc = rb2_inv_r3_exact(__dadd_rn(dxy2, __dmul_rn(w.dz, w.dz)));
    w.U = w.wc;
    w.Zopp = 0.0;
    w.Zsame = 0.0;
    if (NIC < 0) return w;
    const double z_a = swapped ? zj : zi, z_b = swapped ? zi : zj;
    double W, zo, zs = 0.0;
    {
        const double z_ic = -1.0 * z_b, dz = z_a - z_ic;
        const double v = rb2_inv_r3_exact(__dadd_rn(dxy2, __dmul_rn(dz, dz)));
        W = -v;
        zo = __dmul_rn(dz, v);
    }
    const int nmax = (NIC == 0) ? 0 : (NIC == 1 ? 1 : P.nic);
    for (int n = 1; n <= nmax; ++n) {
        const double h = P.two_d * (double)n;  // 2.0d0*n*d_loc
        double z_ic, dz, v;
        z_ic = h - z_b; dz = z_a - z_ic;
        v = rb2_inv_r3_exact(__dadd_rn(dxy2, __dmul_rn(dz, dz))); W -= v; zo = __dadd_rn(zo, __dmul_rn(dz, v));
        z_ic = -h - z_b; dz = z_a - z_ic;
        v = rb2_inv_r3_exact(__dadd_rn(dxy2, __dmul_rn(dz, dz))); W -= v; zo = __dadd_rn(zo, __dmul_rn(dz, v));
        z_ic = h + z_b; dz = z_a - z_ic;
        v = rb2_inv_r3_exact(__dadd_rn(dxy2, __dmul_rn(dz, dz))); W += v; zs = __dadd_rn(zs, __dmul_rn(dz, v));
        z_ic = -h + z_b; dz = z_a - z_ic;
        v = rb2_inv_r3_exact(__dadd_rn(dxy2, __dmul_rn(dz, dz))); W += v; zs = __dadd_rn(zs, __dmul_rn(dz, v));
    }
    w.U = w.wc + W;
    w.Zopp = zo;
    w.Zsame = swapped ? -zs : zs;
    return w;
}

// Gather form: accumulate q_j * field(i <- j) into a (a.t collects the same-charge z-sum with
// the charge qs, which the caller signs: per tile, or per element through qs itself).
template <int NIC>
__device__ __forceinline__ void planar_apply(const PairW &w, double qj, double qs, Acc4 &a)
{
    const double t = qj * w.U;
    a.x = fma(w.dx, t, a.x);
    a.y = fma(w.dy, t, a.y);
    if (NIC < 0) {
        a.z = fma(w.dz, t, a.z);
        return;
    }
    a.z = fma(qj, fma(w.dz, w.wc, -w.Zopp), a.z);
    a.t = fma(qs, w.Zsame, a.t);
}
template <int NIC, bool FAR = false>
__device__ __forceinline__ void planar_term(double xi, double yi, double zi, const double4 pj, double qj, double qs,
                                            const PlanarParams &P, Acc4 &a, bool &close)
{
    planar_apply<NIC>(planar_weights<NIC, FAR>(xi, yi, zi, pj.x, pj.y, pj.z, P, close), qj, qs, a);
}
// slow path of a laterally close pair; swapped: the source has the lower index (qs then carries the minus sign)
template <int NIC>
__device__ __forceinline__ void planar_term_exact(double xi, double yi, double zi, const double4 pj, double qj, double qs,
                                                  const PlanarParams &P, Acc4 &a, bool swapped)
{
    planar_apply<NIC>(planar_weights_exact<NIC>(xi, yi, zi, pj.x, pj.y, pj.z, P, swapped), qj, qs, a);
}

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

// NIC: -1 image charge off, 0, 1, >= 2 (runtime N_ic_max loop).  EXACT: the reference's sqrt / divide for the inverse
// cubes (slow path of pairs flagged `close`); otherwise `close` collects "lateral offset below 1e-11 m" over the calls.
template <int NIC, bool EXACT = false>
__device__ __forceinline__ PairW planar_weights(double xi, double yi, double zi, double xj, double yj, double zj,
                                                const PlanarParams &P, bool &close)
{
    PairW w;
    w.dx = xi - xj;
    w.dy = yi - yj;
    w.dz = zi - zj;
    const double dxy2 = fma(w.dy, w.dy, fma(w.dx, w.dx, RB2_S_FLOOR));
    if (!EXACT) close = close || rb2_is_close(dxy2);
    w.wc = rb2_inv_r3_sel<EXACT>(fma(w.dz, w.dz, dxy2));
    if (NIC < 0) {
        w.U = w.wc;
        w.Zopp = 0.0;
        w.Zsame = 0.0;
        return w;
    }
    const double S = zi + zj;
    const double w0 = rb2_inv_r3_sel<EXACT>(fma(S, S, dxy2));
    double W = -w0;
    w.Zopp = S * w0;
    w.Zsame = 0.0;
    if (NIC == 1) {
        const double a1 = S - P.two_d, a2 = S + P.two_d, b1 = w.dz - P.two_d, b2 = w.dz + P.two_d;
        const double w1 = rb2_inv_r3_sel<EXACT>(fma(a1, a1, dxy2));
        const double w2 = rb2_inv_r3_sel<EXACT>(fma(a2, a2, dxy2));
        const double w3 = rb2_inv_r3_sel<EXACT>(fma(b1, b1, dxy2));
        const double w4 = rb2_inv_r3_sel<EXACT>(fma(b2, b2, dxy2));
        W = (w3 + w4) - ((w0 + w1) + w2);
        w.Zopp = fma(a2, w2, fma(a1, w1, w.Zopp));
        w.Zsame = fma(b2, w4, b1 * w3);
    } else if (NIC >= 2) {
        for (int n = 1; n <= P.nic; ++n) {
            const double h = P.two_d * (double)n;
            const double a1 = S - h, a2 = S + h, b1 = w.dz - h, b2 = w.dz + h;
            const double w1 = rb2_inv_r3_sel<EXACT>(fma(a1, a1, dxy2));
            const double w2 = rb2_inv_r3_sel<EXACT>(fma(a2, a2, dxy2));
            const double w3 = rb2_inv_r3_sel<EXACT>(fma(b1, b1, dxy2));
            const double w4 = rb2_inv_r3_sel<EXACT>(fma(b2, b2, dxy2));
            W += (w3 + w4) - (w1 + w2);
            w.Zopp = fma(a2, w2, fma(a1, w1, w.Zopp));
            w.Zsame = fma(b2, w4, fma(b1, w3, w.Zsame));
        }
    }
    w.U = w.wc + W;
    return w;
}

// Gather form: accumulate q_j * field(i <- j) into a (a.t collects the same-charge z-sum with
// the charge qs, which the caller signs: per tile, or per element through qs itself).
template <int NIC, bool EXACT = false>
__device__ __forceinline__ void planar_term(double xi, double yi, double zi, const double4 pj, double qj, double qs,
                                            const PlanarParams &P, Acc4 &a, bool &close)
{
    const PairW w = planar_weights<NIC, EXACT>(xi, yi, zi, pj.x, pj.y, pj.z, P, close);
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

// rb2_collisions.cu -- electron / N2 collisions of the Ion configuration, collision_mode 1 and 2:
//   Do_Electron_Atom_Collisions        src/mod_collisions.F90:30-76
//   Update_Collision_Data[_All_ots]    :2103-2169   (energy, cross sections, Kramers recombination radius)
//   Do_Continuous_Ionization_ots       :558-705     (O(nrElec), random)
//   Do_Discrete_Recombination_ots      :86-245      (O(nrIon x nrElec), quartic of src/mod_polynomialroots.F90)
//
// Recombination is the other quadratic loop of the program: the reference solves a quartic (entry time of the
// electron's parabola into the Kramers sphere of the ion) for EVERY (ion, electron) pair.  Here:
//   k_recomb_prepare  one pass over the particles: ions past their life time are marked (remove_top), live ions
//                     and live electrons go to compact lists; an electron's record is {x, y, z, b^2} with
//                     b = R + 2 (|v| dt + |a| dt^2 / 2): outside that radius its parabola cannot reach the sphere
//                     of radius R within the step (all real roots of the quartic then have |t| > 1.4 dt).
//   k_recomb_sweep    thread per ion, electron records tiled through shared memory, 7 FP64 instructions per pair
//                     (3 subtractions, 3 FMAs, 1 compare); survivors are appended to a candidate list.
//   k_recomb_solve    thread per candidate: the reference's test, operation by operation -- inside the radius
//                     already, else SolvePolynomial and the root selection by return code (:153-185).  This file
//                     is compiled with -fmad=false so that the solver's rounding follows the uncontracted source.
//   host              sorts the hits by (ion, electron) and replays the serial claim rule (an ion takes its first
//                     colliding electron that no lower-indexed ion has taken, :196-236), then marks both with
//                     remove_recom.  Hits are rare (Kramers radii are ~1e-12 m), so this list is tiny.
// Ionisation is one thread per particle with a counter-based generator (Philox4x32-10 keyed by seed, step and slot);
// the rare events are appended to a list, sorted by slot on the host and added in the reference's order (ejected
// electron, then ion).  The reference's RANDOM_NUMBER stream is compiler specific: parity is statistical there.
#include <algorithm>
#include <unordered_set>

#include "rb2_internal.cuh"

namespace {

constexpr int TPB = 128;
constexpr int ETILE = 256;

struct CollParams {
    const double *tot_e, *tot_d, *ion_e, *ion_d;
    int    n_tot, n_ion;
    double Ryd, Z_eff, N_n, N_bind;  // src/mod_global.F90:50-68
    double n_d, cyl_radius, dt;
    int    step, ion_life_time;
    unsigned long long seed;
    double inj_max;                  // folded_normal_max(5, 25), the envelope of Get_Injected_Vec
};

struct CollCounts {  // device counters of one call
    int n_elec, n_ion, n_cand, n_hit, n_ionev, n_coll, n_expired, pad;
};

struct CollState {
    bool   ready = false;
    int    mode = 0, ion_life_time = 0;
    double n_d = 0.0, cyl_radius = 0.0;
    double *tab = nullptr;
    int    n_tot = 0, n_ion = 0;
    double4 *erec = nullptr; int *eidx = nullptr, *ions = nullptr; int list_cap = 0;
    int2   *cand = nullptr; int cand_cap = 0;
    rb2_recomb_record *hits = nullptr; int hit_cap = 0;
    rb2_ionization_record *ionev = nullptr; int ionev_cap = 0;
    CollCounts *d_cnt = nullptr, *h_cnt = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::vector<rb2_recomb_record> host_recomb;
    std::vector<rb2_ionization_record> host_ion;
    rb2_collision_result last{};
};
CollState g_coll;

// ---- Update_Collision_Data pieces ----------------------------------------------------------------------------
// BinarySearch, src/mod_global.F90:649-693 (0-based; the midpoint is the reference's 1-based (first+last)/2)
__device__ void binary_search(const double *list, int n, double value, int &i1, int &i2)
{
    int first = 0, last = n - 1;
    if ((last - first) >= 1) {
        while ((last - first) != 1) {
            const int mid = ((first + 1) + (last + 1)) / 2 - 1;
            if (list[mid] > value) last = mid;
            else first = mid;
        }
    }
    i1 = first;
    i2 = last;
}
__device__ double cross_interp(const double *en, const double *dat, int n, double energy)
{
    double e = fmin(fmax(energy, en[0]), en[n - 1]);
    int i1, i2;
    binary_search(en, n, e, i1, i2);
    const double y1 = dat[i1], y2 = dat[i2], x1 = en[i1], x2 = en[i2];
    const double h = (y1 - y2) / (x1 - x2);
    const double q = (y2 * x1 - y1 * x2) / (x1 - x2);
    return (h * e + q) * 1.0e-20;
}
// Find_Cross_tot_data / Find_Cross_ion_data, src/mod_collisions.F90:1985-2044
__device__ double find_cross_tot(const CollParams &P, double energy)
{
    if ((energy > 70.0) && (energy <= 3000.0)) return (7.98 * exp(-0.005845 * energy) + 4.628 * exp(-0.0007864 * energy)) * 1.0e-20;
    return cross_interp(P.tot_e, P.tot_d, P.n_tot, energy);
}
__device__ double find_cross_ion(const CollParams &P, double energy)
{
    if ((energy > 180.0) && (energy <= 3000.0)) return (2.251 * exp(-0.00311 * energy) + 1.04 * exp(-0.0003378 * energy)) * 1.0e-20;
    return cross_interp(P.ion_e, P.ion_d, P.n_ion, energy);
}
// Calculate_Kramers_Cross_Section, :1443-1449
__device__ __forceinline__ double kramers(const CollParams &P, double energy)
{
    const double Z2 = P.Z_eff * P.Z_eff;
    return 2.105e-26 * (P.Ryd * P.Ryd) * (Z2 * Z2) / (P.N_n * energy * ((P.N_n * P.N_n) * energy + P.Ryd * Z2));
}
__device__ __forceinline__ double speed2_of(double vx, double vy, double vz)
{
    const double nrm = sqrt(vx * vx + vy * vy + vz * vz);  // norm2(..)**2, :2110
    return nrm * nrm;
}
__device__ __forceinline__ double recom_radius(const CollParams &P, double vx, double vy, double vz)
{
    const double e = 0.5 * rb2k::m_0 * speed2_of(vx, vy, vz) / rb2k::q_0;
    return sqrt(kramers(P, e) / RB2_PI);
}
// Update_Collision_Data, :2103-2134: out = {cur_energy, ion_cross_sec, ion_cross_rad, recom_cross_rad, tot_cross_sec}
__device__ void collision_data(const CollParams &P, double vx, double vy, double vz, double out[5])
{
    const double elec_max_speed2 = (2.0 * rb2k::q_0 * 5000.0 / rb2k::m_0);
    const double s2 = speed2_of(vx, vy, vz);
    double e;
    if (s2 > elec_max_speed2) e = 0.5 * rb2k::m_0 * elec_max_speed2 / rb2k::q_0;
    else e = 0.5 * rb2k::m_0 * s2 / rb2k::q_0;
    out[1] = find_cross_ion(P, e);
    out[2] = sqrt(out[1] / RB2_PI);
    out[4] = find_cross_tot(P, e);
    e = 0.5 * rb2k::m_0 * s2 / rb2k::q_0;
    out[0] = e;
    out[3] = sqrt(kramers(P, e) / RB2_PI);
}

__global__ void k_coll_data(int n, const double *__restrict__ vel, const int *__restrict__ species, CollParams P,
                            double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double o[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (species[i] == RB2_SPECIES_ELEC) collision_data(P, vel[3 * i], vel[3 * i + 1], vel[3 * i + 2], o);  // :2160
#pragma unroll
    for (int k = 0; k < 5; ++k) out[5 * (size_t)i + k] = o[k];
}

// ---- src/mod_polynomialroots.F90 on the device -------------------------------------------------------------------
struct Cx { double re, im; };
#define POLY_EPS 2.220446049250313e-16
__device__ __forceinline__ double f_sign(double a, double b) { return copysign(fabs(a), b); }
__device__ __forceinline__ void swapd(double &a, double &b) { const double t = b; b = a; a = t; }
__device__ __forceinline__ double cube_root(double x)  // CubeRoot, :61-77
{
    if (x < 0.0) return -exp(log(-x) / 3.0);
    if (x > 0.0) return exp(log(x) / 3.0);
    return 0.0;
}
// QuadraticRoots, :125-175
__device__ void quadratic_roots(const double *a, Cx *z, int &code)
{
    if (a[0] == 0.0) { z[0] = {0.0, 0.0}; z[1] = {-a[1] / a[2], 0.0}; code = 21; return; }
    const double d = a[1] * a[1] - 4.0 * a[0] * a[2];
    if (fabs(d) <= 2.0 * POLY_EPS * a[1] * a[1]) { z[0] = {-0.5 * a[1] / a[2], 0.0}; z[1] = z[0]; code = 22; return; }
    const double r = sqrt(fabs(d));
    if (d < 0.0) {
        const double x = -0.5 * a[1] / a[2], y = fabs(0.5 * r / a[2]);
        z[0] = {x, y}; z[1] = {x, -y}; code = 23;
        return;
    }
    if (a[1] != 0.0) {
        const double w = -(a[1] + f_sign(r, a[1]));
        z[0] = {2.0 * a[0] / w, 0.0}; z[1] = {0.5 * w / a[2], 0.0}; code = 22;
        return;
    }
    const double x = fabs(0.5 * r / a[2]);
    z[0] = {x, 0.0}; z[1] = {-x, 0.0}; code = 22;
}
// CubicRoots, :178-333 (labels keep the numbers of the Fortran statement labels)
__device__ void cubic_roots(const double *a, Cx *z, int &code)
{
    const double RT3 = 1.7320508075689;
    double aq[3], arg, c, cf, d, p, p1, q, q1, r, ra, rb, rq, rt, r1, s, sf, sq, sum, t, tol, t1, w, w1, w2;
    double x, x1, x2, x3, y, y1, y2, y3;
    if (a[0] == 0.0) { z[0] = {0.0, 0.0}; quadratic_roots(a + 1, z + 1, code); return; }
    p = a[2] / (3.0 * a[3]);
    q = a[1] / a[3];
    r = a[0] / a[3];
    tol = 4.0 * POLY_EPS;
    c = 0.0;
    t = a[1] - p * a[2];
    if (fabs(t) > tol * fabs(a[1])) c = t / a[3];
    t = 2.0 * p * p - q;
    if (fabs(t) <= tol * fabs(q)) t = 0.0;
    d = r + p * t;
    if (fabs(d) <= tol * fabs(r)) goto L110;

    s = fmax(fmax(fabs(a[0]), fabs(a[1])), fabs(a[2]));
    p1 = a[2] / (3.0 * s);
    q1 = a[1] / s;
    r1 = a[0] / s;
    t1 = q - 2.25 * p * p;
    if (fabs(t1) <= tol * fabs(q)) t1 = 0.0;
    w = 0.25 * r1 * r1;
    w1 = 0.5 * p1 * r1 * t;
    w2 = q1 * q1 * t1 / 27.0;
    if (w1 >= 0.0) { w = w + w1; sq = w + w2; }
    else if (w2 < 0.0) { sq = w + (w1 + w2); }
    else { w = w + w2; sq = w + w1; }
    if (fabs(sq) <= tol * w) sq = 0.0;
    rq = fabs(s / a[3]) * sqrt(fabs(sq));
    if (sq >= 0.0) goto L40;

    arg = atan2(rq, -0.5 * d);  // all roots are real
    cf = cos(arg / 3.0);
    sf = sin(arg / 3.0);
    rt = sqrt(-c / 3.0);
    y1 = 2.0 * rt * cf;
    y2 = -rt * (cf + RT3 * sf);
    y3 = -(d / y1) / y2;
    x1 = y1 - p;
    x2 = y2 - p;
    x3 = y3 - p;
    if (fabs(x1) > fabs(x2)) swapd(x1, x2);
    if (fabs(x2) > fabs(x3)) swapd(x2, x3);
    if (fabs(x1) > fabs(x2)) swapd(x1, x2);
    w = x3;
    if (fabs(x2) < 0.1 * fabs(x3)) goto L70;
    if (fabs(x1) < 0.1 * fabs(x2)) x1 = -(r / x3) / x2;
    z[0] = {x1, 0.0}; z[1] = {x2, 0.0}; z[2] = {x3, 0.0};
    return;

L40:  // real and complex roots
    ra = cube_root(-0.5 * d - f_sign(rq, d));
    rb = -c / (3.0 * ra);
    t = ra + rb;
    w = -p;
    x = -p;
    if (fabs(t) <= tol * fabs(ra)) goto L41;
    w = t - p;
    x = -0.5 * t - p;
    if (fabs(x) <= tol * fabs(p)) x = 0.0;
L41:
    t = fabs(ra - rb);
    y = 0.5 * RT3 * t;
    if (t <= tol * fabs(ra)) goto L60;
    if (fabs(x) < fabs(y)) goto L50;
    s = fabs(x);
    t = y / x;
    goto L51;
L50:
    s = fabs(y);
    t = x / y;
L51:
    if (s < 0.1 * fabs(w)) goto L70;
    w1 = w / s;
    sum = 1.0 + t * t;
    if (w1 * w1 < 0.01 * sum) w = -((r / sum) / s) / s;
    z[0] = {w, 0.0}; z[1] = {x, y}; z[2] = {x, -y};
    return;

L60:  // at least two roots are equal
    if (fabs(x) < fabs(w)) goto L61;
    if (fabs(w) < 0.1 * fabs(x)) w = -(r / x) / x;
    z[0] = {w, 0.0}; z[1] = {x, 0.0}; z[2] = z[1];
    return;
L61:
    if (fabs(x) < 0.1 * fabs(w)) goto L70;
    z[0] = {x, 0.0}; z[1] = z[0]; z[2] = {w, 0.0};
    return;

L70:  // w is much larger in magnitude than the other roots
    aq[0] = a[0];
    aq[1] = a[1] + a[0] / w;
    aq[2] = -a[3] * w;
    quadratic_roots(aq, z, code);
    z[2] = {w, 0.0};
    if (z[0].im == 0.0) return;
    z[2] = z[1];
    z[1] = z[0];
    z[0] = {w, 0.0};
    return;

L110:  // case when d = 0
    z[0] = {-p, 0.0};
    w = sqrt(fabs(c));
    if (c < 0.0) goto L120;
    z[1] = {-p, w}; z[2] = {-p, -w};
    return;
L120:
    if (p != 0.0) goto L130;
    z[1] = {w, 0.0}; z[2] = {-w, 0.0};
    return;
L130:
    x = -(p + f_sign(w, p));
    z[2] = {x, 0.0};
    t = 3.0 * a[0] / (a[2] * x);
    if (fabs(p) > fabs(t)) goto L131;
    z[1] = {t, 0.0};
    return;
L131:
    z[1] = z[0];
    z[0] = {t, 0.0};
}

// principal square root of a complex number with a non-zero imaginary part (Fortran SQRT on COMPLEX(DP))
__device__ __forceinline__ Cx csqrt_dev(Cx v)
{
    const double m = hypot(v.re, v.im);
    Cx w;
    if (v.re >= 0.0) {
        w.re = sqrt(0.5 * (m + v.re));
        w.im = v.im / (2.0 * w.re);
    } else {
        const double s = sqrt(0.5 * (m - v.re));
        w.re = fabs(v.im) / (2.0 * s);
        w.im = copysign(s, v.im);
    }
    return w;
}

// QuarticRoots, :336-510
__device__ void quartic_roots(const double *a, Cx *z, int &code)
{
    Cx w;
    double b, b2, c, d, e, h, p, q, r, t, temp[4], u, v, v1, v2, x, x1, x2, x3, y;
    if (a[0] == 0.0) { z[0] = {0.0, 0.0}; cubic_roots(a + 1, z + 1, code); return; }
    b = a[3] / (4.0 * a[4]);
    c = a[2] / a[4];
    d = a[1] / a[4];
    e = a[0] / a[4];
    b2 = b * b;
    p = 0.5 * (c - 6.0 * b2);
    q = d - 2.0 * b * (c - 4.0 * b2);
    r = b2 * (c - 3.0 * b2) - b * d + e;
    temp[0] = -q * q / 64.0;
    temp[1] = 0.25 * (p * p - r);
    temp[2] = p;
    temp[3] = 1.0;
    cubic_roots(temp, z, code);
    if (z[1].im != 0.0) goto L60;

    x1 = z[0].re;  // the resolvent cubic has only real roots
    x2 = z[1].re;
    x3 = z[2].re;
    if (x1 > x2) swapd(x1, x2);
    if (x2 > x3) swapd(x2, x3);
    if (x1 > x2) swapd(x1, x2);
    u = 0.0;
    if (x3 > 0.0) u = sqrt(x3);
    if (x2 <= 0.0) goto L41;
    if (x1 >= 0.0) goto L30;
    if (fabs(x1) > x2) goto L40;
    x1 = 0.0;
L30:
    x1 = sqrt(x1);
    x2 = sqrt(x2);
    if (q > 0.0) x1 = -x1;
    temp[0] = ((x1 + x2) + u) - b;
    temp[1] = ((-x1 - x2) + u) - b;
    temp[2] = ((x1 - x2) - u) - b;
    temp[3] = ((-x1 + x2) - u) - b;
    for (int j = 0; j < 3; ++j) {  // SelectSort, :512-526
        int k = j;
        for (int m = j + 1; m < 4; ++m) if (temp[m] < temp[k]) k = m;
        if (j != k) swapd(temp[k], temp[j]);
    }
    if (fabs(temp[0]) >= 0.1 * fabs(temp[3])) goto L31;
    t = temp[1] * temp[2] * temp[3];
    if (t != 0.0) temp[0] = e / t;
L31:
    z[0] = {temp[0], 0.0}; z[1] = {temp[1], 0.0}; z[2] = {temp[2], 0.0}; z[3] = {temp[3], 0.0};
    code = 31;
    return;
L40:
    v1 = sqrt(fabs(x1));
    v2 = 0.0;
    goto L50;
L41:
    v1 = sqrt(fabs(x1));
    v2 = sqrt(fabs(x2));
    if (q < 0.0) u = -u;
L50:
    x = -u - b;
    y = v1 - v2;
    z[0] = {x, y}; z[1] = {x, -y};
    x = u - b;
    y = v1 + v2;
    z[2] = {x, y}; z[3] = {x, -y};
    code = 44;
    return;

L60:  // the resolvent cubic has complex roots
    t = z[0].re;
    x = 0.0;
    if (t < 0.0) goto L61;
    else if (t == 0.0) goto L70;
    else goto L62;
L61:
    h = fabs(z[1].re) + fabs(z[1].im);
    if (fabs(t) <= h) goto L70;
    goto L80;
L62:
    x = sqrt(t);
    if (q > 0.0) x = -x;
L70:
    w = csqrt_dev(z[1]);
    u = 2.0 * w.re;
    v = 2.0 * fabs(w.im);
    t = x - b;
    x1 = t + u;
    x2 = t - u;
    if (fabs(x1) <= fabs(x2)) goto L71;
    t = x1;
    x1 = x2;
    x2 = t;
L71:
    u = -x - b;
    h = u * u + v * v;
    if (x1 * x1 < 0.01 * fmin(x2 * x2, h)) x1 = e / (x2 * h);
    z[0] = {x1, 0.0}; z[1] = {x2, 0.0}; z[2] = {u, v}; z[3] = {u, -v};
    code = 42;
    return;
L80:
    v = sqrt(fabs(t));
    z[0] = {-b, v}; z[1] = {-b, -v}; z[2] = z[0]; z[3] = z[1];
    code = 23;
}

// One (ion, electron) test, src/mod_collisions.F90:128-196.  The quartic coefficient is non-zero whenever the
// electron has an acceleration; with a == 0 the reference falls through to CubicRoots with a stale return code
// (module variable), which has no counterpart on a parallel machine: such a pair reports no entry.
__device__ bool recomb_pair(const double ion[3], const double ep[3], const double ev[3], const double ea[3], double recom_rad,
                            double dt, double &t_out, double &dist_out)
{
    const double rel[3] = {ep[0] - ion[0], ep[1] - ion[1], ep[2] - ion[2]};
    const double cur_dist2 = rel[0] * rel[0] + rel[1] * rel[1] + rel[2] * rel[2];
    const double recom_rad2 = recom_rad * recom_rad;
    bool hit = false;
    double t = 0.0;
    if (cur_dist2 <= recom_rad2) {
        hit = true;
    } else {
        double co[5];
        co[4] = 0.25 * (ea[0] * ea[0] + ea[1] * ea[1] + ea[2] * ea[2]);
        co[3] = ev[0] * ea[0] + ev[1] * ea[1] + ev[2] * ea[2];
        co[2] = (ev[0] * ev[0] + ev[1] * ev[1] + ev[2] * ev[2]) + (rel[0] * ea[0] + rel[1] * ea[1] + rel[2] * ea[2]);
        co[1] = 2.0 * (rel[0] * ev[0] + rel[1] * ev[1] + rel[2] * ev[2]);
        co[0] = cur_dist2 - recom_rad2;
        if (co[4] != 0.0) {
            Cx z[5];
            int code = 0;
            quartic_roots(co, z, code);
            // :153-185: code 31 looks at root1..root3 (root4 is never assigned by SolvePolynomial, :561), 42 at root1, root2
            const int nroots = (code == 31) ? 3 : ((code == 42) ? 2 : 0);
            for (int k = 0; k < nroots; ++k)
                if (z[k].im == 0.0 && z[k].re > 0.0 && z[k].re <= dt) { hit = true; t = z[k].re; break; }
        }
    }
    if (!hit) return false;
    double nx[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) nx[c] = (ep[c] + ev[c] * t + 0.5 * ea[c] * (t * t)) - ion[c];
    t_out = t;
    dist_out = sqrt(nx[0] * nx[0] + nx[1] * nx[1] + nx[2] * nx[2]);
    return true;
}

// Test hook: QuarticRoots on the device for a list of polynomials (the solver is otherwise only reachable through the
// recombination test).  coeffs = {quartic, cubic, quadratic, linear, constant} per polynomial, like SolvePolynomial.
__global__ void k_quartic_probe(int n, const double *__restrict__ coeffs, int *__restrict__ codes, double *__restrict__ roots)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a[5] = {coeffs[5 * i + 4], coeffs[5 * i + 3], coeffs[5 * i + 2], coeffs[5 * i + 1], coeffs[5 * i]};
    Cx z[5];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int k = 0; k < 5; ++k) z[k] = {nan, nan};
    int code = 0;
    if (a[4] != 0.0) quartic_roots(a, z, code);
    codes[i] = code;
    for (int k = 0; k < 4; ++k) { roots[8 * i + 2 * k] = z[k].re; roots[8 * i + 2 * k + 1] = z[k].im; }
}

// ---- recombination kernels -------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_append(int *counter, bool want)
{
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (m == 0u) return -1;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return want ? base + __popc(m & ((1u << lane) - 1u)) : -1;
}

__global__ void __launch_bounds__(256)
k_recomb_prepare(int n, DevArrays A, int *__restrict__ mask, DevCounters *__restrict__ C, CollParams P,
                 double4 *__restrict__ erec, int *__restrict__ eidx, int *__restrict__ ions, CollCounts *__restrict__ K)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool is_e = false, is_i = false;
    double4 rec = make_double4(0.0, 0.0, 0.0, 0.0);
    if (i < n && mask[i] != 0) {
        const int sp = A.species[i];
        if (sp == RB2_SPECIES_ION) {
            if (P.step >= A.life[i]) {  // end of life: Mark_Particles_Remove(i, remove_top), :121-124
                mask[i] = 0;
                A.pq[i].w = 0.0;
                atomicAdd(&C->mark_part, 1); atomicAdd(&C->mark_ion, 1);
                atomicAdd(&C->top_part, 1); atomicAdd(&C->top_ion, 1);
                atomicAdd(&K->n_expired, 1);
            } else {
                is_i = true;
            }
        } else if (sp == RB2_SPECIES_ELEC) {
            is_e = true;
            const double4 p = A.pq[i];
            const double vx = A.vel[3 * i], vy = A.vel[3 * i + 1], vz = A.vel[3 * i + 2];
            const double ax = A.acc[3 * i], ay = A.acc[3 * i + 1], az = A.acc[3 * i + 2];
            const double R = recom_radius(P, vx, vy, vz);
            const double travel = sqrt(vx * vx + vy * vy + vz * vz) * P.dt + 0.5 * sqrt(ax * ax + ay * ay + az * az) * (P.dt * P.dt);
            const double b = (R + 2.0 * travel) * (1.0 + 1.0e-9);
            rec = make_double4(p.x, p.y, p.z, b * b);
        }
    }
    const int ke = warp_append(&K->n_elec, is_e);
    if (is_e) { erec[ke] = rec; eidx[ke] = i; }
    const int ki = warp_append(&K->n_ion, is_i);
    if (is_i) ions[ki] = i;
}

// IPT ions per thread: every electron record read from shared memory (two LDS.128) serves IPT distance tests; with one
// ion per thread the shared-memory pipe, not the FP64 pipe, was the limiter (ncu: LSU wavefronts 77 %, FP64 68 %).
template <int IPT>
__global__ void __launch_bounds__(TPB)
k_recomb_sweep(const double4 *__restrict__ pq, const int *__restrict__ ions, const double4 *__restrict__ erec,
               const int *__restrict__ eidx, int e_chunk, int2 *__restrict__ cand, int cand_cap, CollCounts *__restrict__ K)
{
    __shared__ double4 tile[ETILE];
    const int n_ion = K->n_ion, n_elec = K->n_elec;
    if ((int)(blockIdx.x * TPB * IPT) >= n_ion) return;
    const int j0 = blockIdx.y * e_chunk, j1 = min(n_elec, j0 + e_chunk);
    int ion[IPT];
    double x[IPT], y[IPT], z[IPT];
#pragma unroll
    for (int u = 0; u < IPT; ++u) {
        const int k = (blockIdx.x * IPT + u) * TPB + threadIdx.x;
        ion[u] = -1;
        x[u] = y[u] = z[u] = 1.0e30;  // no electron record is within reach of a padding lane
        if (k < n_ion) {
            ion[u] = ions[k];
            const double4 p = pq[ion[u]];
            x[u] = p.x; y[u] = p.y; z[u] = p.z;
        }
    }
    for (int t0 = j0; t0 < j1; t0 += ETILE) {
        __syncthreads();
        for (int s = threadIdx.x; s < ETILE; s += TPB)
            tile[s] = (t0 + s < j1) ? erec[t0 + s] : make_double4(0.0, 0.0, 0.0, -1.0);
        __syncthreads();
        const int cnt = min(ETILE, j1 - t0);
#pragma unroll 4
        for (int s = 0; s < cnt; ++s) {
            const double4 e = tile[s];
#pragma unroll
            for (int u = 0; u < IPT; ++u) {
                const double dx = e.x - x[u], dy = e.y - y[u], dz = e.z - z[u];
                const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                if (d2 <= e.w) {  // rare (an infinite bound -- electron at rest -- admits every ion, like the reference)
                    if (ion[u] >= 0) {
                        const int slot = atomicAdd(&K->n_cand, 1);
                        if (slot < cand_cap) cand[slot] = make_int2(ion[u], eidx[t0 + s]);
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(TPB)
k_recomb_solve(DevArrays A, const int2 *__restrict__ cand, int cand_cap, CollParams P, rb2_recomb_record *__restrict__ hits,
               CollCounts *__restrict__ K)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int nc = min(K->n_cand, cand_cap);
    if (k >= nc) return;
    const int i = cand[k].x, j = cand[k].y;
    const double4 pi = A.pq[i], pj = A.pq[j];
    const double ion[3] = {pi.x, pi.y, pi.z}, ep[3] = {pj.x, pj.y, pj.z};
    const double ev[3] = {A.vel[3 * j], A.vel[3 * j + 1], A.vel[3 * j + 2]};
    const double ea[3] = {A.acc[3 * j], A.acc[3 * j + 1], A.acc[3 * j + 2]};
    const double R = recom_radius(P, ev[0], ev[1], ev[2]) * 1.0;  // multiplicator = 1, :96
    double t, dist;
    if (!recomb_pair(ion, ep, ev, ea, R, P.dt, t, dist)) return;
    const int slot = atomicAdd(&K->n_hit, 1);
    rb2_recomb_record r;
    r.step = P.step; r.elec_slot = j; r.ion_slot = i;
    r.elec_emit = A.emitter[j]; r.ion_life = P.step - A.step[i];
    r.elec_sec = A.section[j]; r.elec_id = A.id[j];
    r.ion_emit = A.emitter[i]; r.ion_sec = A.section[i]; r.ion_id = A.id[i];
    for (int c = 0; c < 3; ++c) { r.ion_pos[c] = ion[c]; r.elec_pos[c] = ep[c]; }
    r.elec_speed = sqrt(ev[0] * ev[0] + ev[1] * ev[1] + ev[2] * ev[2]);
    r.dist = dist; r.recom_rad = R; r.t = t;
    hits[slot] = r;  // n_hit <= n_cand <= cand_cap == hit capacity
}

// ---- ionisation ----------------------------------------------------------------------------------------------------------
struct Philox {
    uint2 key; uint4 ctr; uint4 buf; int have;
    __device__ Philox(unsigned long long seed, int step, int slot)
    {
        key = make_uint2((unsigned)seed ^ ((unsigned)slot * 0x9E3779B1u), (unsigned)(seed >> 32) + (unsigned)step * 0x85EBCA6Bu);
        ctr = make_uint4(0u, 0x636f6c6cu, (unsigned)step, (unsigned)slot);
        have = 0;
    }
    __device__ uint4 round10(uint4 c, uint2 k) const
    {
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
            const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
            c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
            k.x += 0x9E3779B9u;
            k.y += 0xBB67AE85u;
        }
        return c;
    }
    __device__ double next()  // uniform in [0, 1), 53 bits
    {
        unsigned hi, lo;
        if (have) { hi = buf.z; lo = buf.w; have = 0; }
        else { buf = round10(ctr, key); ctr.x += 1u; hi = buf.x; lo = buf.y; have = 1; }
        return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
    }
};

// folded_normal_dist / folded_normal_max, src/mod_collisions.F90:1880-1904
__device__ __forceinline__ double folded_normal_dist(double mu, double sigma, double x)
{
    const double sigma2 = sigma * sigma;
    return sqrt(2.0 / (RB2_PI * sigma2)) * exp(-1.0 * (mu * mu + x * x) / (2.0 * sigma2)) * cosh(mu * x / sigma2);
}
// The reference scans a 0.1 degree grid over [0, 180] for the largest value (1801 evaluations).  The folded normal is
// unimodal on that interval (rising to its mode, then falling; the mode is 0 for mu <= sigma), so the first grid point
// whose successor is smaller is the grid maximum: 11 bisection steps instead of the scan -- the scan alone made one
// ionising thread the critical path of the whole kernel (0.7 ms).
__device__ double folded_normal_max(double mu, double sigma)
{
    int lo = 0, hi = 1800;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (folded_normal_dist(mu, sigma, 180.0 * (mid + 1) / 1800.0) < folded_normal_dist(mu, sigma, 180.0 * mid / 1800.0)) hi = mid;
        else lo = mid + 1;
    }
    return folded_normal_dist(mu, sigma, 180.0 * lo / 1800.0);
}
// The acceptance-rejection loop shared by Get_Injected_Vec (:1452-1517) and Get_Ejected_Vec (:1525-1577)
__device__ void scatter_direction(Philox &g, double mu, double sigma, double m_factor, const double pv[3], double out[3])
{
    const double len_vel = sqrt(pv[0] * pv[0] + pv[1] * pv[1] + pv[2] * pv[2]);
    double v[3], len_vec;
    int n_tries = 0;
    for (;;) {
        v[0] = g.next() - 0.5; v[1] = g.next() - 0.5; v[2] = g.next() - 0.5;
        len_vec = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        double alpha = 1.0;
        if ((len_vel > 0.0) && (len_vec > 0.0)) {
            const double dot_p = pv[0] * v[0] + pv[1] * v[1] + pv[2] * v[2];
            const double angle = acos(dot_p / (len_vec * len_vel)) * 180.0 / RB2_PI;
            alpha = folded_normal_dist(mu, sigma, angle) / m_factor;
        }
        if (g.next() < alpha) break;
        if (++n_tries >= 1000000) break;
    }
    out[0] = v[0] / len_vec; out[1] = v[1] / len_vec; out[2] = v[2] / len_vec;
}

__global__ void __launch_bounds__(TPB)
k_ionize(int n, DevArrays A, const int *__restrict__ mask, CollParams P, rb2_ionization_record *__restrict__ ev, int ev_cap,
         CollCounts *__restrict__ K)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if ((A.species[i] != RB2_SPECIES_ELEC) || (mask[i] == 0)) return;
    const double4 p = A.pq[i];
    if (!(sqrt(p.x * p.x + p.y * p.y) <= P.cyl_radius)) return;  // :594
    const double pv[3] = {A.vel[3 * i], A.vel[3 * i + 1], A.vel[3 * i + 2]};
    double cd[5];
    collision_data(P, pv[0], pv[1], pv[2], cd);
    const double elec_energy = cd[0];
    if (!(elec_energy > P.N_bind)) return;
    const double dx = p.x - A.prev_pos[3 * i], dy = p.y - A.prev_pos[3 * i + 1], dz = p.z - A.prev_pos[3 * i + 2];
    const double elec_cur_path = sqrt(dx * dx + dy * dy + dz * dz);
    const double elec_cur_speed = sqrt(pv[0] * pv[0] + pv[1] * pv[1] + pv[2] * pv[2]);
    const double cross_tot = cd[4];
    const double mean_path = 1.0 / (P.n_d * cross_tot);
    Philox g(P.seed, P.step, i);
    if (!(g.next() < elec_cur_path / mean_path)) return;
    atomicAdd(&K->n_coll, 1);
    if (!(g.next() < cd[1] / cross_tot)) return;  // ionising or not, :619-624

    rb2_ionization_record r;
    r.step = P.step; r.in_slot = i; r.new_id = -1; r.ion_id = -1; r.elec_emit = A.emitter[i]; r.pad = 0;
    r.pos[0] = p.x; r.pos[1] = p.y; r.pos[2] = p.z;
    const double E1 = elec_energy, E2 = E1 - P.N_bind;
    const double collE = E2 * g.next();
    const double ejecE = E2 - collE;
    r.E1 = E1; r.collE = collE; r.ejecE = ejecE;
    double dir[3], nrm;
    // colliding electron: Get_Injected_Vec (mu = 5, sigma = 25 degrees)
    scatter_direction(g, 5.0, 25.0, P.inj_max, pv, dir);
    nrm = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    const double s_in = sqrt(2.0 * rb2k::q_0 * collE / rb2k::m_0);
#pragma unroll
    for (int c = 0; c < 3; ++c) r.new_vel[c] = (dir[c] / nrm) * s_in;
    r.in_speed = elec_cur_speed;
    r.out_speed = sqrt(r.new_vel[0] * r.new_vel[0] + r.new_vel[1] * r.new_vel[1] + r.new_vel[2] * r.new_vel[2]);
    // ejected electron
    const double pc[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) r.ejec_pos[c] = pc[c] + (2.0 * (g.next() - 0.5)) * rb2k::length_scale;
    {
        const double T = elec_energy;
        const double angle_max = (T < 100.0) ? (-430.5 * pow(100.0, -0.5445) + 89.32) : (-430.5 * pow(T, -0.5445) + 89.32);
        scatter_direction(g, angle_max, 48.0, folded_normal_max(angle_max, 48.0), pv, dir);
    }
    nrm = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    const double s_ej = sqrt(2.0 * rb2k::q_0 * ejecE / rb2k::m_0);
#pragma unroll
    for (int c = 0; c < 3; ++c) r.ejec_vel[c] = (dir[c] / nrm) * s_ej;
    r.new_speed = sqrt(r.ejec_vel[0] * r.ejec_vel[0] + r.ejec_vel[1] * r.ejec_vel[1] + r.ejec_vel[2] * r.ejec_vel[2]);
    // created ion
#pragma unroll
    for (int c = 0; c < 3; ++c) r.ion_pos[c] = pc[c] + (2.0 * (g.next() - 0.5)) * rb2k::length_scale;
    const int slot = atomicAdd(&K->n_ionev, 1);
    if (slot < ev_cap) ev[slot] = r;
}

// The state changes of the colliding electrons (:635, :692), applied once the host knows that every record fitted
// (k_ionize itself changes nothing, so that it can simply be run again with a larger buffer).
__global__ void k_ionize_apply(int ne, const rb2_ionization_record *__restrict__ ev, DevArrays A)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ne) return;
    const int i = ev[k].in_slot;
    for (int c = 0; c < 3; ++c) A.vel[3 * i + c] = ev[k].new_vel[c];
    A.emitter[i] = 2;  // ion_emitter, src/mod_global.F90:121
}

// ---- host side ----------------------------------------------------------------------------------------------------------------
double host_folded_normal_max(double mu, double sigma)
{
    double best = 0.0;
    const double sigma2 = sigma * sigma;
    for (int k = 0; k <= 1800; ++k) {
        const double x = 180.0 * k / 1800.0;
        const double f = sqrt(2.0 / (RB2_PI * sigma2)) * exp(-1.0 * (mu * mu + x * x) / (2.0 * sigma2)) * cosh(mu * x / sigma2);
        if (f > best) best = f;
    }
    return best;
}

CollParams make_params(const Rb2Ctx &c, int step, unsigned long long seed)
{
    const CollState &S = g_coll;
    CollParams P{};
    P.tot_e = S.tab; P.tot_d = S.tab + S.n_tot; P.ion_e = S.tab + 2 * (size_t)S.n_tot; P.ion_d = P.ion_e + S.n_ion;
    P.n_tot = S.n_tot; P.n_ion = S.n_ion;
    // src/mod_global.F90:50-68
    const double h = 6.62607015e-34;
    const double R_inf = rb2k::m_0 * (rb2k::q_0 * rb2k::q_0 * rb2k::q_0 * rb2k::q_0) /
                         (8.0 * (rb2k::epsilon_0 * rb2k::epsilon_0) * (h * h * h) * rb2k::c);
    P.Ryd = h * rb2k::c * R_inf / rb2k::q_0;
    P.N_n = 2.0;
    P.N_bind = 15.581;
    P.Z_eff = sqrt(P.N_bind * (P.N_n * P.N_n) / P.Ryd);
    P.n_d = S.n_d; P.cyl_radius = S.cyl_radius; P.dt = c.cfg.time_step;
    P.step = step; P.ion_life_time = S.ion_life_time; P.seed = seed;
    static const double inj_max = host_folded_normal_max(5.0, 25.0);
    P.inj_max = inj_max;
    return P;
}

int ensure_lists(Rb2Ctx &c)
{
    CollState &S = g_coll;
    if (S.list_cap >= c.cap) return RB2_OK;
    cudaFree(S.erec); cudaFree(S.eidx); cudaFree(S.ions);
    S.erec = nullptr; S.eidx = nullptr; S.ions = nullptr; S.list_cap = 0;
    RB2_CUDA(cudaMalloc(&S.erec, (size_t)c.cap * sizeof(double4)));
    RB2_CUDA(cudaMalloc(&S.eidx, (size_t)c.cap * sizeof(int)));
    RB2_CUDA(cudaMalloc(&S.ions, (size_t)c.cap * sizeof(int)));
    S.list_cap = c.cap;
    return RB2_OK;
}
int ensure_cand(int want)
{
    CollState &S = g_coll;
    if (S.cand_cap >= want) return RB2_OK;
    cudaFree(S.cand); cudaFree(S.hits);
    S.cand = nullptr; S.hits = nullptr; S.cand_cap = S.hit_cap = 0;
    RB2_CUDA(cudaMalloc(&S.cand, (size_t)want * sizeof(int2)));
    RB2_CUDA(cudaMalloc(&S.hits, (size_t)want * sizeof(rb2_recomb_record)));
    S.cand_cap = S.hit_cap = want;
    return RB2_OK;
}
int ensure_ionev(int want)
{
    CollState &S = g_coll;
    if (S.ionev_cap >= want) return RB2_OK;
    cudaFree(S.ionev);
    S.ionev = nullptr; S.ionev_cap = 0;
    RB2_CUDA(cudaMalloc(&S.ionev, (size_t)want * sizeof(rb2_ionization_record)));
    S.ionev_cap = want;
    return RB2_OK;
}

void fill_result(Rb2Ctx &c, rb2_collision_result *out)
{
    CollState &S = g_coll;
    S.last.nrPart_remove_recom = c.h_counters->recom_part;
    S.last.nrElec_remove_recom = c.h_counters->recom_elec;
    S.last.nrIon_remove_recom = c.h_counters->recom_ion;
    S.last.counts = c.counts;
    if (out) *out = S.last;
}

int require_ready()
{
    if (!g_rb2.init) return rb2_fail(RB2_ERR_NOT_INIT, "rb2_init has not been called");
    if (!g_coll.ready) return rb2_fail(RB2_ERR_NOT_INIT, "rb2_collisions_init has not been called");
    return RB2_OK;
}

// Do_Discrete_Recombination_ots on the current store; leaves the counters fetched.
int do_recombination(Rb2Ctx &c, int step)
{
    CollState &S = g_coll;
    S.host_recomb.clear();
    S.last.nrRecombinations = 0; S.last.nrIonsExpired = 0; S.last.n_candidates = 0;
    const int n = c.n;
    if (n < 1) return rb2_fetch_counters(c);
    int rc = ensure_lists(c);
    if (rc) return rc;
    if ((rc = ensure_cand(S.cand_cap > 0 ? S.cand_cap : 65536))) return rc;
    const CollParams P = make_params(c, step, 0ull);
    cudaStream_t st = c.stream;
    RB2_CUDA(cudaMemsetAsync(S.d_cnt, 0, sizeof(CollCounts), st));
    k_recomb_prepare<<<(n + 255) / 256, 256, 0, st>>>(n, c.a, c.mask, c.d_counters, P, S.erec, S.eidx, S.ions, S.d_cnt);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(1);
    RB2_CUDA(cudaMemcpyAsync(S.h_cnt, S.d_cnt, sizeof(CollCounts), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaStreamSynchronize(st));
    const int n_ion = S.h_cnt->n_ion, n_elec = S.h_cnt->n_elec;
    S.last.nrIonsExpired = S.h_cnt->n_expired;
    S.h_cnt->n_cand = 0;
    if (n_ion > 0 && n_elec > 0) {
        // grid: ion blocks x electron chunks (whole tiles), about 8 CTAs per SM
        const int ipt = (n_ion >= 2 * TPB * c.sm_count) ? 2 : 1;  // two ions per thread once that still fills the machine
        const int iblocks = (n_ion + TPB * ipt - 1) / (TPB * ipt);
        int chunks = (8 * c.sm_count + iblocks - 1) / iblocks;
        const int max_chunks = (n_elec + ETILE - 1) / ETILE;
        if (chunks > max_chunks) chunks = max_chunks;
        int e_chunk = (n_elec + chunks - 1) / chunks;
        e_chunk = ((e_chunk + ETILE - 1) / ETILE) * ETILE;
        chunks = (n_elec + e_chunk - 1) / e_chunk;
        for (;;) {
            RB2_CUDA(cudaMemsetAsync(&S.d_cnt->n_cand, 0, 2 * sizeof(int), st));  // n_cand, n_hit
            if (ipt == 2)
                k_recomb_sweep<2><<<dim3(iblocks, chunks), TPB, 0, st>>>(c.a.pq, S.ions, S.erec, S.eidx, e_chunk, S.cand, S.cand_cap, S.d_cnt);
            else
                k_recomb_sweep<1><<<dim3(iblocks, chunks), TPB, 0, st>>>(c.a.pq, S.ions, S.erec, S.eidx, e_chunk, S.cand, S.cand_cap, S.d_cnt);
            RB2_CUDA(cudaGetLastError());
            RB2_LAUNCHED(1);
            RB2_CUDA(cudaMemcpyAsync(S.h_cnt, S.d_cnt, sizeof(CollCounts), cudaMemcpyDeviceToHost, st));
            RB2_CUDA(cudaStreamSynchronize(st));
            if (S.h_cnt->n_cand <= S.cand_cap) break;
            if ((rc = ensure_cand(S.h_cnt->n_cand + S.h_cnt->n_cand / 2))) return rc;  // grow and sweep again
        }
    }
    const int n_cand = S.h_cnt->n_cand;
    S.last.n_candidates = n_cand;
    if (n_cand > 0) {
        k_recomb_solve<<<(n_cand + TPB - 1) / TPB, TPB, 0, st>>>(c.a, S.cand, S.cand_cap, P, S.hits, S.d_cnt);
        RB2_CUDA(cudaGetLastError());
        RB2_LAUNCHED(1);
        RB2_CUDA(cudaMemcpyAsync(S.h_cnt, S.d_cnt, sizeof(CollCounts), cudaMemcpyDeviceToHost, st));
        RB2_CUDA(cudaStreamSynchronize(st));
        const int nh = S.h_cnt->n_hit;
        if (nh > 0) {
            std::vector<rb2_recomb_record> hits((size_t)nh);
            RB2_CUDA(cudaMemcpyAsync(hits.data(), S.hits, (size_t)nh * sizeof(rb2_recomb_record), cudaMemcpyDeviceToHost, st));
            RB2_CUDA(cudaStreamSynchronize(st));
            std::sort(hits.begin(), hits.end(), [](const rb2_recomb_record &a, const rb2_recomb_record &b) {
                return a.ion_slot != b.ion_slot ? a.ion_slot < b.ion_slot : a.elec_slot < b.elec_slot;
            });
            // the serial claim rule: ions in ascending order, each takes its first colliding electron still unclaimed
            std::unordered_set<int> taken;
            std::vector<int> idx, reason;
            int cur_ion = -1;
            bool ion_done = false;
            for (const rb2_recomb_record &h : hits) {
                if (h.ion_slot != cur_ion) { cur_ion = h.ion_slot; ion_done = false; }
                if (ion_done || taken.count(h.elec_slot)) continue;
                taken.insert(h.elec_slot);
                ion_done = true;
                S.host_recomb.push_back(h);
                idx.push_back(h.ion_slot); reason.push_back(RB2_REMOVE_RECOM);   // Mark(i) then Mark(j), :213-216
                idx.push_back(h.elec_slot); reason.push_back(RB2_REMOVE_RECOM);
            }
            rc = rb2_mark_remove((int)idx.size(), idx.data(), reason.data());
            if (rc) return rc;
        }
    }
    S.last.nrRecombinations = (int)S.host_recomb.size();
    return rb2_fetch_counters(c);
}

// Do_Continuous_Ionization_ots on the current store.
int do_ionization(Rb2Ctx &c, int step, unsigned long long seed)
{
    CollState &S = g_coll;
    S.host_ion.clear();
    S.last.nrCollisions = 0; S.last.nrIonizations = 0;
    const int n = c.n;
    if (n < 1 || c.counts.nrElec == 0) return RB2_OK;  // :580
    int rc = ensure_ionev(S.ionev_cap > 0 ? S.ionev_cap : 4096);
    if (rc) return rc;
    const CollParams P = make_params(c, step, seed);
    cudaStream_t st = c.stream;
    int ne = 0;
    for (;;) {
        RB2_CUDA(cudaMemsetAsync(S.d_cnt, 0, sizeof(CollCounts), st));
        k_ionize<<<(n + TPB - 1) / TPB, TPB, 0, st>>>(n, c.a, c.mask, P, S.ionev, S.ionev_cap, S.d_cnt);
        RB2_CUDA(cudaGetLastError());
        RB2_LAUNCHED(1);
        RB2_CUDA(cudaMemcpyAsync(S.h_cnt, S.d_cnt, sizeof(CollCounts), cudaMemcpyDeviceToHost, st));
        RB2_CUDA(cudaStreamSynchronize(st));
        ne = S.h_cnt->n_ionev;
        if (ne <= S.ionev_cap) break;
        if ((rc = ensure_ionev(ne + ne / 2))) return rc;  // same generator keys: the second run finds the same events
    }
    S.last.nrCollisions = S.h_cnt->n_coll;
    S.last.nrIonizations = ne;
    if (ne > 0) {
        k_ionize_apply<<<(ne + 127) / 128, 128, 0, st>>>(ne, S.ionev, c.a);
        RB2_CUDA(cudaGetLastError());
        RB2_LAUNCHED(1);
        S.host_ion.resize((size_t)ne);
        RB2_CUDA(cudaMemcpyAsync(S.host_ion.data(), S.ionev, (size_t)ne * sizeof(rb2_ionization_record), cudaMemcpyDeviceToHost, st));
        RB2_CUDA(cudaStreamSynchronize(st));
        std::sort(S.host_ion.begin(), S.host_ion.end(),
                  [](const rb2_ionization_record &a, const rb2_ionization_record &b) { return a.in_slot < b.in_slot; });
        // Add_Particle(ejected electron) then Add_Particle(ion) per event, :668-685
        std::vector<double> pos((size_t)6 * ne), vel((size_t)6 * ne, 0.0);
        std::vector<int> sp((size_t)2 * ne), emit((size_t)2 * ne, 2), sec((size_t)2 * ne, 1), life((size_t)2 * ne);
        int room = c.cap - c.n, id = c.counts.nrID;
        for (int e = 0; e < ne; ++e) {
            rb2_ionization_record &r = S.host_ion[(size_t)e];
            for (int k = 0; k < 3; ++k) {
                pos[(size_t)6 * e + k] = r.ejec_pos[k]; vel[(size_t)6 * e + k] = r.ejec_vel[k];
                pos[(size_t)6 * e + 3 + k] = r.ion_pos[k];
            }
            sp[(size_t)2 * e] = RB2_SPECIES_ELEC; life[(size_t)2 * e] = -1;
            sp[(size_t)2 * e + 1] = RB2_SPECIES_ION; life[(size_t)2 * e + 1] = step + S.ion_life_time;
            if (room > 0) { r.new_id = id++; --room; }   // newID = nrID before the call, :670
            if (room > 0) { r.ion_id = id++; --room; }
        }
        rc = rb2_add_particles(2 * ne, pos.data(), vel.data(), sp.data(), step, emit.data(), sec.data(), life.data());
        if (rc) return rc;
    }
    return RB2_OK;
}

}  // namespace

void rb2_collisions_release(Rb2Ctx &ctx)
{
    (void)ctx;
    CollState &S = g_coll;
    cudaFree(S.tab); cudaFree(S.erec); cudaFree(S.eidx); cudaFree(S.ions); cudaFree(S.cand); cudaFree(S.hits); cudaFree(S.ionev);
    cudaFree(S.d_cnt);
    if (S.h_cnt) cudaFreeHost(S.h_cnt);
    if (S.e0) cudaEventDestroy(S.e0);
    if (S.e1) cudaEventDestroy(S.e1);
    S = CollState{};
}

extern "C" {

int rb2_collisions_init(const rb2_collision_config *cfg)
{
    RB2_REQUIRE_INIT();
    if (!cfg) return rb2_fail(RB2_ERR_ARG, "config is NULL");
    if (g_rb2_ndev > 1) return rb2_fail(RB2_ERR_ARG, "the collision step keeps its state for one device: not available with rb2_set_devices");
    if (cfg->collision_mode != 1 && cfg->collision_mode != 2)
        return rb2_fail(RB2_ERR_ARG, "collision_mode %d: only 1 (continuous ionisation) and 2 (+ discrete recombination) run on "
                                     "the device", cfg->collision_mode);
    if (cfg->n_tot < 2 || cfg->n_ion < 2 || !cfg->tot_energy || !cfg->tot_data || !cfg->ion_energy || !cfg->ion_data)
        return rb2_fail(RB2_ERR_ARG, "cross-section tables with at least two rows each are required");
    if (!(cfg->n_d > 0.0)) return rb2_fail(RB2_ERR_ARG, "n_d must be positive");
    rb2_collisions_release(g_rb2);
    CollState &S = g_coll;
    S.mode = cfg->collision_mode; S.ion_life_time = cfg->ion_life_time; S.n_d = cfg->n_d; S.cyl_radius = cfg->cyl_radius;
    S.n_tot = cfg->n_tot; S.n_ion = cfg->n_ion;
    const size_t nt = (size_t)cfg->n_tot, ni = (size_t)cfg->n_ion;
    std::vector<double> tab(2 * nt + 2 * ni);
    std::copy(cfg->tot_energy, cfg->tot_energy + nt, tab.begin());
    std::copy(cfg->tot_data, cfg->tot_data + nt, tab.begin() + nt);
    std::copy(cfg->ion_energy, cfg->ion_energy + ni, tab.begin() + 2 * nt);
    std::copy(cfg->ion_data, cfg->ion_data + ni, tab.begin() + 2 * nt + ni);
    RB2_CUDA(cudaMalloc(&S.tab, tab.size() * sizeof(double)));
    RB2_CUDA(cudaMemcpy(S.tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    RB2_CUDA(cudaMalloc(&S.d_cnt, sizeof(CollCounts)));
    RB2_CUDA(cudaMallocHost(&S.h_cnt, sizeof(CollCounts)));
    RB2_CUDA(cudaEventCreate(&S.e0));
    RB2_CUDA(cudaEventCreate(&S.e1));
    S.ready = true;
    return RB2_OK;
}

int rb2_collision_data(double *out)
{
    int rc = require_ready();
    if (rc) return rc;
    Rb2Ctx &c = g_rb2;
    const int n = c.n;
    if (n < 1) return RB2_OK;
    if ((rc = rb2_ensure_stage(c, (size_t)5 * n, 0))) return rc;
    const CollParams P = make_params(c, 0, 0ull);
    k_coll_data<<<(n + 255) / 256, 256, 0, c.stream>>>(n, c.a.vel, c.a.species, P, c.d_stage_d);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(1);
    if (out) RB2_CUDA(cudaMemcpyAsync(out, c.d_stage_d, (size_t)5 * n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    RB2_CUDA(cudaStreamSynchronize(c.stream));
    return RB2_OK;
}

static int run_timed(int step, unsigned long long seed, int what, rb2_collision_result *out)
{
    int rc = require_ready();
    if (rc) return rc;
    Rb2Ctx &c = g_rb2;
    CollState &S = g_coll;
    S.last = rb2_collision_result{};
    S.host_ion.clear();
    S.host_recomb.clear();
    RB2_CUDA(cudaEventRecord(S.e0, c.stream));
    if (what & 1) { if ((rc = do_ionization(c, step, seed))) return rc; }
    if (what & 2) {
        if ((rc = do_recombination(c, step))) return rc;
        if (what & 1) S.last.nrCollisions += S.last.nrRecombinations;  // :45, :61
    }
    RB2_CUDA(cudaEventRecord(S.e1, c.stream));
    RB2_CUDA(cudaEventSynchronize(S.e1));
    RB2_CUDA(cudaEventElapsedTime(&S.last.ms, S.e0, S.e1));
    if ((rc = rb2_fetch_counters(c))) return rc;
    fill_result(c, out);
    return RB2_OK;
}

int rb2_continuous_ionization(int step, unsigned long long seed, rb2_collision_result *out)
{
    return run_timed(step, seed, 1, out);
}

int rb2_discrete_recombination(int step, rb2_collision_result *out)
{
    return run_timed(step, 0ull, 2, out);
}

int rb2_do_collisions(int step, unsigned long long seed, rb2_collision_result *out)
{
    int rc = require_ready();
    if (rc) return rc;
    return run_timed(step, seed, g_coll.mode == 2 ? 3 : 1, out);
}

int rb2_probe_quartic_roots(int n, const double *coeffs, int *codes_out, double *roots_out)
{
    RB2_REQUIRE_INIT();
    if (n < 1) return RB2_OK;
    if (!coeffs || !codes_out || !roots_out) return rb2_fail(RB2_ERR_ARG, "NULL argument");
    Rb2Ctx &c = g_rb2;
    int rc = rb2_ensure_stage(c, (size_t)13 * n, (size_t)n);
    if (rc) return rc;
    double *d_co = c.d_stage_d, *d_roots = d_co + (size_t)5 * n;
    RB2_CUDA(cudaMemcpyAsync(d_co, coeffs, (size_t)5 * n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    k_quartic_probe<<<(n + 127) / 128, 128, 0, c.stream>>>(n, d_co, c.d_stage_i, d_roots);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(1);
    RB2_CUDA(cudaMemcpyAsync(codes_out, c.d_stage_i, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    RB2_CUDA(cudaMemcpyAsync(roots_out, d_roots, (size_t)8 * n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    RB2_CUDA(cudaStreamSynchronize(c.stream));
    return RB2_OK;
}

int rb2_get_recombination_records(int max_records, rb2_recomb_record *out, int *n_out)
{
    RB2_REQUIRE_INIT();
    const int n = (int)g_coll.host_recomb.size();
    if (n_out) *n_out = n;
    if (out)
        for (int k = 0; k < n && k < max_records; ++k) out[k] = g_coll.host_recomb[(size_t)k];
    return RB2_OK;
}

int rb2_get_ionization_records(int max_records, rb2_ionization_record *out, int *n_out)
{
    RB2_REQUIRE_INIT();
    const int n = (int)g_coll.host_ion.size();
    if (n_out) *n_out = n;
    if (out)
        for (int k = 0; k < n && k < max_records; ++k) out[k] = g_coll.host_ion[(size_t)k];
    return RB2_OK;
}

}  // extern "C"

// rb2_mh.cu -- device-resident lock-step Metropolis-Hastings sampler for the planar emitters.
//
// Replaces the host loop of Metropolis_Hastings_rectangle_J_batch (reference
// src/mod_field_emission_v2.F90:1284-1458): M chains advance together.  ONE persistent cooperative kernel
// runs the search for favourable start spots and all jump iterations; per iteration
//   field sums   the sequence of (tile of 32 chains) x (128-particle sub-tile) work items is cut into one equal
//                contiguous range per CTA (even finish, no work queue): the proposals (Marsaglia polar normals,
//                reflection at the emitter edges :1466-1516) are recomputed from the counter-based generator wherever
//                they are needed, and the surface field sum runs over the particle records staged in shared memory;
//   accept       the CTA that delivers the LAST partial sum of a tile (per-tile arrival counter, "last block" pattern)
//                joins the tile's sums in a fixed order, adds the vacuum field and does the accept / reject step of
//                its 32 chains, one per lane, on the log electron supply (Elec_Supply_log :589, or ln J_GTF for the
//                thermal-field mode, src/mod_field_thermo_emission.F90:257);
// then ONE grid barrier (an arrival counter, two CTAs per SM), after which every thread applies the same MH_std update
// (:603-612, one per iteration after the warm-up) from the iteration's accept / reject counters.  (The first version
// had a cooperative-groups grid barrier after each of the two phases and up to four CTAs per SM: 20 us per jump at
// 324 chains x 8.7e3 electrons for 5 us of field sums.)  At most 512 chains: k_mh_small below.
//
// Surface field.  The chains live on the cathode plane z = 0, where the image series of
// src/acc_ic_planar_series.inc is mirror-antisymmetric: the partner of charge q at height h = z_j + 2nd is
// the charge -q at -h, so E_x = E_y = 0 and
//   E_z(p) = E_vac - 2 / (4 pi eps0) * sum_j q_j sum_{n=-N..N} h_jn / (rho^2 + h_jn^2)^(3/2)
// (one partner of each mirrored couple is evaluated and doubled; without image charges only n = 0, not
// doubled).  h_jn and q_j h_jn are precomputed once per call into 64-byte particle records, which leaves
// 31 FP64 instructions per (chain, particle) at N_ic_max = 1 instead of the 74 of the general field kernel.
// Random numbers: Philox4x32-10 keyed by (seed, chain), counter = (iteration, purpose, attempt): the
// reference's RANDOM_NUMBER is compiler specific, so parity is statistical (tests/test_emission.py).
#include "rb2_internal.cuh"
#include "rb2_tip_math.cuh"

#include <algorithm>

namespace {

constexpr int MHB = 128;
constexpr double TINY = 2.2250738585072014e-308;
constexpr double HUGE_NEG = -1.7976931348623157e308;

struct MhParams {
    rb2_mh_config c;
    const double *w_theta;  // [y_num][x_num] on the device
    double b_FN, l_const;
};

// ---- Philox4x32-10 ---------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox(uint4 ctr, uint2 key)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
__device__ __forceinline__ double u53(unsigned hi, unsigned lo)
{
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}
// two uniforms in [0,1) for (chain, iteration, purpose, attempt)
__device__ __forceinline__ void rand2(unsigned long long seed, int chain, int iter, int purpose, int attempt, double &u, double &v)
{
    const uint2 key = make_uint2((unsigned)seed ^ (unsigned)(chain * 0x9E3779B1u), (unsigned)(seed >> 32) + (unsigned)chain);
    const uint4 r = philox(make_uint4((unsigned)iter, (unsigned)purpose, (unsigned)attempt, (unsigned)chain), key);
    u = u53(r.x, r.y);
    v = u53(r.z, r.w);
}

// ---- work function + target densities ------------------------------------------------------------------
// w_theta_checkerboard, src/mod_work_function.F90:389-487
__device__ __forceinline__ double w_theta_xy(const MhParams &P, double x, double y)
{
    const rb2_mh_config &c = P.c;
    const double xs = (x - c.emit_pos[0]) / c.emit_dim[0], ys = (y - c.emit_pos[1]) / c.emit_dim[1];
    int x_i = (int)floor(xs / (1.0 / c.x_num)) + 1, y_i = (int)floor(ys / (1.0 / c.y_num)) + 1;
    x_i = min(max(x_i, 1), c.x_num);
    y_i = min(max(y_i, 1), c.y_num);
    y_i = c.y_num - y_i + 1;
    return P.w_theta[(y_i - 1) * c.x_num + (x_i - 1)];
}
// t_y / v_y, src/mod_field_emission_v2.F90:515-553
__device__ __forceinline__ double fn_l(const MhParams &P, double F, double w)
{
    double l = P.l_const * (-1.0 * F) / (w * w);
    return l > 1.0 ? 1.0 : l;
}
__device__ __forceinline__ double t_y(const MhParams &P, double F, double w)
{
    if (!P.c.image_charge) return 1.0;
    const double l = fn_l(P, F, w);
    return 1.0 + l * (1.0 / 9.0 - 1.0 / 18.0 * log(l));
}
__device__ __forceinline__ double v_y(const MhParams &P, double F, double w)
{
    if (!P.c.image_charge) return 1.0;
    const double l = fn_l(P, F, w);
    return 1.0 - l + 1.0 / 6.0 * l * log(l);
}
// Jensen GTF current density, src/mod_kevin_rjgtf_v2.f90:56-175
__device__ double Nns(double n, double s)
{
    if (n == 1.0) return (s + 1.0) * exp(-s);
    const double x = n * n, y = 1.0 / x, z = (n - 1.0) * s;
    double sng;
    if (fabs(z) > 1.0e-5) sng = (x + 1.0) * (x * exp(-s) - exp(-n * s)) / (x - 1.0);
    else sng = (0.5 * (x + 1.0) * exp(-s) / (n + 1.0)) * ((1.0 - n) * s * s + 2.0 * (1.0 + n) + 2.0 * s);
    const double sn = -x * (0.10593434 * x + 0.35506593), sd = -y * (0.10593434 * y + 0.35506593);
    return fmax(sng + sn * exp(-n * s) + x * sd * exp(-s), x * exp(-s));
}
__device__ double kevin_jgtf_v2(double F, double T, double Phi)
{
    const double kpi = 3.14159265358979324, kb = 1.0 / 11604.50635, hbar = 0.6582119571, c = 299.7924580;
    const double mo = 5.685630103, afs = 1.0 / 137.035999084, Qo = afs * hbar * c / 4.0, cm = 1.0e7, Amp = 6.241509074e3;
    const double Arld = (mo * (kb * kb) / (2.0 * (kpi * kpi) * (hbar * hbar * hbar))) * (cm * cm) / Amp;
    const double Fo = fabs(F) * 1.0e-9;
    if (Fo < 1.0e-9) return 0.0;
    const double yo = sqrt(4.0 * Qo * Fo) / Phi, phix = Phi - sqrt(4.0 * Qo * Fo);
    const double ty = 1.0 + (yo * yo) * (1.0 - log(yo)) / 9.0, vy = 1.0 - (yo * yo) * (3.0 - log(yo)) / 3.0;
    const double Tmin = (hbar * Fo / (4.0 * kb * ty)) * sqrt(2.0 / (mo * Phi)), Tmax = hbar * Fo / (kb * kpi * sqrt(mo * Phi * yo));
    const double betaT = 1.0 / (kb * T), betau = (2.0 / (hbar * Fo)) * sqrt(2.0 * mo * Phi) * ty;
    const double betap = (kpi / (hbar * Fo)) * sqrt(mo * Phi * yo), theto = (4.0 * sqrt(2.0 * mo * (Phi * Phi * Phi)) / (3.0 * hbar * Fo)) * vy;
    double nft, sft;
    if (T < Tmin) { nft = betaT / betau; sft = theto; }
    else if (T > Tmax) { nft = betaT / betap; sft = betap * phix; }
    else {
        const double Ap = 3.0 * (betap + betau) - 6.0 * theto / phix, Bp = -2.0 * (betap + 2.0 * betau) + 6.0 * theto / phix, Cp = betau - betaT;
        const double po = (-Bp - sqrt(Bp * Bp - 4.0 * Ap * Cp)) / (2.0 * Ap);
        const double theta = ((1.0 - po) * (1.0 - po)) * (2.0 * po + 1.0) * theto - phix * po * (1.0 - po) * ((1.0 - po) * betau - po * betap);
        nft = 1.0;
        sft = theta + betaT * (po * phix);
    }
    return (Arld * Nns(nft, sft) * (T * T)) * 1.0e4;
}
// log of the chain target at a favourable field F < 0, work function w at the spot
__device__ __forceinline__ double target_log_w(const MhParams &P, double F, double w)
{
    if (P.c.kind == 2) return log(fmax(kevin_jgtf_v2(F, P.c.T_temp, w), TINY));
    return 2.0 * log(-1.0 * F) - 2.0 * log(t_y(P, F, w)) - log(w);  // Elec_Supply_log
}
__device__ __forceinline__ double target_log(const MhParams &P, double F, double x, double y)
{
    return target_log_w(P, F, w_theta_xy(P, x, y));
}

struct __align__(16) SurfRec {
    double x, y, h0, g0, h1, g1, h2, g2;  // g_n = q * h_n;  N_ic_max >= 2: h0 = z, g1 = q, the rest on the fly
};

struct MhState {
    double *cur_x, *cur_y, *sup_cur, *F_cur;  // [M]
    int *ok;                                   // [M]
    double *partial;                           // [nsplit][M]
    int *cnt;                                  // [2 * (ndim + 1)] accepted / rejected per iteration
    int *bad;                                  // [max_init] chains still without a favourable spot per round
    const SurfRec *recs;                       // [n]
    const long long *wstart;                   // [G + 1] first work item of every CTA (host-computed W*g/G)
    const int *tfirst;                         // [G] tile of that work item
    const int *kfirst;                         // [G] how many CTAs contribute to that tile before this one
    const int *tcount;                         // [n_tiles] contributions per tile
    unsigned *tile_arrive;                     // [n_tiles] partial sums delivered so far (all phases, cumulative)
    unsigned *bar;                             // arrival counter of the grid barrier
    double *df_out, *F_out, *pos_out, *scal_out;
};

struct MhPlan {
    int M, n, n_tiles, nsplit, j_chunk, units, max_init;
    // persistent sampler: the sequence of (tile, 128-record sub-tile) work items, S per tile, W in total, is cut
    // into gridDim.x equal contiguous ranges; a CTA writes one partial sum per tile its range touches, into
    // partial[tile][k][32] with k = its rank among the tile's contributors (maxslots = most contributors)
    int S, maxslots;
    int resident;  // sub-tiles of records every CTA keeps in shared memory (0: stream them from L2 every iteration)
    long long W;
    unsigned long long seed;
    double two_d, E_vac, fac, mh_std0, a_rate0;
    int nic;
    int far;  // d >= 1 um and option sym_far: the n = 1 partners of a cathode-plane point skip the softening term
};

// Grid-wide barrier on an arrival counter in global memory (zeroed by the host): release add by one thread per CTA,
// acquire spin until `target` arrivals.  Needs all CTAs resident (cooperative launch).  Bounded: a CTA that never
// arrives traps the kernel instead of hanging the GPU.
__device__ __forceinline__ void small_barrier_arrive(unsigned *bar)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
    }
}
__device__ __forceinline__ void small_barrier_wait(unsigned *bar, unsigned target)
{
    if (threadIdx.x == 0) {
        unsigned v, spins = 0;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if (v >= target) break;
            if (++spins > (1u << 27)) __trap();  // seconds: a peer CTA is gone
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void small_barrier(unsigned *bar, unsigned target)
{
    small_barrier_arrive(bar);
    small_barrier_wait(bar, target);
}

// one particle against one surface point: sum of g_n / (rho^2 + h_n^2)^(3/2).  EXACT: reference sqrt / divide (slow
// path of laterally close pairs, see rb2_is_close); otherwise `close` collects the flag.
template <int NIC, bool EXACT, bool FAR = false>
__device__ __forceinline__ double surf_term(const SurfRec &r, double px, double py, double acc, const MhPlan &L, bool &close)
{
    const double dx = px - r.x, dy = py - r.y;
    const double d2 = fma(dy, dy, fma(dx, dx, RB2_S_FLOOR));
    if (!EXACT) close = close || rb2_is_close(d2);
    acc = fma(r.g0, rb2_inv_r3_sel<EXACT>(fma(r.h0, r.h0, d2)), acc);
    if (NIC == 1) {
        // FAR (MhPlan.far: d >= 1 um): both n = 1 partners are at least d from the cathode plane -- no softening term,
        // like the pair kernels' far partners (rb2_inv_r3_far)
        acc = fma(r.g1, (FAR && !EXACT) ? rb2_inv_r3_far(fma(r.h1, r.h1, d2)) : rb2_inv_r3_sel<EXACT>(fma(r.h1, r.h1, d2)), acc);
        acc = fma(r.g2, (FAR && !EXACT) ? rb2_inv_r3_far(fma(r.h2, r.h2, d2)) : rb2_inv_r3_sel<EXACT>(fma(r.h2, r.h2, d2)), acc);
    } else if (NIC >= 2) {
        for (int n = 1; n <= L.nic; ++n) {
            const double h = L.two_d * (double)n, hm = r.h0 - h, hp = r.h0 + h;
            acc = fma(r.g1 * hm, rb2_inv_r3_sel<EXACT>(fma(hm, hm, d2)), acc);
            acc = fma(r.g1 * hp, rb2_inv_r3_sel<EXACT>(fma(hp, hp, d2)), acc);
        }
    }
    return acc;
}
// CNT consecutive records against this lane's point, added to acc as one block sum
template <int NIC, int CNT>
__device__ __forceinline__ double surf_block(const SurfRec *rr, double px, double py, double acc, const MhPlan &L)
{
    double part = 0.0;
    bool close = false;
    if (NIC == 1 && L.far) {
#pragma unroll 4
        for (int k = 0; k < CNT; ++k) part = surf_term<NIC, false, true>(rr[k], px, py, part, L, close);
    } else {
#pragma unroll 4
        for (int k = 0; k < CNT; ++k) part = surf_term<NIC, false>(rr[k], px, py, part, L, close);
    }
    if (close) {
        part = 0.0;
        for (int k = 0; k < CNT; ++k) part = surf_term<NIC, true>(rr[k], px, py, part, L, close);
    }
    return acc + part;
}

// the same for a run-time number of records (the partly filled last sub-tile of a CTA of k_mh_small)
template <int NIC>
__device__ __forceinline__ double surf_block_n(const SurfRec *rr, int cnt, double px, double py, double acc, const MhPlan &L)
{
    double part = 0.0;
    bool close = false;
    if (NIC == 1 && L.far) {
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) part = surf_term<NIC, false, true>(rr[k], px, py, part, L, close);
    } else {
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) part = surf_term<NIC, false>(rr[k], px, py, part, L, close);
    }
    if (close) {
        part = 0.0;
        for (int k = 0; k < cnt; ++k) part = surf_term<NIC, true>(rr[k], px, py, part, L, close);
    }
    return acc + part;
}

__global__ void k_surf_pack(const double4 *__restrict__ pq, int n, double two_d, int nic, SurfRec *__restrict__ recs)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double4 p = pq[j];
    SurfRec r;
    r.x = p.x; r.y = p.y; r.h0 = p.z; r.g0 = p.w * p.z;
    if (nic >= 2) { r.h1 = 0.0; r.g1 = p.w; r.h2 = 0.0; r.g2 = 0.0; }
    else { r.h1 = p.z - two_d; r.g1 = p.w * r.h1; r.h2 = p.z + two_d; r.g2 = p.w * r.h2; }
    recs[j] = r;
}

// Proposal of chain c at iteration iter: a pure function of (seed, chain, iteration, chain state, step), so any
// thread that needs it recomputes it.  iter < 0: search rounds for a favourable start (uniform over the emitter).
// The two standard normals of chain k's jump `iter`: a function of (seed, chain, iteration) only, so they can be drawn
// ahead of time (k_mh_small draws the next jump's while it waits at the barrier).
__device__ __forceinline__ void draw_jump_normals(const MhPlan &L, int iter, int k, double &g0, double &g1)
{
    g0 = 0.0; g1 = 0.0;
    // Marsaglia polar method, src/mod_global.F90:578-595.  The lanes of a warp leave the rejection loop at different
    // attempts (78 % per attempt: ~3 rounds until all 32 are through), so only the cheap part is inside it; the
    // log / divide / sqrt of the accepted point runs once, behind the loop (same values: same operations).
    double a = 0.0, b = 0.0, w = 0.0;
    bool found = false;
    for (int attempt = 0; attempt < 64 && !found; ++attempt) {
        double u, v;
        rand2(L.seed, k, iter, 1, attempt, u, v);
        a = 2.0 * u - 1.0; b = 2.0 * v - 1.0; w = a * a + b * b;
        found = (w < 1.0 && w > 0.0);
    }
    if (found) {
        const double f = sqrt((-2.0 * log(w)) / w);
        g0 = a * f; g1 = b * f;
    }
}
__device__ __forceinline__ void propose_apply(const MhParams &P, int iter, double mh_std, double g0, double g1, double &x, double &y);
// (x, y) enter as the chain's current position and leave as the proposal.
__device__ __forceinline__ void propose_from(const MhParams &P, const MhPlan &L, int iter, int k, double mh_std, int ok,
                                             double &x, double &y)
{
    const rb2_mh_config &c = P.c;
    if (iter < 0) {
        if (!ok) {
            double u, v;
            rand2(L.seed, k, iter, 0, 0, u, v);
            x = u * c.emit_dim[0] + c.emit_pos[0];
            y = v * c.emit_dim[1] + c.emit_pos[1];
        }
        return;
    }
    if (!ok) return;
    double g0, g1;
    draw_jump_normals(L, iter, k, g0, g1);
    propose_apply(P, iter, mh_std, g0, g1, x, y);
}
__device__ __forceinline__ void propose_apply(const MhParams &P, int iter, double mh_std, double g0, double g1, double &x, double &y)
{
    const rb2_mh_config &c = P.c;
    const double frac = (iter > c.ndim_first) ? mh_std : c.init_std;
    x += g0 * (c.emit_dim[0] * frac);
    y += g1 * (c.emit_dim[1] * frac);
    if (c.kind == 2) {  // src/mod_field_thermo_emission.F90:369-389
        double qx = (x - c.emit_pos[0]) / c.emit_dim[0], qy = (y - c.emit_pos[1]) / c.emit_dim[1];
        if (qx > 1.0 || qx < 0.0) qx = 1.0 - (qx - floor(qx));
        if (qy > 1.0 || qy < 0.0) qy = 1.0 - (qy - floor(qy));
        x = qx * c.emit_dim[0] + c.emit_pos[0];
        y = qy * c.emit_dim[1] + c.emit_pos[1];
    } else {  // src/mod_field_emission_v2.F90:1466-1516
        const double x_max = c.emit_pos[0] + c.emit_dim[0], x_min = c.emit_pos[0];
        const double y_max = c.emit_pos[1] + c.emit_dim[1], y_min = c.emit_pos[1];
        if (x > x_max) x = x_max - (x - x_max); else if (x < x_min) x = (x_min - x) + x_min;
        if (y > y_max) y = y_max - (y - y_max); else if (y < y_min) y = (y_min - y) + y_min;
    }
}
__device__ __forceinline__ void propose(const MhParams &P, const MhState &S, const MhPlan &L, int iter, int k, double mh_std,
                                        double &x, double &y)
{
    x = __ldcg(&S.cur_x[k]);
    y = __ldcg(&S.cur_y[k]);
    propose_from(P, L, iter, k, mh_std, __ldcg(&S.ok[k]), x, y);
}

// One work unit: 32 surface points (one per lane, the same in all four warps) against the particle records
// [j0, j1).  Of every 128 records warp w stages records 32w .. 32w+31 in its own slice of shared memory
// (double buffered, register prefetch) and reads them back as broadcasts; the four warp sums are joined in a
// fixed order.  The result is valid in warp 0.
template <int NIC>
__device__ __forceinline__ double surf_unit_sum(const SurfRec *__restrict__ g_recs, int j0, int j1, double px, double py,
                                                const MhPlan &L, SurfRec (*recs)[MHB], double (*red)[32])
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nsub = (j1 - j0 + MHB - 1) / MHB;
    SurfRec nxt;
    auto fetch = [&](int t) {
        const int j = j0 + t * MHB + tid;
        if (j < j1) nxt = g_recs[j];
        else { nxt.x = 1.0; nxt.y = 1.0; nxt.h0 = 1.0; nxt.g0 = 0.0; nxt.h1 = 1.0; nxt.g1 = 0.0; nxt.h2 = 1.0; nxt.g2 = 0.0; }
    };
    fetch(0);
    double acc = 0.0;
    // each warp stages and consumes its own 32 records: only warp-level synchronisation inside the loop, so
    // the four warps of a CTA drift apart freely (a CTA barrier per sub-tile cost ~25 % in stalls)
    for (int t = 0; t < nsub; ++t) {
        recs[t & 1][tid] = nxt;
        __syncwarp();
        if (t + 1 < nsub) fetch(t + 1);
        const SurfRec *rr = &recs[t & 1][warp * 32];
        acc = surf_block<NIC, 32>(rr, px, py, acc, L);
        __syncwarp();
    }
    red[warp][lane] = acc;
    __syncthreads();
    const double sum = ((red[0][lane] + red[1][lane]) + red[2][lane]) + red[3][lane];
    __syncthreads();
    return sum;
}

// Accept / reject step of one tile of 32 chains (lane = chain), run by the warp that delivered the tile's last partial
// sum (joined by the caller): finish the field, decide, count.  (x, y) is the lane's proposal.
__device__ __forceinline__ void tile_accept(const MhParams &P, const MhState &S, const MhPlan &L, int iter, int tile, double x,
                                            double y, double sum)
{
    const int lane = threadIdx.x & 31;
    const int c = tile * 32 + lane;
    bool acc = false, rej = false, bad = false;
    if (c < L.M) {
        const double Fz = L.E_vac - L.fac * sum;
        const int ok = __ldcg(&S.ok[c]);
        if (iter < 0) {
            if (!ok) {
                if (Fz < 0.0) {
                    S.cur_x[c] = x; S.cur_y[c] = y; S.F_cur[c] = Fz;
                    S.sup_cur[c] = target_log(P, Fz, x, y);
                    S.ok[c] = 1;
                } else bad = true;
            }
        } else if (ok) {
            const bool unfav = (P.c.kind == 2) ? (Fz > 0.0) : (Fz >= 0.0);
            bool accept = false;
            if (!unfav) {
                const double sup_new = target_log(P, Fz, x, y), sup_old = __ldcg(&S.sup_cur[c]);
                accept = sup_new >= sup_old;
                if (!accept) {
                    double u, v;
                    rand2(L.seed, c, iter, 2, 0, u, v);
                    accept = log(u) <= sup_new - sup_old;
                }
                if (accept) { S.cur_x[c] = x; S.cur_y[c] = y; S.sup_cur[c] = sup_new; S.F_cur[c] = Fz; }
            }
            acc = accept; rej = !accept;
        }
    }
    const int n_acc = __popc(__ballot_sync(0xffffffffu, acc)), n_rej = __popc(__ballot_sync(0xffffffffu, rej));
    const int n_bad = __popc(__ballot_sync(0xffffffffu, bad));
    if (lane == 0) {
        if (iter < 0) { if (n_bad) atomicAdd(&S.bad[-iter - 1], n_bad); }
        else {
            if (n_acc) atomicAdd(&S.cnt[2 * iter], n_acc);
            if (n_rej) atomicAdd(&S.cnt[2 * iter + 1], n_rej);
        }
    }
}

// The same unit sum over `nsub` sub-tiles that this CTA keeps resident in shared memory (see k_mh_persistent).
template <int NIC>
__device__ __forceinline__ double surf_unit_sum_resident(const SurfRec *__restrict__ res, int nsub, double px, double py,
                                                         const MhPlan &L, double (*red)[32])
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc = 0.0;
    for (int t = 0; t < nsub; ++t) {
        const SurfRec *rr = res + (size_t)t * MHB + warp * 32;
        acc = surf_block<NIC, 32>(rr, px, py, acc, L);
    }
    red[warp][lane] = acc;
    __syncthreads();
    const double sum = ((red[0][lane] + red[1][lane]) + red[2][lane]) + red[3][lane];
    __syncthreads();
    return sum;
}

// Field sums of this CTA's range of (chain tile, 128-record sub-tile) work items; the CTA that delivers the last
// partial sum of a tile also runs that tile's accept / reject step.
template <int NIC>
__device__ __forceinline__ void phase_field(const MhParams &P, const MhState &S, const MhPlan &L, int iter, double mh_std,
                                            unsigned phase, const SurfRec *resident, SurfRec (*recs)[MHB], double (*red)[32],
                                            int *s_last)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long w0 = S.wstart[blockIdx.x], w1 = S.wstart[blockIdx.x + 1];
    const int tile_first = S.tfirst[blockIdx.x];
    for (long long w = w0; w < w1;) {
        const int tile = (int)(w / L.S), s0 = (int)(w - (long long)tile * L.S);
        const int s1 = (int)min((long long)L.S, (long long)s0 + (w1 - w));
        const int c = tile * 32 + lane;
        double px = 0.0, py = 0.0;
        if (c < L.M) propose(P, S, L, iter, c, mh_std, px, py);
        const double sum = resident ? surf_unit_sum_resident<NIC>(resident + (size_t)(w - w0) * MHB, s1 - s0, px, py, L, red)
                                    : surf_unit_sum<NIC>(S.recs, s0 * MHB, min(L.n, s1 * MHB), px, py, L, recs, red);
        // contribution index within the tile: only the first tile of a range can have earlier contributors
        const int k = (tile == tile_first) ? S.kfirst[blockIdx.x] : 0;
        if (warp == 0) {
            S.partial[((size_t)tile * L.maxslots + k) * 32 + lane] = sum;
            // "last block" pattern (fence, count, fence): whoever delivers the LAST partial sum of a tile does that tile's
            // accept / reject step right away -- no grid barrier between the two phases, and the accept work of early
            // tiles overlaps the field sums of late ones
            __threadfence();
            if (lane == 0) *s_last = (atomicAdd(&S.tile_arrive[tile], 1u) + 1u == (unsigned)S.tcount[tile] * (phase + 1u));
        }
        __syncthreads();
        if (*s_last) {
            // join in a fixed order with all four warps: warp w adds contributions w, w + 4, ... (ascending), then the
            // four strands in warp order.  (One warp walking all contributions paid the L2 latency ~nk / 8 times.)
            __threadfence();
            const int nk = S.tcount[tile];
            const double *pp = S.partial + ((size_t)tile * L.maxslots) * 32 + lane;
            double strand = 0.0;
#pragma unroll 16
            for (int q = warp; q < nk; q += MHB / 32) strand += __ldcg(pp + (size_t)q * 32);
            red[warp][lane] = strand;
            __syncthreads();
            const double joined = ((red[0][lane] + red[1][lane]) + red[2][lane]) + red[3][lane];
            __syncthreads();
            if (warp == 0) tile_accept(P, S, L, iter, tile, px, py, joined);
        }
        w += s1 - s0;
    }
}

// rb2_field_surface_z: the same unit sums for M caller-supplied surface points, then a warp per point joins them
template <int NIC>
__global__ void __launch_bounds__(MHB, 4) k_surface_field(const double *__restrict__ pts, const SurfRec *__restrict__ g_recs,
                                                          MhPlan L, double *__restrict__ partial)
{
    __shared__ SurfRec recs[2][MHB];
    __shared__ double red[MHB / 32][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int u = blockIdx.x; u < L.units; u += gridDim.x) {
        const int tile = u / L.nsplit, s = u - tile * L.nsplit;
        const int c = tile * 32 + lane;
        double px = 0.0, py = 0.0;
        if (c < L.M) { px = pts[3 * c]; py = pts[3 * c + 1]; }
        const int j0 = s * L.j_chunk, j1 = min(L.n, j0 + L.j_chunk);
        const double sum = surf_unit_sum<NIC>(g_recs, j0, j1, px, py, L, recs, red);
        if (warp == 0 && c < L.M) partial[(size_t)s * L.M + c] = sum;
    }
}
__global__ void k_surface_join(const double *__restrict__ partial, MhPlan L, double *__restrict__ Ez)
{
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= L.M) return;
    double sum = 0.0;
    for (int s = lane; s < L.nsplit; s += 32) sum += partial[(size_t)s * L.M + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) Ez[c] = L.E_vac - L.fac * sum;
}

template <int NIC>
__global__ void __launch_bounds__(MHB, 4) k_mh_persistent(MhParams P, MhState S, MhPlan L)
{
    __shared__ SurfRec recs[2][MHB];
    __shared__ double red[MHB / 32][32];
    __shared__ int s_last;
    extern __shared__ __align__(16) unsigned char mh_dyn_smem[];
    // The work range of a CTA is the same in every iteration and the particle records do not change while the chains
    // run: when the range is short (L.resident sub-tiles fit every range) the CTA loads its records ONCE and keeps them
    // in shared memory -- otherwise every unit starts with an exposed L2 round trip (~1 us for ~1 us of arithmetic).
    const SurfRec *resident = nullptr;
    if (L.resident > 0) {
        SurfRec *res = reinterpret_cast<SurfRec *>(mh_dyn_smem);
        const long long w0 = S.wstart[blockIdx.x], w1 = S.wstart[blockIdx.x + 1];
        for (long long w = w0; w < w1; ++w) {
            const int sub = (int)(w % L.S), j = sub * MHB + threadIdx.x;
            SurfRec r;
            if (j < L.n) r = S.recs[j];
            else { r.x = 1.0; r.y = 1.0; r.h0 = 1.0; r.g0 = 0.0; r.h1 = 1.0; r.g1 = 0.0; r.h2 = 1.0; r.g2 = 0.0; }
            res[(size_t)(w - w0) * MHB + threadIdx.x] = r;
        }
        __syncthreads();
        resident = res;
    }
    double mh_std = L.mh_std0, a_rate = L.a_rate0;
    unsigned phase = 0;  // field phases so far (the same in every CTA)
    // a favourable start for every chain (:1303-1361): rounds of uniform draws until the field is negative
    int bad = L.M;
    for (int r = 0; r < L.max_init && bad > 0; ++r) {
        phase_field<NIC>(P, S, L, -(r + 1), mh_std, phase, resident, recs, red, &s_last);
        ++phase;
        small_barrier(S.bar, phase * gridDim.x);
        bad = __ldcg(&S.bad[r]);
    }
    for (int i = 1; i <= P.c.ndim; ++i) {
        phase_field<NIC>(P, S, L, i, mh_std, phase, resident, recs, red, &s_last);
        ++phase;
        small_barrier(S.bar, phase * gridDim.x);
        const int a = __ldcg(&S.cnt[2 * i]), r = __ldcg(&S.cnt[2 * i + 1]);
        if (i > P.c.ndim_first && a + r > 0) {  // MH_std_update, :603-612 -- identical in every thread
            a_rate = (double)a / (double)(a + r);
            mh_std = fmin(fmax(mh_std * exp(P.c.std_gain * (a_rate - P.c.target_rate)), P.c.std_min), P.c.std_max);
        }
    }
    // outputs: position, surface field and the log escape probability (Escape_Prob_log :568) of every chain
    for (int k = blockIdx.x * MHB + threadIdx.x; k < L.M; k += gridDim.x * MHB) {
        if (__ldcg(&S.ok[k])) {
            const double x = __ldcg(&S.cur_x[k]), y = __ldcg(&S.cur_y[k]), F = __ldcg(&S.F_cur[k]);
            const double w = w_theta_xy(P, x, y), sw = sqrt(w);
            S.pos_out[3 * k] = x; S.pos_out[3 * k + 1] = y; S.pos_out[3 * k + 2] = 0.0;
            S.F_out[k] = F;
            S.df_out[k] = (P.c.kind == 2) ? 0.0 : P.b_FN * (sw * sw * sw) * v_y(P, F, w) / (-1.0 * F);
        } else {  // failed chain: defined outputs, no emission (:1346-1361)
            S.pos_out[3 * k] = P.c.emit_pos[0]; S.pos_out[3 * k + 1] = P.c.emit_pos[1]; S.pos_out[3 * k + 2] = 0.0;
            S.F_out[k] = 1.0;
            S.df_out[k] = HUGE_NEG;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { S.scal_out[0] = mh_std; S.scal_out[1] = a_rate; }
}

// ---- single-barrier variant for at most 512 chains ---------------------------------------------------------------
// In the reference's own regime (1e3 - 1e4 electrons in the gap, ~100 emission candidates per step) an iteration of
// k_mh_persistent is a chain of global-memory round trips (chain state, partial sums, arrival counters, accept
// counters, barrier: ~8 us even with no particles at all) around ~1 us of arithmetic.  With at most sixteen tiles
// of chains (one warp per tile) all of that state fits one CTA:
//   * CTA b works for ONE tile t_b = b / Gs on its own share of the particle records, which stays RESIDENT in shared
//     memory for the whole call (the records do not change while the chains run);
//   * every CTA keeps the state of ALL chains in registers: warp w holds tile w (lane = chain).  Per iteration warp t_b
//     computes the tile's 32 proposals and hands them to the other warps through shared memory, the warps sum the
//     resident records (128 / WPB each per sub-tile), and the CTA publishes one partial sum per chain of its tile;
//   * ONE barrier (arrival counter in global memory);
//   * behind it EVERY CTA joins the partial sums of every tile (all warps, fixed order) and warp w does the
//     accept / reject step of tile w -- redundantly: same inputs, same instructions, same result in every CTA, so no
//     chain state, no accept counters and no second barrier in global memory.  The MH_std update uses the counts of
//     the warps, exchanged through shared memory.  The partial sums are double buffered by iteration
//     parity (a CTA can only be one barrier ahead of the slowest one).
// Proposals, targets and the generator keys are those of k_mh_persistent (propose_from, target_log, rand2).
struct MhSmall {
    int T, Gs, G, S, R, U;  // tiles, CTAs per tile, CTAs = T * Gs, 128-record sub-tiles in total, most sub-tiles per CTA, 16-record units in total
    double *partial;        // [2][G][32]
    unsigned *bar;          // arrival counter, zeroed by the host
};

// WPB warps per CTA: 4 (up to 4 tiles = 128 chains) or 16 (up to 16 tiles = 512 chains, one CTA per SM).  A sub-tile is
// always MHB = 128 records; warp w sums records w * (128 / WPB) ... of every resident sub-tile.
template <int NIC, int WPB>
__global__ void __launch_bounds__(WPB * 32) k_mh_small(MhParams P, MhState S, MhPlan L, MhSmall Q)
{
    constexpr int NT = WPB * 32, RPW = MHB / WPB;
    extern __shared__ __align__(16) unsigned char small_smem[];
    // dynamic shared memory: resident sub-tiles (zero-weight padding beyond the particle list), then the join strands
    SurfRec *mine = reinterpret_cast<SurfRec *>(small_smem);
    double *strands = reinterpret_cast<double *>(small_smem + (size_t)Q.R * MHB * sizeof(SurfRec));  // [T][WPB][32]
    __shared__ double red[WPB][32];
    __shared__ double sp_x[32], sp_y[32];
    __shared__ int s_acc[WPB], s_rej[WPB], s_bad[WPB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t_b = blockIdx.x / Q.Gs, g_b = blockIdx.x - t_b * Q.Gs;
    // this CTA's records: an even share in units of 16 records (whole 128-record sub-tiles left 6 sub-tiles to some CTAs
    // and 5 to others at the deck's size, and every jump waited for the long ones: 7.2 vs 6.0 us of field sums)
    const int u0 = (int)((long long)Q.U * g_b / Q.Gs), u1 = (int)((long long)Q.U * (g_b + 1) / Q.Gs);
    const int cnt = (u1 - u0) * 16, nfull = cnt / MHB, rem = cnt - nfull * MHB, rem_w = rem / WPB;  // rem: a multiple of 16
    for (int q = tid; q < (nfull + (rem ? 1 : 0)) * MHB; q += NT) {
        const int j = u0 * 16 + q;
        SurfRec r;
        if (q < cnt && j < L.n) r = S.recs[j];
        else { r.x = 1.0; r.y = 1.0; r.h0 = 1.0; r.g0 = 0.0; r.h1 = 1.0; r.g1 = 0.0; r.h2 = 1.0; r.g2 = 0.0; }
        mine[q] = r;
    }
    if (tid < WPB) { s_acc[tid] = 0; s_rej[tid] = 0; s_bad[tid] = 0; }
    __syncthreads();
    // chain state of tile `warp` (meaningful for warp < T): lane = chain, identical in every CTA
    const int chain = warp * 32 + lane;
    const bool live = (warp < Q.T) && (chain < L.M);
    double cx = 0.0, cy = 0.0, sup = 0.0, Fc = 0.0, mh_std = L.mh_std0, a_rate = L.a_rate0;
    int ok = 0;
    unsigned phase = 0;
    int bad = L.M, round = 0, jump = 1;
    bool searching = true;
    double ahead_g0 = 0.0, ahead_g1 = 0.0;
    int ahead_iter = 0;  // jump whose normals are in ahead_g0 / ahead_g1
    // -DRB2_MH_TIMING: clock64 per phase, printed by the first and the last CTA (profiles/mh_small_phases_r02.log)
#ifdef RB2_MH_TIMING
    long long tk[6] = {0, 0, 0, 0, 0, 0}, tc = clock64();
#define RB2_TK(i) do { const long long now_ = clock64(); tk[i] += now_ - tc; tc = now_; } while (0)
#else
#define RB2_TK(i) do { } while (0)
#endif
    for (;;) {
        // rounds of the search for a favourable start (generator iteration -(round + 1), like k_mh_persistent), then the
        // jump iterations 1 .. ndim
        if (searching && !(round < L.max_init && bad > 0)) searching = false;
        if (!searching && jump > P.c.ndim) break;
        const int iter = searching ? -(round + 1) : jump;
        // proposals of every tile by the warp that holds it (kept: the accept step below needs them); the tile this CTA
        // sums for goes through shared memory to the other warps
        double qx = cx, qy = cy;
        if (live) {
            if (iter >= 1 && ok && ahead_iter == iter) propose_apply(P, iter, mh_std, ahead_g0, ahead_g1, qx, qy);
            else propose_from(P, L, iter, chain, mh_std, ok, qx, qy);
        }
        if (warp == t_b) { sp_x[lane] = qx; sp_y[lane] = qy; }
        __syncthreads();
        RB2_TK(0);  // propose
        const double px = sp_x[lane], py = sp_y[lane];
        double acc = 0.0;
        for (int t = 0; t < nfull; ++t) {
            const SurfRec *rr = &mine[t * MHB + warp * RPW];
            acc = surf_block<NIC, RPW>(rr, px, py, acc, L);
        }
        if (rem_w > 0) acc = surf_block_n<NIC>(&mine[nfull * MHB + warp * rem_w], rem_w, px, py, acc, L);
        red[warp][lane] = acc;
        __syncthreads();
        RB2_TK(1);  // field sums
        double *part = Q.partial + (size_t)(phase & 1u) * Q.G * 32;
        // The LAST warp joins the CTA's sums, stores them and arrives at the barrier (release fence + counter): it owns no
        // tile unless there are 16 of them, so the fence (~1 us until the stores are visible) no longer sits in front of
        // the look-ahead of the warp that does (clock64 profile: arrive + look-ahead 2.3 us of a 13 us jump, serial in warp 0)
        if (warp == WPB - 1) {
            double sum = red[0][lane];
#pragma unroll
            for (int w = 1; w < WPB; ++w) sum += red[w][lane];
            part[(size_t)blockIdx.x * 32 + lane] = sum;
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                atomicAdd(Q.bar, 1u);
            }
        }
        ++phase;
        // while the other CTAs arrive: the normals of the NEXT jump (Philox + log + sqrt, ~1 us of dependent latency that
        // would otherwise open the next iteration)
        // ... and what the accept step needs besides the field sum: the work function at the proposal, log(u) of the test
        double w_q = 0.0, log_u = 0.0;
        if (live) {
            w_q = w_theta_xy(P, qx, qy);
            if (iter >= 1 && ok) {
                double u, v;
                rand2(L.seed, chain, iter, 2, 0, u, v);
                log_u = log(u);
            }
        }
        if (live && !searching && jump < P.c.ndim) { ahead_iter = jump + 1; draw_jump_normals(L, ahead_iter, chain, ahead_g0, ahead_g1); }
        RB2_TK(2);  // store, arrive, look-ahead
        small_barrier_wait(Q.bar, phase * (unsigned)Q.G);
        RB2_TK(3);  // wait
        // join: for every tile, warp w adds the partial sums of that tile's CTAs w, w + WPB, ... (ascending); the WPB
        // strands are added in warp order by the warp that owns the tile.  The loads of four tiles are issued together:
        // tile after tile, each join paid its own L2 round trip.
        {
            const double *pt = part + lane;
            const int T = Q.T, Gs = Q.Gs;
            for (int t0 = 0; t0 < T; t0 += 4) {
                double st[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
                for (int k = warp; k < Gs; k += WPB) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (t0 + u < T) st[u] += __ldcg(pt + ((size_t)(t0 + u) * Gs + k) * 32);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (t0 + u < T) strands[((size_t)(t0 + u) * WPB + warp) * 32 + lane] = st[u];
            }
        }
        __syncthreads();
        RB2_TK(4);  // join loads
        bool acc_ = false, rej_ = false, bad_ = false;
        if (warp < Q.T) {
            const double *sw = strands + (size_t)warp * WPB * 32 + lane;
            double sum = sw[0];
#pragma unroll
            for (int w = 1; w < WPB; ++w) sum += sw[(size_t)w * 32];
            const double Fz = L.E_vac - L.fac * sum;
            if (live) {
                if (iter < 0) {
                    if (!ok) {
                        if (Fz < 0.0) { cx = qx; cy = qy; Fc = Fz; sup = target_log_w(P, Fz, w_q); ok = 1; }
                        else bad_ = true;
                    }
                } else if (ok) {
                    const bool unfav = (P.c.kind == 2) ? (Fz > 0.0) : (Fz >= 0.0);
                    bool accept = false;
                    if (!unfav) {
                        const double sup_new = target_log_w(P, Fz, w_q);
                        accept = (sup_new >= sup) || (log_u <= sup_new - sup);
                        if (accept) { cx = qx; cy = qy; sup = sup_new; Fc = Fz; }
                    }
                    acc_ = accept; rej_ = !accept;
                }
            }
            const int na = __popc(__ballot_sync(0xffffffffu, acc_)), nr = __popc(__ballot_sync(0xffffffffu, rej_));
            const int nb = __popc(__ballot_sync(0xffffffffu, bad_));
            if (lane == 0) { s_acc[warp] = na; s_rej[warp] = nr; s_bad[warp] = nb; }
        }
        __syncthreads();
        {   // every thread: the same counts, the same update (the next write to s_* sits behind two more barriers)
            int a = 0, r = 0, nb = 0;
#pragma unroll
            for (int w = 0; w < WPB; ++w) { a += s_acc[w]; r += s_rej[w]; nb += s_bad[w]; }
            if (iter > P.c.ndim_first && a + r > 0) {  // MH_std_update, :603-612
                a_rate = (double)a / (double)(a + r);
                mh_std = fmin(fmax(mh_std * exp(P.c.std_gain * (a_rate - P.c.target_rate)), P.c.std_min), P.c.std_max);
            }
            if (searching) { bad = nb; ++round; } else ++jump;
        }
        RB2_TK(5);  // accept + bookkeeping
    }
#ifdef RB2_MH_TIMING
    if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))
        printf("k_mh_small cta %d (tile %d, %d records): clocks propose %lld field %lld arrive+ahead %lld wait %lld join %lld accept %lld\n",
               (int)blockIdx.x, t_b, cnt, tk[0], tk[1], tk[2], tk[3], tk[4], tk[5]);
#endif
    if (blockIdx.x == 0) {
        if (live) {
            const int k = chain;
            if (ok) {
                const double w = w_theta_xy(P, cx, cy), sw = sqrt(w);
                S.pos_out[3 * k] = cx; S.pos_out[3 * k + 1] = cy; S.pos_out[3 * k + 2] = 0.0;
                S.F_out[k] = Fc;
                S.df_out[k] = (P.c.kind == 2) ? 0.0 : P.b_FN * (sw * sw * sw) * v_y(P, Fc, w) / (-1.0 * Fc);
            } else {
                S.pos_out[3 * k] = P.c.emit_pos[0]; S.pos_out[3 * k + 1] = P.c.emit_pos[1]; S.pos_out[3 * k + 2] = 0.0;
                S.F_out[k] = 1.0;
                S.df_out[k] = HUGE_NEG;
            }
        }
        if (tid == 0) { S.scal_out[0] = mh_std; S.scal_out[1] = a_rate; }
    }
}

// ---- the reference's DEFAULT sampler: serial chains (mh_batch = .false.) ---------------------------------------------
// Metropolis_Hastings_rectangle_J (src/mod_field_emission_v2.F90:1122-1265) is called once per emission candidate from
// the insert loop of Do_Field_Emission_Planar_rectangle (:322-380): chain s runs on the field of the store PLUS the
// electrons that the chains before it emitted in this same time step, and the shared step MH_std is adapted once per
// chain from that chain's own acceptance rate.  Both couplings are strict -- chain s needs the complete outcome of
// chain s - 1 -- so the N_round x (1 + 200) field evaluations of a time step are sequential by construction.  The host
// loop paid one M = 1 round trip (~40 us) per evaluation; here ONE CTA of 1024 threads runs the whole step: per jump the
// threads sum the particle records (resident in shared memory as far as they fit, the rest from L2) and the pending
// electrons of this step for the single proposal, a fixed-shape tree joins them, and warp 0 does the accept step, the
// chain bookkeeping, the emission test (ln u <= D_f, :339-379) and the next proposal: two CTA barriers per jump, no
// global memory traffic, ~2.5 us per evaluation.  kind 2: the chains of src/mod_field_thermo_emission.F90:198-364 (25
// jumps, every chain that found a start emits).  Generator keys (seed, chain, iteration) as in the lock-step kernels.
struct MhSerial {
    int M, n, resident;           // chains, particle records, how many of them live in shared memory
    double mh_std0, a_rate0;
    const SurfRec *recs;          // [n]
    double *df_out, *F_out, *pos_out, *scal_out;
    int *emit_out;                // [M] 1: the candidate was emitted
};
constexpr int SER_T = 1024, SER_PMAX = 1024;  // threads; most electrons one time step can emit through this path

template <int NIC>
__global__ void __launch_bounds__(SER_T, 1) k_mh_serial(MhParams P, MhPlan L, MhSerial Q)
{
    extern __shared__ __align__(16) unsigned char ser_smem[];
    SurfRec *res = reinterpret_cast<SurfRec *>(ser_smem);            // [resident]
    SurfRec *pend = res + Q.resident;                                // [SER_PMAX] electrons emitted in this step
    __shared__ double red[SER_T / 32];
    __shared__ double sp_x, sp_y;
    __shared__ int s_npend, s_done;
    // look-ahead of warp 1 (while warp 0 does the accept step): the random numbers of the three iterations that can
    // follow the current one -- the next jump of this chain, the next search round, the first search round of the next chain
    __shared__ double la_g0, la_g1, la_lu, la_nu, la_nv, la_ru, la_rv;
    __shared__ int st_s, st_ok, st_jump, st_round;   // state of the evaluation in flight, published by warp 0
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < Q.resident; i += SER_T) res[i] = Q.recs[i];
    if (tid == 0) { s_npend = 0; s_done = 0; }
    const rb2_mh_config &c = P.c;
    // chain state, kept by every lane of warp 0 (identical)
    int s = 0, round = 0, jump = 0, ok = 0, jump_a = 0, jump_r = 0, npend = 0;
    double cx = 0.0, cy = 0.0, sup = 0.0, Fc = 1.0, mh_std = Q.mh_std0, a_rate = Q.a_rate0;
    double qx = 0.0, qy = 0.0, w_q = 0.0, log_u = 0.0;
    int iter = 0;
    bool have_la = false;      // warp 0: the look-ahead of the previous iteration is valid
    int prev_s = -1;           // warp 0: chain of the previous evaluation
    for (;;) {
        if (warp == 0) {
            // the proposal of the current (chain, iteration)
            if (s >= Q.M) { if (lane == 0) s_done = 1; }
            else {
                qx = cx; qy = cy;
                if (!ok) {  // search for a favourable start, :1150-1180 (uniform over the emitter)
                    iter = -(round + 1);
                    double u, v;
                    if (have_la && s != prev_s && round == 0) { u = la_nu; v = la_nv; }
                    else if (have_la && s == prev_s && round >= 1) { u = la_ru; v = la_rv; }
                    else rand2(L.seed, s, iter, 0, 0, u, v);
                    qx = u * c.emit_dim[0] + c.emit_pos[0];
                    qy = v * c.emit_dim[1] + c.emit_pos[1];
                } else {
                    iter = jump;
                    double g0, g1;
                    if (have_la && s == prev_s) { g0 = la_g0; g1 = la_g1; log_u = la_lu; }
                    else {
                        draw_jump_normals(L, iter, s, g0, g1);
                        double u, v;
                        rand2(L.seed, s, iter, 2, 0, u, v);
                        log_u = log(u);
                    }
                    propose_apply(P, iter, mh_std, g0, g1, qx, qy);
                }
                w_q = w_theta_xy(P, qx, qy);
                if (lane == 0) { sp_x = qx; sp_y = qy; st_s = s; st_ok = ok; st_jump = jump; st_round = round; }
                prev_s = s;
            }
        }
        __syncthreads();
        if (s_done) break;
        const double px = sp_x, py = sp_y;
        const int n_tot = Q.n + s_npend;
        double acc = 0.0;
        bool close = false;
        if (NIC == 1 && L.far) {
            for (int i = tid; i < n_tot; i += SER_T) {
                const SurfRec &r = (i < Q.resident) ? res[i] : (i < Q.n ? Q.recs[i] : pend[i - Q.n]);
                acc = surf_term<NIC, false, true>(r, px, py, acc, L, close);
            }
        } else {
            for (int i = tid; i < n_tot; i += SER_T) {
                const SurfRec &r = (i < Q.resident) ? res[i] : (i < Q.n ? Q.recs[i] : pend[i - Q.n]);
                acc = surf_term<NIC, false>(r, px, py, acc, L, close);
            }
        }
        if (close) {  // a laterally close record: this thread's records again with the reference's sqrt / divide
            acc = 0.0;
            for (int i = tid; i < n_tot; i += SER_T) {
                const SurfRec &r = (i < Q.resident) ? res[i] : (i < Q.n ? Q.recs[i] : pend[i - Q.n]);
                acc = surf_term<NIC, true>(r, px, py, acc, L, close);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (warp == 1) {
            // the randoms of whatever iteration follows (a function of seed, chain, iteration only), one item per lane
            const int cs = st_s, cok = st_ok;
            const int jn = cok ? st_jump + 1 : 1;
            if (lane == 0 && jn <= c.ndim) { double g0, g1; draw_jump_normals(L, jn, cs, g0, g1); la_g0 = g0; la_g1 = g1; }
            else if (lane == 1 && jn <= c.ndim) { double u, v; rand2(L.seed, cs, jn, 2, 0, u, v); la_lu = log(u); }
            else if (lane == 2 && cs + 1 < Q.M) { double u, v; rand2(L.seed, cs + 1, -1, 0, 0, u, v); la_nu = u; la_nv = v; }
            else if (lane == 3 && !cok) { double u, v; rand2(L.seed, cs, -(st_round + 2), 0, 0, u, v); la_ru = u; la_rv = v; }
            __syncwarp();
            asm volatile("bar.sync 1, 64;" ::: "memory");  // with warp 0, behind its accept step
        }
        if (warp == 0) {
            double sum = red[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const double Fz = L.E_vac - L.fac * sum;
            bool chain_done = false;
            if (!ok) {
                if (Fz < 0.0) { cx = qx; cy = qy; Fc = Fz; sup = target_log_w(P, Fz, w_q); ok = 1; jump = 1; jump_a = 0; jump_r = 0; }
                else if (++round >= L.max_init) chain_done = true;  // "Failed to find spot for emission", :1160-1170
                if (ok && c.ndim < 1) chain_done = true;
            } else {
                const bool counted = jump > c.ndim_first;
                const bool unfav = (c.kind == 2) ? (Fz > 0.0) : (Fz >= 0.0);
                bool accept = false;
                if (!unfav) {
                    const double sup_new = target_log_w(P, Fz, w_q);
                    accept = (sup_new >= sup) || (log_u <= sup_new - sup);
                    if (accept) { cx = qx; cy = qy; sup = sup_new; Fc = Fz; }
                }
                if (counted) { if (accept) ++jump_a; else ++jump_r; }
                if (++jump > c.ndim) chain_done = true;
            }
            if (chain_done) {
                int emitted = 0;
                double D_f = HUGE_NEG, F_out = 1.0, ox = c.emit_pos[0], oy = c.emit_pos[1];
                if (ok) {
                    if (jump_a + jump_r > 0) {  // the per-chain step adaptation, :1250-1256 / MH_std_update :603-612
                        a_rate = (double)jump_a / (double)(jump_a + jump_r);
                        mh_std = fmin(fmax(mh_std * exp(c.std_gain * (a_rate - c.target_rate)), c.std_min), c.std_max);
                    }
                    const double w = w_theta_xy(P, cx, cy), sw = sqrt(w);
                    F_out = Fc; ox = cx; oy = cy;
                    if (c.kind == 2) { D_f = 0.0; emitted = 1; }
                    else {
                        D_f = P.b_FN * (sw * sw * sw) * v_y(P, Fc, w) / (-1.0 * Fc);  // Escape_Prob_log
                        double u, v;
                        rand2(L.seed, s, c.ndim + 1, 3, 0, u, v);
                        emitted = (Fc < 0.0 && log(u) <= D_f) ? 1 : 0;                // :339-379
                    }
                    if (emitted && npend >= SER_PMAX) emitted = 0;  // (never in practice: ~N_round / 10 emit)
                    if (emitted) {  // Add_Particle at z = 1 nm: the chains that follow see it
                        if (lane == 0) {
                            const double z = 1.0 * rb2k::length_scale, q = -1.0 * rb2k::q_0;
                            SurfRec r;
                            r.x = cx; r.y = cy; r.h0 = z; r.g0 = q * z;
                            if (L.nic >= 2) { r.h1 = 0.0; r.g1 = q; r.h2 = 0.0; r.g2 = 0.0; }
                            else { r.h1 = z - L.two_d; r.g1 = q * r.h1; r.h2 = z + L.two_d; r.g2 = q * r.h2; }
                            pend[npend] = r;
                        }
                        ++npend;
                    }
                }
                if (lane == 0) {
                    Q.df_out[s] = D_f; Q.F_out[s] = F_out; Q.emit_out[s] = emitted;
                    Q.pos_out[3 * s] = ox; Q.pos_out[3 * s + 1] = oy; Q.pos_out[3 * s + 2] = 0.0;
                    s_npend = npend;
                }
                ++s; ok = 0; round = 0; jump = 0; Fc = 1.0;
            }
            asm volatile("bar.sync 1, 64;" ::: "memory");  // warp 1's look-ahead is in shared memory
            have_la = true;
        }
        // (warp 0 publishes the next proposal behind the barrier at the top of the loop)
    }
    if (tid == 0) { Q.scal_out[0] = mh_std; Q.scal_out[1] = a_rate; }
}

// work units: (tiles of 32 points) x (particle chunks, whole 128-record sub-tiles)
MhPlan make_plan(const Rb2Ctx &ctx, int M, int G_max)
{
    const rb2_config &gc = ctx.cfg;
    const int n = ctx.n;
    MhPlan L{};
    L.M = M; L.n = n; L.n_tiles = (M + 31) / 32;
    if (n > 0) {
        const long long max_ns = (n + MHB - 1) / MHB;
        // one wave when the work is small (cheap joins), up to 16 waves when it is large (even finish)
        const long long avail = (long long)L.n_tiles * max_ns;
        long long units = std::min<long long>(avail, (long long)G_max * std::min<long long>(16, std::max<long long>(1, avail / ((long long)8 * G_max))));
        long long ns = std::min<long long>(std::max<long long>((units + L.n_tiles - 1) / L.n_tiles, 1), max_ns);
        int chunk = (int)((n + ns - 1) / ns);
        chunk = ((chunk + MHB - 1) / MHB) * MHB;
        L.j_chunk = chunk;
        L.nsplit = (n + chunk - 1) / chunk;
    }
    L.units = L.n_tiles * L.nsplit;
    L.two_d = 2.0 * gc.d;
    L.E_vac = rb2_make_step_params(gc).pl.E_z;
    L.fac = (gc.image_charge ? 2.0 : 1.0) * rb2k::div_fac_c;
    L.nic = gc.N_ic_max;
    L.far = (rb2_make_step_params(gc).pl.far_ok && rb2_far_allowed(ctx)) ? 1 : 0;
    return L;
}

}  // namespace

// E_z at M points of the cathode plane (z = 0), planar geometry: the mirror-antisymmetric form of the image
// series (file header), 31 FP64 instructions per (point, particle) at N_ic_max = 1.
int rb2_launch_surface_field(Rb2Ctx &ctx, const double *d_pts, int M, double *d_Ez)
{
    if (M < 1) return RB2_OK;
    const rb2_config &gc = ctx.cfg;
    const int n = ctx.n;
    const int NIC = !gc.image_charge ? -1 : (gc.N_ic_max >= 2 ? 2 : gc.N_ic_max);
    const MhPlan L = make_plan(ctx, M, 4 * ctx.sm_count);
    int rc = rb2_ensure_stage(ctx, (size_t)L.nsplit * M + (size_t)8 * n + 2, 0);
    if (rc) return rc;
    cudaStream_t st = ctx.stream;
    double *partial = ctx.d_stage_d;
    SurfRec *d_recs = reinterpret_cast<SurfRec *>(ctx.d_stage_d + ((((size_t)L.nsplit * M) + 1) & ~(size_t)1));
    int launches = 1;
    if (n > 0) {
        k_surf_pack<<<(n + 255) / 256, 256, 0, st>>>(ctx.a.pq, n, L.two_d, gc.image_charge ? gc.N_ic_max : 0, d_recs);
#define RB2_GO(N) k_surface_field<N><<<L.units, MHB, 0, st>>>(d_pts, d_recs, L, partial)
        if (NIC < 0) RB2_GO(-1); else if (NIC == 0) RB2_GO(0); else if (NIC == 1) RB2_GO(1); else RB2_GO(2);
#undef RB2_GO
        launches += 2;
    }
    k_surface_join<<<(M + 3) / 4, 128, 0, st>>>(partial, L, d_Ez);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(launches);
    return RB2_OK;
}

// ---- lock-step chains on the hyperboloid tip ----------------------------------------------------------------------
// Metro_algo_tip_v3 (src/mod_emission_tip.f90:1241-1390) for all candidates of a time step together, the scheme of
// Metropolis_Hastings_rectangle_J_batch applied to the tip's (xi, phi) chains (the host mirror's
// Metro_algo_tip_v3_batch).  A jump is three launches queued back to back -- proposals -> the tip field kernel of
// rb2_field_batch (k_pair<tip, field> + finalise, device pointers) -> accept / reject + MH_std update -- and all
// `ndim` jumps are queued without a host wait in between; only the rounds of the search for a favourable start are
// host-stepped (one or two per call).  The host loop paid one rb2_field_batch round trip (~70 us) and ~60 us of host
// math per jump.  One CTA runs the proposal / accept kernels (a few hundred chains; the shared step needs the
// accept counts of all chains anyway).
struct TipMh {
    int M, ndim_first;
    unsigned long long seed;
    double max_xi, eta_1, a_foci, shift_z;
    double sup_fac;              // (time_step / q_0) * a_FN / w_theta
    double esc_fac;              // b_FN * w_theta^(3/2)
    double *xi, *phi, *eta_f, *sup, *cur;   // chain state: [M], [M], [M], [M], [3M]
    double *w_xi, *w_phi;                  // proposals
    double *pts, *fld;                     // [3M] field points / fields of the current jump
    double *scal;                          // [0] MH_std, [1] a_rate
    int *ok, *valid, *bad;                 // [M], [M], [1]
    double *df_out;                        // [M]
};
constexpr double TIP_W = 4.7;  // w_theta of the tip, src/mod_emission_tip.f90:45
constexpr int TSUB = 32;       // particles per sub-tile of the persistent tip kernel (few particles: spread them over many CTAs)
constexpr int TIPB = 256;

__device__ __forceinline__ void tip_xyz(const TipMh &T, double xi, double phi, double *out)  // xyz_corr, src/mod_hyperboloid_tip.f90:78-89
{
    const double xy = T.a_foci * sqrt((xi * xi - 1.0) * (1.0 - T.eta_1 * T.eta_1));
    out[0] = xy * cos(phi);
    out[1] = xy * sin(phi);
    out[2] = T.a_foci * xi * T.eta_1 + T.shift_z;
}
__device__ __forceinline__ void tip_surface_normal(const TipMh &T, double x, double y, double *n)  // surface_normal, :25-34
{
    const double eta_fac = T.eta_1 / sqrt(1 - T.eta_1 * T.eta_1);
    const double div_fac = -1.0 / sqrt(x * x + y * y + (T.a_foci * T.a_foci) * (1 - T.eta_1 * T.eta_1));
    const double nx = eta_fac * x * div_fac, ny = eta_fac * y * div_fac, nz = 1.0;
    const double nrm = sqrt(nx * nx + ny * ny + nz * nz);
    n[0] = nx / nrm; n[1] = ny / nrm; n[2] = nz / nrm;
}
__device__ __forceinline__ double tip_field_normal(const TipMh &T, const double *pos, const double *f)  // :156-163, :25-34
{
    const double eta_fac = T.eta_1 / sqrt(1 - T.eta_1 * T.eta_1);
    const double div_fac = -1.0 / sqrt(pos[0] * pos[0] + pos[1] * pos[1] + (T.a_foci * T.a_foci) * (1 - T.eta_1 * T.eta_1));
    const double nx = eta_fac * pos[0] * div_fac, ny = eta_fac * pos[1] * div_fac, nz = 1.0;
    const double nrm = sqrt(nx * nx + ny * ny + nz * nz);
    return (nx / nrm) * f[0] + (ny / nrm) * f[1] + (nz / nrm) * f[2];
}
__device__ __forceinline__ double tip_target_log(const MhParams &P, const TipMh &T, double eta_f, double xi)  // Tip_fe_target_log, :1213-1219
{
    const double t = t_y(P, eta_f, TIP_W);
    const double sup = T.sup_fac / (t * t) * (eta_f * eta_f);
    return log(fmax(sup, TINY)) + 0.5 * log(xi * xi - T.eta_1 * T.eta_1);
}

// iter < 0: search round -iter (uniform over the surface for chains without a start); iter >= 1: jump iter
__global__ void __launch_bounds__(TIPB) k_tip_propose(MhParams P, TipMh T, int iter)
{
    const double two_pi = 2.0 * RB2_PI;
    for (int c = threadIdx.x; c < T.M; c += TIPB) {
        const int ok = T.ok[c];
        int valid = 0;
        double nxi = T.xi[c], nphi = T.phi[c];
        if (iter < 0) {
            if (!ok) {
                double u, v;
                rand2(T.seed, c, iter, 0, 0, u, v);
                nxi = 1.0 + (T.max_xi - 1.0) * u;
                nphi = two_pi * v;
                valid = 1;
            }
        } else if (ok) {
            const double frac = (iter > T.ndim_first) ? T.scal[0] : 0.10;
            double g0 = 0.0, g1 = 0.0;
            for (int attempt = 0; attempt < 64; ++attempt) {  // box_muller = Marsaglia polar, src/mod_global.F90:578-595
                double u, v;
                rand2(T.seed, c, iter, 1, attempt, u, v);
                const double a = 2.0 * u - 1.0, b = 2.0 * v - 1.0, w = a * a + b * b;
                if (w < 1.0 && w > 0.0) {
                    const double f = sqrt((-2.0 * log(w)) / w);
                    g0 = a * f; g1 = b * f;
                    break;
                }
            }
            nxi = nxi + g0 * ((T.max_xi - 1.0) * frac);
            nphi = fmod(nphi + g1 * (two_pi * frac), two_pi);
            if (nphi < 0.0) nphi += two_pi;                 // Fortran modulo()
            if (nxi > T.max_xi) nxi = 2.0 * T.max_xi - nxi;  // reflection, :1300-1310
            if (nxi < 1.0) nxi = 2.0 - nxi;
            valid = !(nxi < 1.0 || nxi > T.max_xi);
        }
        T.valid[c] = valid;
        T.w_xi[c] = nxi; T.w_phi[c] = nphi;
        if (valid) tip_xyz(T, nxi, nphi, &T.pts[3 * c]);
        else { T.pts[3 * c] = T.cur[3 * c]; T.pts[3 * c + 1] = T.cur[3 * c + 1]; T.pts[3 * c + 2] = T.cur[3 * c + 2]; }
    }
}

__global__ void __launch_bounds__(TIPB) k_tip_accept(MhParams P, TipMh T, int iter)
{
    __shared__ int s_acc, s_rej, s_bad;
    if (threadIdx.x == 0) { s_acc = 0; s_rej = 0; s_bad = 0; }
    __syncthreads();
    int n_acc = 0, n_rej = 0, n_bad = 0;
    for (int c = threadIdx.x; c < T.M; c += TIPB) {
        const int ok = T.ok[c], valid = T.valid[c];
        if (iter < 0) {
            if (ok) continue;
            const double ef = tip_field_normal(T, &T.pts[3 * c], &T.fld[3 * c]);
            if (ef < 0.0) {
                T.xi[c] = T.w_xi[c]; T.phi[c] = T.w_phi[c]; T.eta_f[c] = ef; T.ok[c] = 1;
                for (int k = 0; k < 3; ++k) T.cur[3 * c + k] = T.pts[3 * c + k];
                T.sup[c] = tip_target_log(P, T, ef, T.w_xi[c]);
            } else n_bad++;
            continue;
        }
        if (!ok) continue;
        if (!valid) { n_rej++; continue; }
        const double ef = tip_field_normal(T, &T.pts[3 * c], &T.fld[3 * c]);
        if (ef >= 0.0) { n_rej++; continue; }
        const double sup_new = tip_target_log(P, T, ef, T.w_xi[c]), sup_old = T.sup[c];
        bool accept = sup_new >= sup_old;
        if (!accept) {
            double u, v;
            rand2(T.seed, c, iter, 2, 0, u, v);
            accept = log(u) <= sup_new - sup_old;
        }
        if (accept) {
            T.xi[c] = T.w_xi[c]; T.phi[c] = T.w_phi[c]; T.eta_f[c] = ef; T.sup[c] = sup_new;
            for (int k = 0; k < 3; ++k) T.cur[3 * c + k] = T.pts[3 * c + k];
            n_acc++;
        } else n_rej++;
    }
    if (n_acc) atomicAdd(&s_acc, n_acc);
    if (n_rej) atomicAdd(&s_rej, n_rej);
    if (n_bad) atomicAdd(&s_bad, n_bad);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (iter < 0) *T.bad = s_bad;
        else if (iter > T.ndim_first && s_acc + s_rej > 0) {  // :1370-1380
            const double a_rate = (double)s_acc / (double)(s_acc + s_rej);
            T.scal[1] = a_rate;
            T.scal[0] = fmin(fmax(T.scal[0] * exp(0.025 * (a_rate - 0.35)), 0.0005), 0.125);
        }
    }
}

__global__ void __launch_bounds__(TIPB) k_tip_finish(MhParams P, TipMh T)
{
    for (int c = threadIdx.x; c < T.M; c += TIPB) {
        if (T.ok[c]) {
            const double F = T.eta_f[c];
            T.df_out[c] = exp(T.esc_fac * v_y(P, F, TIP_W) / fabs(F));  // Escape_Prob_Tip, :1734-1760
        } else {  // no favourable spot: defined outputs, no emission
            tip_xyz(T, 1.0, 0.0, &T.cur[3 * c]);
            T.eta_f[c] = 1.0;
            T.df_out[c] = 0.0;
        }
    }
}

// ---- the tip chains as ONE persistent kernel (at most 512 chains, a few thousand particles) -------------------------
// The scheme of k_mh_small applied to the tip: CTA b works for tile t_b = b / Gs of 32 chains on its own share of the
// particles, resident in shared memory; every CTA keeps the state of all chains in registers (warp w holds tile w) and
// does every accept step redundantly behind the single barrier of a jump.  Three field components per chain instead of
// one; the pair arithmetic is tip_point_field (the same as k_tip_field_point / k_pair<tip, field>), the vacuum field
// and the normal component are added by the warp that owns the tile.  Proposals, targets and generator keys are those
// of k_tip_propose / k_tip_accept, so for one seed the two paths run the same chains up to rounding in the field sums.
struct TipSmall {
    int T, Gs, G, S, R;   // tiles, CTAs per tile, CTAs, 128-particle sub-tiles in total, most sub-tiles per CTA
    int n, ndim, max_init, do_ic;
    double mh_std0, a_rate0;
    double *partial;      // [2][G][3][32]
    unsigned *bar;
    double *eta_f_out, *df_out, *pos_out, *scal_out;
};

__device__ __forceinline__ void tip_normals(unsigned long long seed, int iter, int k, double &g0, double &g1)
{
    g0 = 0.0; g1 = 0.0;
    // box_muller = Marsaglia polar, src/mod_global.F90:578-595; the tail behind the rejection loop (draw_jump_normals)
    double a = 0.0, b = 0.0, w = 0.0;
    bool found = false;
    for (int attempt = 0; attempt < 64 && !found; ++attempt) {
        double u, v;
        rand2(seed, k, iter, 1, attempt, u, v);
        a = 2.0 * u - 1.0; b = 2.0 * v - 1.0; w = a * a + b * b;
        found = (w < 1.0 && w > 0.0);
    }
    if (found) {
        const double f = sqrt((-2.0 * log(w)) / w);
        g0 = a * f; g1 = b * f;
    }
}

template <int WPB>
__global__ void __launch_bounds__(WPB * 32) k_mh_tip_small(MhParams P, TipMh T, TipParams TP, TipSmall Q, const double4 *__restrict__ pq)
{
    constexpr int NT = WPB * 32, RPW = TSUB / WPB;
    extern __shared__ __align__(16) unsigned char tip_smem[];
    double4 *mine = reinterpret_cast<double4 *>(tip_smem);                                                // [R][TSUB]
    double *strands = reinterpret_cast<double *>(tip_smem + (size_t)Q.R * TSUB * sizeof(double4));      // [T][WPB][3][32]
    __shared__ double red[WPB][3][32];
    __shared__ double sp[3][32];
    __shared__ int s_acc[WPB], s_rej[WPB], s_bad[WPB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t_b = blockIdx.x / Q.Gs, g_b = blockIdx.x - t_b * Q.Gs;
    const int s0 = (int)((long long)Q.S * g_b / Q.Gs), s1 = (int)((long long)Q.S * (g_b + 1) / Q.Gs);
    for (int q = tid; q < (s1 - s0) * TSUB; q += NT) {
        const int j = s0 * TSUB + q;
        mine[q] = (j < Q.n) ? pq[j] : make_double4(1.0 + (double)q, 1.0, 1.0, 0.0);  // padding: charge 0, metres away
    }
    if (tid < WPB) { s_acc[tid] = 0; s_rej[tid] = 0; s_bad[tid] = 0; }
    __syncthreads();
    const double two_pi = 2.0 * RB2_PI;
    const int chain = warp * 32 + lane;
    const bool live = (warp < Q.T) && (chain < T.M);
    double xi = 1.0, phi = 0.0, eta_f = 1.0, sup = 0.0, cx = 0.0, cy = 0.0, cz = 0.0;
    double mh_std = Q.mh_std0, a_rate = Q.a_rate0;
    int ok = 0;
    if (live) { double c3[3]; tip_xyz(T, 1.0, 0.0, c3); cx = c3[0]; cy = c3[1]; cz = c3[2]; }
    unsigned phase = 0;
    int bad = T.M, round = 0, jump = 1;
    bool searching = true;
    double ahead_g0 = 0.0, ahead_g1 = 0.0;
    int ahead_iter = 0;
    for (;;) {
        if (searching && !(round < Q.max_init && bad > 0)) searching = false;
        if (!searching && jump > Q.ndim) break;
        const int iter = searching ? -(round + 1) : jump;
        // proposal of my chain (k_tip_propose)
        double nxi = xi, nphi = phi, qx = cx, qy = cy, qz = cz;
        int valid = 0;
        if (live) {
            if (iter < 0) {
                if (!ok) {
                    double u, v;
                    rand2(T.seed, chain, iter, 0, 0, u, v);
                    nxi = 1.0 + (T.max_xi - 1.0) * u;
                    nphi = two_pi * v;
                    valid = 1;
                }
            } else if (ok) {
                const double frac = (iter > T.ndim_first) ? mh_std : 0.10;
                double g0, g1;
                if (ahead_iter == iter) { g0 = ahead_g0; g1 = ahead_g1; }
                else tip_normals(T.seed, iter, chain, g0, g1);
                nxi = nxi + g0 * ((T.max_xi - 1.0) * frac);
                nphi = fmod(nphi + g1 * (two_pi * frac), two_pi);
                if (nphi < 0.0) nphi += two_pi;                 // Fortran modulo()
                if (nxi > T.max_xi) nxi = 2.0 * T.max_xi - nxi;  // reflection, :1300-1310
                if (nxi < 1.0) nxi = 2.0 - nxi;
                valid = !(nxi < 1.0 || nxi > T.max_xi);
            }
            if (valid) { double c3[3]; tip_xyz(T, nxi, nphi, c3); qx = c3[0]; qy = c3[1]; qz = c3[2]; }
        }
        if (warp == t_b) { sp[0][lane] = qx; sp[1][lane] = qy; sp[2][lane] = qz; }
        __syncthreads();
        const double px = sp[0][lane], py = sp[1][lane], pz = sp[2][lane];
        TipImage im_i{};
        double4 img_i = make_double4(0.0, 0.0, 0.0, 0.0);
        if (Q.do_ic) { im_i = tip_image_point(TP, px, py, pz); img_i = tip_image_packed(TP, px, py, pz); }
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (int t = 0; t < s1 - s0; ++t) {
            const double4 *rr = &mine[t * TSUB + warp * RPW];
#pragma unroll
            for (int k = 0; k < RPW; ++k) {
                const double4 pj = rr[k];
                double fx, fy, fz;
                bool close = false;
                tip_pair_fast_upper(img_i, Q.do_ic != 0, px, py, pz, pj, fx, fy, fz, close);
                if (close) tip_point_field(TP, im_i, Q.do_ic != 0, px, py, pz, pj, fx, fy, fz);  // within 1e-11 m: literal arithmetic
                ax = fma(pj.w, fx, ax); ay = fma(pj.w, fy, ay); az = fma(pj.w, fz, az);
            }
        }
        red[warp][0][lane] = ax; red[warp][1][lane] = ay; red[warp][2][lane] = az;
        __syncthreads();
        double *part = Q.partial + (size_t)(phase & 1u) * Q.G * 96;
        // the last warp (tile-free unless there are 16 tiles) joins, stores and arrives: see k_mh_small
        if (warp == WPB - 1) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double sum = red[0][c][lane];
#pragma unroll
                for (int w = 1; w < WPB; ++w) sum += red[w][c][lane];
                part[((size_t)blockIdx.x * 3 + c) * 32 + lane] = sum;
            }
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                atomicAdd(Q.bar, 1u);
            }
        }
        ++phase;
        // while the other CTAs arrive: everything of the accept step that does not need the field sums -- the vacuum
        // field and the surface normal at the proposal, log(u) of the accept test -- and the normals of the NEXT jump
        double fE_x = 0.0, fE_y = 0.0, fE_z = 0.0, nrm[3] = {0.0, 0.0, 1.0}, log_u = 0.0;
        if (live && valid) {
            rb2_tip_field_E(TP, qx, qy, qz, fE_x, fE_y, fE_z);
            tip_surface_normal(T, qx, qy, nrm);
            if (iter >= 1) {
                double u, v;
                rand2(T.seed, chain, iter, 2, 0, u, v);
                log_u = log(u);
            }
        }
        if (live && !searching && jump < Q.ndim) { ahead_iter = jump + 1; tip_normals(T.seed, ahead_iter, chain, ahead_g0, ahead_g1); }
        small_barrier_wait(Q.bar, phase * (unsigned)Q.G);
        // join: warp w adds the partial sums of CTAs w, w + WPB, ... of every tile (ascending), three components
        // (the loads of four tiles are issued together: tile after tile, each join paid its own L2 round trip)
        for (int t0 = 0; t0 < Q.T; t0 += 4) {
            double st[4][3];
#pragma unroll
            for (int u = 0; u < 4; ++u) { st[u][0] = 0.0; st[u][1] = 0.0; st[u][2] = 0.0; }
            for (int k = warp; k < Q.Gs; k += WPB) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (t0 + u < Q.T) {
                        const double *src = part + ((size_t)((t0 + u) * Q.Gs + k) * 3) * 32 + lane;
#pragma unroll
                        for (int c = 0; c < 3; ++c) st[u][c] += __ldcg(src + c * 32);
                    }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (t0 + u < Q.T) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) strands[(((size_t)(t0 + u) * WPB + warp) * 3 + c) * 32 + lane] = st[u][c];
                }
        }
        __syncthreads();
        bool acc_ = false, rej_ = false, bad_ = false;
        if (warp < Q.T) {
            double f3[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double *sw = strands + (((size_t)warp * WPB) * 3 + c) * 32 + lane;
                double sum = sw[0];
#pragma unroll
                for (int w = 1; w < WPB; ++w) sum += sw[(size_t)w * 96];
                f3[c] = sum;
            }
            if (live) {
                // E = E_vac(point) + 1/(4 pi eps0) * sum (k_tip_field_point), then the normal component (Field_normal)
                const double ef = nrm[0] * (fE_x + rb2k::div_fac_c * f3[0]) + nrm[1] * (fE_y + rb2k::div_fac_c * f3[1]) +
                                  nrm[2] * (fE_z + rb2k::div_fac_c * f3[2]);
                if (iter < 0) {
                    if (!ok) {
                        if (ef < 0.0) { xi = nxi; phi = nphi; eta_f = ef; ok = 1; cx = qx; cy = qy; cz = qz; sup = tip_target_log(P, T, ef, nxi); }
                        else bad_ = true;
                    }
                } else if (ok) {
                    bool accept = false;
                    if (valid && ef < 0.0) {
                        const double sup_new = tip_target_log(P, T, ef, nxi);
                        accept = (sup_new >= sup) || (log_u <= sup_new - sup);
                        if (accept) { xi = nxi; phi = nphi; eta_f = ef; sup = sup_new; cx = qx; cy = qy; cz = qz; }
                    }
                    acc_ = accept; rej_ = !accept;
                }
            }
            const int na = __popc(__ballot_sync(0xffffffffu, acc_)), nr = __popc(__ballot_sync(0xffffffffu, rej_));
            const int nb = __popc(__ballot_sync(0xffffffffu, bad_));
            if (lane == 0) { s_acc[warp] = na; s_rej[warp] = nr; s_bad[warp] = nb; }
        }
        __syncthreads();
        {
            int a = 0, r = 0, nb = 0;
#pragma unroll
            for (int w = 0; w < WPB; ++w) { a += s_acc[w]; r += s_rej[w]; nb += s_bad[w]; }
            if (iter > T.ndim_first && a + r > 0) {  // :1370-1380
                a_rate = (double)a / (double)(a + r);
                mh_std = fmin(fmax(mh_std * exp(0.025 * (a_rate - 0.35)), 0.0005), 0.125);
            }
            if (searching) { bad = nb; ++round; } else ++jump;
        }
    }
    if (blockIdx.x == 0) {
        if (live) {
            const int k = chain;
            if (ok) {
                Q.pos_out[3 * k] = cx; Q.pos_out[3 * k + 1] = cy; Q.pos_out[3 * k + 2] = cz;
                Q.eta_f_out[k] = eta_f;
                Q.df_out[k] = exp(T.esc_fac * v_y(P, eta_f, TIP_W) / fabs(eta_f));  // Escape_Prob_Tip, :1734-1760
            } else {  // no favourable spot: defined outputs, no emission
                double c3[3];
                tip_xyz(T, 1.0, 0.0, c3);
                Q.pos_out[3 * k] = c3[0]; Q.pos_out[3 * k + 1] = c3[1]; Q.pos_out[3 * k + 2] = c3[2];
                Q.eta_f_out[k] = 1.0;
                Q.df_out[k] = 0.0;
            }
        }
        if (tid == 0) { Q.scal_out[0] = mh_std; Q.scal_out[1] = a_rate; }
    }
}

// Returns RB2_ERR_ARG - 1000 ("does not apply") when the problem is too large for the persistent kernel.
static int launch_mh_tip_small(Rb2Ctx &ctx, const MhParams &P, TipMh T, int M, int ndim, double *eta_f_out, double *df_out,
                               double *pos_out, double *a_rate_io, double *mh_std_io)
{
    const rb2_config &gc = ctx.cfg;
    const int n = ctx.n;
    TipSmall Q{};
    Q.T = (M + 31) / 32;
    if (Q.T > 16) return RB2_ERR_ARG - 1000;
    const int WPB = Q.T <= 4 ? 4 : 16;
    const int max_ctas = (WPB == 4 ? 2 : 1) * ctx.sm_count;
    Q.S = std::max(1, (n + TSUB - 1) / TSUB);
    Q.Gs = std::max(1, std::min(Q.S, max_ctas / Q.T));
    Q.G = Q.T * Q.Gs;
    Q.R = std::max(1, (Q.S + Q.Gs - 1) / Q.Gs);
    if (Q.R > 32) return RB2_ERR_ARG - 1000;
    Q.n = n; Q.ndim = ndim; Q.max_init = 10000; Q.do_ic = gc.image_charge ? 1 : 0;
    Q.mh_std0 = fmin(fmax(*mh_std_io, 0.0005), 0.125);  // the clamp of :1250-1254 at entry
    Q.a_rate0 = *a_rate_io;
    const size_t off_part = (size_t)5 * M + 2;
    int rc = rb2_ensure_stage(ctx, off_part + (size_t)2 * Q.G * 96 + 2, 4);
    if (rc) return rc;
    double *d = ctx.d_stage_d;
    Q.eta_f_out = d; Q.df_out = d + M; Q.pos_out = d + 2 * (size_t)M; Q.scal_out = d + 5 * (size_t)M;
    Q.partial = d + off_part;
    Q.bar = reinterpret_cast<unsigned *>(ctx.d_stage_i);
    cudaStream_t st = ctx.stream;
    RB2_CUDA(cudaMemsetAsync(ctx.d_stage_i, 0, 4 * sizeof(int), st));
    void *kern = WPB == 4 ? (void *)k_mh_tip_small<4> : (void *)k_mh_tip_small<16>;
    const size_t smem = (size_t)Q.R * TSUB * sizeof(double4) + (size_t)Q.T * WPB * 96 * sizeof(double);
    if (smem > 48 * 1024) RB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        int occ = 0;
        RB2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WPB * 32, smem));
        if ((long long)occ * ctx.sm_count < Q.G) return RB2_ERR_ARG - 1000;
    }
    const StepParams SP = rb2_make_step_params(gc);
    TipParams TP = SP.tip;
    MhParams Pm = P;
    const double4 *pq = ctx.a.pq;
    void *args[] = {&Pm, &T, &TP, &Q, &pq};
    RB2_CUDA(cudaLaunchCooperativeKernel(kern, dim3(Q.G), dim3(WPB * 32), args, smem, st));
    RB2_LAUNCHED(1);
    double scal1[2];
    RB2_CUDA(cudaMemcpyAsync(eta_f_out, Q.eta_f_out, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaMemcpyAsync(df_out, Q.df_out, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaMemcpyAsync(pos_out, Q.pos_out, (size_t)3 * M * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaMemcpyAsync(scal1, Q.scal_out, sizeof(scal1), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaStreamSynchronize(st));
    *mh_std_io = scal1[0];
    *a_rate_io = scal1[1];
    return RB2_OK;
}

int rb2_launch_mh_tip(Rb2Ctx &ctx, int M, int ndim, unsigned long long seed, double *eta_f_out, double *df_out, double *pos_out,
                      double *a_rate_io, double *mh_std_io)
{
    if (M < 1) return RB2_OK;
    const rb2_config &gc = ctx.cfg;
    if (gc.geometry != RB2_GEOM_TIP) return rb2_fail(RB2_ERR_GEOMETRY, "rb2_mh_tip: hyperboloid-tip geometry only");
    if (ndim < 1) return rb2_fail(RB2_ERR_ARG, "ndim < 1");
    const double pi = RB2_PI, h_bar = 6.62607015e-34 / (2.0 * pi);
    MhParams P{};
    P.c.image_charge = gc.image_charge;
    P.l_const = rb2k::q_0 / (4.0 * pi * rb2k::epsilon_0);
    P.b_FN = -4.0 / (3.0 * h_bar) * sqrt(2.0 * rb2k::m_0 * rb2k::q_0);
    const double a_FN = (rb2k::q_0 * rb2k::q_0) / (16.0 * (pi * pi) * h_bar);
    TipMh T{};
    T.M = M; T.ndim_first = (int)lround(ndim * 0.25); T.seed = seed;
    T.max_xi = gc.max_xi; T.eta_1 = gc.eta_1; T.a_foci = gc.a_foci; T.shift_z = gc.shift_z;
    T.sup_fac = (gc.time_step / rb2k::q_0) * a_FN / TIP_W;
    { const double sw = sqrt(TIP_W); T.esc_fac = P.b_FN * (sw * sw * sw); }
    if (ctx.mh_small && M <= ctx.mh_small_max) {  // few chains, few particles: everything in one persistent kernel
        const int rc_small = launch_mh_tip_small(ctx, P, T, M, ndim, eta_f_out, df_out, pos_out, a_rate_io, mh_std_io);
        if (rc_small != RB2_ERR_ARG - 1000) return rc_small;
    }
    // scratch: 17 M doubles + 2 scalars, 2 M + 1 ints
    int rc = rb2_ensure_stage(ctx, (size_t)18 * M + 4, (size_t)2 * M + 4);
    if (rc) return rc;
    double *d = ctx.d_stage_d;
    T.xi = d; T.phi = d + M; T.eta_f = d + 2 * (size_t)M; T.sup = d + 3 * (size_t)M; T.cur = d + 4 * (size_t)M;
    T.w_xi = d + 7 * (size_t)M; T.w_phi = d + 8 * (size_t)M; T.pts = d + 9 * (size_t)M; T.fld = d + 12 * (size_t)M;
    T.df_out = d + 15 * (size_t)M; T.scal = d + 16 * (size_t)M;
    T.ok = ctx.d_stage_i; T.valid = T.ok + M; T.bad = T.valid + M;
    cudaStream_t st = ctx.stream;
    double scal[2] = {fmin(fmax(*mh_std_io, 0.0005), 0.125), *a_rate_io};  // the clamp of :1250-1254 at entry
    RB2_CUDA(cudaMemsetAsync(d, 0, (size_t)16 * M * sizeof(double), st));
    RB2_CUDA(cudaMemsetAsync(ctx.d_stage_i, 0, ((size_t)2 * M + 4) * sizeof(int), st));
    RB2_CUDA(cudaMemcpyAsync(T.scal, scal, sizeof(scal), cudaMemcpyHostToDevice, st));
    // search for favourable starts: host-stepped rounds (one or two in practice, 10000 at most like the reference)
    int bad = M;
    for (int r = 0; r < 10000 && bad > 0; ++r) {
        k_tip_propose<<<1, TIPB, 0, st>>>(P, T, -(r + 1));
        rc = rb2_launch_field(ctx, ctx.a.pq, ctx.n, nullptr, 0, T.pts, M, T.fld);
        if (rc) return rc;
        k_tip_accept<<<1, TIPB, 0, st>>>(P, T, -(r + 1));
        RB2_CUDA(cudaGetLastError());
        RB2_LAUNCHED(2);
        RB2_CUDA(cudaMemcpyAsync(&bad, T.bad, sizeof(int), cudaMemcpyDeviceToHost, st));
        RB2_CUDA(cudaStreamSynchronize(st));
    }
    // the jumps: queued back to back, no host wait
    for (int i = 1; i <= ndim; ++i) {
        k_tip_propose<<<1, TIPB, 0, st>>>(P, T, i);
        rc = rb2_launch_field(ctx, ctx.a.pq, ctx.n, nullptr, 0, T.pts, M, T.fld);
        if (rc) return rc;
        k_tip_accept<<<1, TIPB, 0, st>>>(P, T, i);
        RB2_LAUNCHED(2);
    }
    k_tip_finish<<<1, TIPB, 0, st>>>(P, T);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(1);
    RB2_CUDA(cudaMemcpyAsync(eta_f_out, T.eta_f, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaMemcpyAsync(df_out, T.df_out, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaMemcpyAsync(pos_out, T.cur, (size_t)3 * M * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaMemcpyAsync(scal, T.scal, sizeof(scal), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaStreamSynchronize(st));
    *mh_std_io = scal[0];
    *a_rate_io = scal[1];
    return RB2_OK;
}

// ---- the tip's supply grid on the device ----------------------------------------------------------------------------
// Do_Field_Emission_Tip_OLDCODE (src/mod_emission_tip.f90:431-481) sums Elec_Supply(A, F) (:1710-1718) over a 100 x 100
// (xi, phi) midpoint grid of the tip surface every time step.  Nodes, unit normals and patch areas depend on the geometry
// only: the host hands them over once (rb2_tip_supply_set_grid); a time step is then the tip field kernel on the
// resident nodes + one kernel that projects the field on the normal, applies the supply function and reduces 256 nodes
// per CTA in a fixed tree -- two numbers per CTA go back (the host adds them in CTA order) instead of 3M doubles each way
// and M exp / log calls on the host.
constexpr int SUPB = 256;
__global__ void __launch_bounds__(SUPB) k_tip_supply(MhParams P, int M, const double *__restrict__ nrm, const double *__restrict__ area,
                                                     const double *__restrict__ fld, double fac, double *__restrict__ part)
{
    __shared__ double s_ns[SUPB], s_F[SUPB];
    const int k = blockIdx.x * SUPB + threadIdx.x;
    double ns = 0.0, F = 0.0;
    if (k < M) {
        F = nrm[3 * k] * fld[3 * k] + nrm[3 * k + 1] * fld[3 * k + 1] + nrm[3 * k + 2] * fld[3 * k + 2];  // Field_normal
        if (F < 0.0) {
            const double t = t_y(P, F, TIP_W);
            ns = area[k] * fac * (F * F) / (t * t);
        }
    }
    s_ns[threadIdx.x] = ns; s_F[threadIdx.x] = F;
    __syncthreads();
    for (int o = SUPB / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s_ns[threadIdx.x] += s_ns[threadIdx.x + o]; s_F[threadIdx.x] += s_F[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = s_ns[0]; part[2 * blockIdx.x + 1] = s_F[0]; }
}

int rb2_tip_supply_set_grid_impl(Rb2Ctx &ctx, int M, const double *pts, const double *nrm, const double *area)
{
    if (ctx.cfg.geometry != RB2_GEOM_TIP) return rb2_fail(RB2_ERR_GEOMETRY, "rb2_tip_supply_set_grid: hyperboloid-tip geometry only");
    RB2_CUDA(cudaStreamSynchronize(ctx.stream));
    if (ctx.d_sup_grid) RB2_CUDA(cudaFree(ctx.d_sup_grid));
    if (ctx.h_sup) RB2_CUDA(cudaFreeHost(ctx.h_sup));
    ctx.d_sup_grid = nullptr; ctx.h_sup = nullptr; ctx.sup_M = 0;
    const int nb = (M + SUPB - 1) / SUPB;
    RB2_CUDA(cudaMallocHost(&ctx.h_sup, (size_t)2 * nb * sizeof(double)));
    // [3M] nodes | [3M] normals | [M] areas | [3M] fields | [2 nb] partial sums
    RB2_CUDA(cudaMalloc(&ctx.d_sup_grid, ((size_t)10 * M + 2 * (size_t)nb) * sizeof(double)));
    RB2_CUDA(cudaMemcpy(ctx.d_sup_grid, pts, (size_t)3 * M * sizeof(double), cudaMemcpyHostToDevice));
    RB2_CUDA(cudaMemcpy(ctx.d_sup_grid + (size_t)3 * M, nrm, (size_t)3 * M * sizeof(double), cudaMemcpyHostToDevice));
    RB2_CUDA(cudaMemcpy(ctx.d_sup_grid + (size_t)6 * M, area, (size_t)M * sizeof(double), cudaMemcpyHostToDevice));
    ctx.sup_M = M;
    return RB2_OK;
}

int rb2_tip_supply_impl(Rb2Ctx &ctx, double *n_s_out, double *F_sum_out)
{
    const rb2_config &gc = ctx.cfg;
    if (gc.geometry != RB2_GEOM_TIP) return rb2_fail(RB2_ERR_GEOMETRY, "rb2_tip_supply: hyperboloid-tip geometry only");
    const int M = ctx.sup_M;
    if (M < 1 || !ctx.d_sup_grid) return rb2_fail(RB2_ERR_ARG, "rb2_tip_supply without rb2_tip_supply_set_grid");
    const int nb = (M + SUPB - 1) / SUPB;
    double *pts = ctx.d_sup_grid, *nrm = pts + (size_t)3 * M, *area = nrm + (size_t)3 * M, *fld = area + M, *part = fld + (size_t)3 * M;
    int rc = rb2_launch_field(ctx, ctx.a.pq, ctx.n, nullptr, 0, pts, M, fld);
    if (rc) return rc;
    const double pi = RB2_PI, h_bar = 6.62607015e-34 / (2.0 * pi);
    MhParams P{};
    P.c.image_charge = gc.image_charge;
    P.l_const = rb2k::q_0 / (4.0 * pi * rb2k::epsilon_0);
    const double a_FN = (rb2k::q_0 * rb2k::q_0) / (16.0 * (pi * pi) * h_bar);
    const double fac = a_FN * gc.time_step / (rb2k::q_0 * TIP_W);  // Elec_Supply = A a_FN F^2 dt / (q_0 w t^2)
    k_tip_supply<<<nb, SUPB, 0, ctx.stream>>>(P, M, nrm, area, fld, fac, part);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(1);
    double *h = ctx.h_sup;
    RB2_CUDA(cudaMemcpyAsync(h, part, (size_t)2 * nb * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    RB2_CUDA(cudaStreamSynchronize(ctx.stream));
    double ns = 0.0, Fs = 0.0;
    for (int b = 0; b < nb; ++b) { ns += h[2 * b]; Fs += h[2 * b + 1]; }
    *n_s_out = ns;
    if (F_sum_out) *F_sum_out = Fs;
    return RB2_OK;
}

// ---- one level of the planar supply quadrature on the device -------------------------------------------------------
// The host's stand-in for Cuba_Integrate (rh_emission.cpp) evaluates the integrand of Do_Surface_Integration_FE
// (integrand_cuba_fe_v, src/mod_field_emission_v2.F90:668-745) or ..._Simple (integrand_cuba_simple,
// src/mod_field_thermo_emission.F90:394-446) on K shifted copies of a rank-1 lattice, level by level.  A level here is:
// the nodes generated on the device from the K shifts, the cathode-plane field kernel on them, the integrand
// (Elec_Supply_V2 or J_GTF dt / q_0 at the node's work function) and a fixed-tree sum per shift and CTA -- 2 K nb
// numbers come back instead of M fields, and the M exp / log evaluations leave the host.
struct SupLevel {
    int K, n_done, n_new, kind;
    double shift[16];
    double fe_fac, gtf_fac;  // (dt / q_0) a_FN;  dt / q_0
};
__global__ void k_supply_nodes(MhParams P, SupLevel S, double *__restrict__ pts)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= S.K * S.n_new) return;
    const int r = m / S.n_new, k = m - r * S.n_new;
    const double a1 = 0.7548776662466927600495088963585286919, a2 = 0.5698402909980532659113999581195686488;
    const double kk = (double)(S.n_done + k + 1);
    double u = kk * a1 + S.shift[2 * r], v = kk * a2 + S.shift[2 * r + 1];
    u -= floor(u); v -= floor(v);
    pts[3 * m] = P.c.emit_pos[0] + u * P.c.emit_dim[0];
    pts[3 * m + 1] = P.c.emit_pos[1] + v * P.c.emit_dim[1];
    pts[3 * m + 2] = 0.0;
}
__global__ void __launch_bounds__(SUPB) k_supply_planar(MhParams P, SupLevel S, const double *__restrict__ pts,
                                                        const double *__restrict__ Ez, double *__restrict__ part)
{
    __shared__ double s_f[SUPB], s_E[SUPB];
    const int r = blockIdx.y, k = blockIdx.x * SUPB + threadIdx.x;
    double ff = 0.0, F = 0.0;
    if (k < S.n_new) {
        const int m = r * S.n_new + k;
        F = Ez[m];
        if (F < 0.0) {
            const double w = w_theta_xy(P, pts[3 * m], pts[3 * m + 1]);
            if (S.kind == 1) { const double t = t_y(P, F, w); ff = S.fe_fac / ((t * t) * w) * (F * F); }  // Elec_Supply_V2
            else ff = kevin_jgtf_v2(F, P.c.T_temp, w) * S.gtf_fac;
        }
    }
    s_f[threadIdx.x] = ff; s_E[threadIdx.x] = F;
    __syncthreads();
    for (int o = SUPB / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s_f[threadIdx.x] += s_f[threadIdx.x + o]; s_E[threadIdx.x] += s_E[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const size_t o = ((size_t)r * gridDim.x + blockIdx.x) * 2;
        part[o] = s_f[0]; part[o + 1] = s_E[0];
    }
}

int rb2_planar_supply_level_impl(Rb2Ctx &ctx, const rb2_mh_config *cfg, const double *w_theta_host, int kind, int K,
                                 const double *shifts, int n_done, int n_new, double *sums_out, double *ez_sum_out)
{
    const rb2_config &gc = ctx.cfg;
    if (gc.geometry != RB2_GEOM_PLANAR) return rb2_fail(RB2_ERR_GEOMETRY, "rb2_planar_supply_level: planar geometry only");
    if (K < 1 || K > 8 || n_new < 1 || n_done < 0) return rb2_fail(RB2_ERR_ARG, "rb2_planar_supply_level: K must be 1..8, n_new >= 1, n_done >= 0");
    if (kind != 1 && kind != 2) return rb2_fail(RB2_ERR_ARG, "rb2_planar_supply_level: kind must be 1 (field emission) or 2 (thermal-field)");
    const int nw = cfg->y_num * cfg->x_num;
    if (nw < 1 || nw > 96 * 96) return rb2_fail(RB2_ERR_ARG, "work function table must have 1..9216 cells");
    const size_t Mp = (size_t)K * n_new;
    if (Mp > (size_t)1 << 26) return rb2_fail(RB2_ERR_ARG, "rb2_planar_supply_level: too many nodes in one level");
    const int M = (int)Mp, nb = (n_new + SUPB - 1) / SUPB;
    // own buffers (the field kernel reallocates the shared staging area): table | nodes | E_z | partial sums
    const size_t need = (size_t)nw + 4 * (size_t)M + 2 * (size_t)K * nb + 8;
    if (need > ctx.supq_cap) {
        RB2_CUDA(cudaStreamSynchronize(ctx.stream));
        if (ctx.d_supq) RB2_CUDA(cudaFree(ctx.d_supq));
        if (ctx.h_supq) RB2_CUDA(cudaFreeHost(ctx.h_supq));
        ctx.d_supq = nullptr; ctx.h_supq = nullptr; ctx.supq_cap = 0;
        const size_t want = need + need / 2;
        RB2_CUDA(cudaMalloc(&ctx.d_supq, want * sizeof(double)));
        RB2_CUDA(cudaMallocHost(&ctx.h_supq, want * sizeof(double)));
        ctx.supq_cap = want;
    }
    double *d_w = ctx.d_supq, *d_pts = d_w + ((nw + 1) & ~1), *d_Ez = d_pts + (size_t)3 * M, *d_part = d_Ez + M;
    cudaStream_t st = ctx.stream;
    memcpy(ctx.h_supq, w_theta_host, (size_t)nw * sizeof(double));
    RB2_CUDA(cudaMemcpyAsync(d_w, ctx.h_supq, (size_t)nw * sizeof(double), cudaMemcpyHostToDevice, st));
    const double pi = RB2_PI, h_bar = 6.62607015e-34 / (2.0 * pi);
    MhParams P;
    P.c = *cfg;
    P.c.image_charge = gc.image_charge;
    P.w_theta = d_w;
    P.b_FN = -4.0 / (3.0 * h_bar) * sqrt(2.0 * rb2k::m_0 * rb2k::q_0);
    P.l_const = rb2k::q_0 / (4.0 * pi * rb2k::epsilon_0);
    SupLevel S{};
    S.K = K; S.n_done = n_done; S.n_new = n_new; S.kind = kind;
    for (int i = 0; i < 2 * K; ++i) S.shift[i] = shifts[i];
    S.gtf_fac = gc.time_step / rb2k::q_0;
    S.fe_fac = S.gtf_fac * (rb2k::q_0 * rb2k::q_0) / (16.0 * (pi * pi) * h_bar);
    k_supply_nodes<<<(M + 255) / 256, 256, 0, st>>>(P, S, d_pts);
    RB2_CUDA(cudaGetLastError());
    int rc = rb2_launch_surface_field(ctx, d_pts, M, d_Ez);
    if (rc) return rc;
    k_supply_planar<<<dim3(nb, K), SUPB, 0, st>>>(P, S, d_pts, d_Ez, d_part);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(2);
    double *h = ctx.h_supq + ((nw + 1) & ~1);
    RB2_CUDA(cudaMemcpyAsync(h, d_part, (size_t)2 * K * nb * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaStreamSynchronize(st));
    double ez = 0.0;
    for (int r = 0; r < K; ++r) {
        double f = 0.0;
        for (int b = 0; b < nb; ++b) { f += h[((size_t)r * nb + b) * 2]; ez += h[((size_t)r * nb + b) * 2 + 1]; }
        sums_out[r] = f;
    }
    if (ez_sum_out) *ez_sum_out = ez;
    return RB2_OK;
}

// At most 512 chains (T <= 16 tiles), the records of every CTA resident in shared memory (as many 128-record sub-tiles as fit): the single-barrier kernel, with
// 4 warps per CTA (two CTAs per SM) up to 4 tiles and 16 warps (one CTA per SM) beyond.
// Returns RB2_ERR_ARG - 1000 ("does not apply") when the problem is too large for it.
static int launch_mh_small(Rb2Ctx &ctx, const rb2_mh_config *cfg, const double *w_theta_host, int M, unsigned long long seed,
                           int NIC, int max_init, double *df_out, double *F_out, double *pos_out, double *a_rate_io,
                           double *mh_std_io)
{
    const rb2_config &gc = ctx.cfg;
    const int n = ctx.n, nw = cfg->y_num * cfg->x_num;
    MhSmall Q{};
    Q.T = (M + 31) / 32;
    if (Q.T > 16) return RB2_ERR_ARG - 1000;
    const int WPB = Q.T <= 4 ? 4 : 16;
    const int max_ctas = (WPB == 4 ? 2 : 1) * ctx.sm_count;
    Q.S = (n + MHB - 1) / MHB;
    Q.Gs = std::max(1, std::min(Q.S, max_ctas / Q.T));
    Q.G = Q.T * Q.Gs;
    Q.U = std::max(1, (n + 15) / 16);
    Q.R = std::max(1, (((Q.U + Q.Gs - 1) / Q.Gs) * 16 + MHB - 1) / MHB);  // sub-tiles of shared memory for the largest share
    {   // the records of a CTA stay in shared memory: as many 8 KB sub-tiles as fit next to the join strands (one CTA of
        // 16 warps per SM: ~216 KB; two CTAs of 4 warps: ~105 KB each).  Round 2 first capped this at 6 resp. 4 sub-tiles;
        // the Ion deck (20 k particles, 12 per CTA) then fell to the many-chain kernel at 21.8 us per jump.
        const size_t budget = (WPB == 4 ? 105 : 216) * 1024, strands = (size_t)Q.T * WPB * 32 * sizeof(double);
        const int r_max = budget > strands ? (int)((budget - strands) / (MHB * sizeof(SurfRec))) : 0;
        if (Q.R > r_max) return RB2_ERR_ARG - 1000;
    }
    MhPlan L{};
    L.M = M; L.n = n; L.n_tiles = Q.T; L.max_init = max_init; L.seed = seed;
    L.two_d = 2.0 * gc.d;
    L.E_vac = rb2_make_step_params(gc).pl.E_z;
    L.fac = (gc.image_charge ? 2.0 : 1.0) * rb2k::div_fac_c;
    L.nic = gc.N_ic_max;
    L.far = (rb2_make_step_params(gc).pl.far_ok && rb2_far_allowed(ctx)) ? 1 : 0;
    L.mh_std0 = *mh_std_io; L.a_rate0 = *a_rate_io;
    // scratch (doubles): 5M outputs + 2 scalars + table + 2 x G x 32 partials + 8n records; (ints): the arrival counter
    size_t off_part = (size_t)5 * M + 2 + nw;
    size_t off_recs = (off_part + (size_t)2 * Q.G * 32 + 1) & ~(size_t)1;  // 16-byte alignment
    int rc = rb2_ensure_stage(ctx, off_recs + (size_t)8 * n + 2, 4);
    if (rc) return rc;
    cudaStream_t st = ctx.stream;
    double *d = ctx.d_stage_d;
    MhState S{};
    S.df_out = d; S.F_out = d + M; S.pos_out = d + 2 * (size_t)M; S.scal_out = d + 5 * (size_t)M;
    double *d_w = S.scal_out + 2;
    Q.partial = d + off_part;
    SurfRec *d_recs = reinterpret_cast<SurfRec *>(d + off_recs);
    S.recs = d_recs;
    Q.bar = reinterpret_cast<unsigned *>(ctx.d_stage_i);
    MhParams P;
    P.c = *cfg;
    P.w_theta = d_w;
    const double pi = RB2_PI, h_bar = 6.62607015e-34 / (2.0 * pi);
    P.b_FN = -4.0 / (3.0 * h_bar) * sqrt(2.0 * rb2k::m_0 * rb2k::q_0);
    P.l_const = rb2k::q_0 / (4.0 * pi * rb2k::epsilon_0);
    RB2_CUDA(cudaMemcpyAsync(d_w, w_theta_host, (size_t)nw * sizeof(double), cudaMemcpyHostToDevice, st));
    RB2_CUDA(cudaMemsetAsync(ctx.d_stage_i, 0, 4 * sizeof(int), st));
    int launches = 1;
    if (n > 0) {
        k_surf_pack<<<(n + 255) / 256, 256, 0, st>>>(ctx.a.pq, n, L.two_d, gc.image_charge ? gc.N_ic_max : 0, d_recs);
        launches++;
    }
    void *kern = WPB == 4 ? (NIC < 0 ? (void *)k_mh_small<-1, 4> : NIC == 0 ? (void *)k_mh_small<0, 4>
                             : NIC == 1 ? (void *)k_mh_small<1, 4> : (void *)k_mh_small<2, 4>)
                          : (NIC < 0 ? (void *)k_mh_small<-1, 16> : NIC == 0 ? (void *)k_mh_small<0, 16>
                             : NIC == 1 ? (void *)k_mh_small<1, 16> : (void *)k_mh_small<2, 16>);
    // resident records + join strands [T][WPB][32]
    const size_t smem = (size_t)Q.R * MHB * sizeof(SurfRec) + (size_t)Q.T * WPB * 32 * sizeof(double);
    if (smem > 48 * 1024) RB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        int occ = 0;
        RB2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WPB * 32, smem));
        if ((long long)occ * ctx.sm_count < Q.G) return RB2_ERR_ARG - 1000;  // not co-resident: the many-chain kernel takes it
    }
    void *args[] = {&P, &S, &L, &Q};
    // cooperative launch: the arrival-counter barrier needs all G <= 2 x sm_count CTAs resident
    RB2_CUDA(cudaLaunchCooperativeKernel(kern, dim3(Q.G), dim3(WPB * 32), args, smem, st));
    RB2_LAUNCHED(launches);
    double scal1[2];
    RB2_CUDA(cudaMemcpyAsync(df_out, S.df_out, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaMemcpyAsync(F_out, S.F_out, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaMemcpyAsync(pos_out, S.pos_out, (size_t)3 * M * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaMemcpyAsync(scal1, S.scal_out, sizeof(scal1), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaStreamSynchronize(st));
    *mh_std_io = scal1[0];
    *a_rate_io = scal1[1];
    return RB2_OK;
}

int rb2_launch_mh_planar_serial(Rb2Ctx &ctx, const rb2_mh_config *cfg, const double *w_theta_host, int M, unsigned long long seed,
                                double *df_out, double *F_out, double *pos_out, int *emit_out, double *a_rate_io, double *mh_std_io)
{
    if (M < 1) return RB2_OK;
    const int nw = cfg->y_num * cfg->x_num;
    if (nw < 1 || nw > 96 * 96) return rb2_fail(RB2_ERR_ARG, "work function table must have 1..9216 cells");
    const rb2_config &gc = ctx.cfg;
    const int n = ctx.n;
    const int NIC = !gc.image_charge ? -1 : (gc.N_ic_max >= 2 ? 2 : gc.N_ic_max);
    MhPlan L{};
    L.M = M; L.n = n; L.max_init = 10000; L.seed = seed;
    L.two_d = 2.0 * gc.d;
    L.E_vac = rb2_make_step_params(gc).pl.E_z;
    L.fac = (gc.image_charge ? 2.0 : 1.0) * rb2k::div_fac_c;
    L.nic = gc.N_ic_max;
    L.far = (rb2_make_step_params(gc).pl.far_ok && rb2_far_allowed(ctx)) ? 1 : 0;
    MhSerial Q{};
    Q.M = M; Q.n = n;
    Q.mh_std0 = *mh_std_io; Q.a_rate0 = *a_rate_io;
    // scratch (doubles): 5M outputs + 2 scalars + table + 8n records; (ints): M flags
    const size_t off_recs = ((size_t)5 * M + 2 + nw + 1) & ~(size_t)1;
    int rc = rb2_ensure_stage(ctx, off_recs + (size_t)8 * n + 2, (size_t)M + 4);
    if (rc) return rc;
    cudaStream_t st = ctx.stream;
    double *d = ctx.d_stage_d;
    Q.df_out = d; Q.F_out = d + M; Q.pos_out = d + 2 * (size_t)M; Q.scal_out = d + 5 * (size_t)M;
    double *d_w = Q.scal_out + 2;
    SurfRec *d_recs = reinterpret_cast<SurfRec *>(d + off_recs);
    Q.recs = d_recs;
    Q.emit_out = ctx.d_stage_i;
    MhParams P;
    P.c = *cfg;
    P.w_theta = d_w;
    const double pi = RB2_PI, h_bar = 6.62607015e-34 / (2.0 * pi);
    P.b_FN = -4.0 / (3.0 * h_bar) * sqrt(2.0 * rb2k::m_0 * rb2k::q_0);
    P.l_const = rb2k::q_0 / (4.0 * pi * rb2k::epsilon_0);
    RB2_CUDA(cudaMemcpyAsync(d_w, w_theta_host, (size_t)nw * sizeof(double), cudaMemcpyHostToDevice, st));
    int launches = 1;
    if (n > 0) {
        k_surf_pack<<<(n + 255) / 256, 256, 0, st>>>(ctx.a.pq, n, L.two_d, gc.image_charge ? gc.N_ic_max : 0, d_recs);
        launches++;
    }
    void *kern = NIC < 0 ? (void *)k_mh_serial<-1> : NIC == 0 ? (void *)k_mh_serial<0> : NIC == 1 ? (void *)k_mh_serial<1> : (void *)k_mh_serial<2>;
    const size_t smem_max = 200 * 1024;
    const size_t pend_bytes = (size_t)SER_PMAX * sizeof(SurfRec);
    Q.resident = (int)std::min<size_t>((size_t)n, (smem_max - pend_bytes) / sizeof(SurfRec));
    const size_t smem = (size_t)Q.resident * sizeof(SurfRec) + pend_bytes;
    RB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    void *args[] = {&P, &L, &Q};
    RB2_CUDA(cudaLaunchKernel(kern, dim3(1), dim3(SER_T), args, smem, st));
    RB2_LAUNCHED(launches);
    double scal1[2];
    RB2_CUDA(cudaMemcpyAsync(df_out, Q.df_out, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaMemcpyAsync(F_out, Q.F_out, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaMemcpyAsync(pos_out, Q.pos_out, (size_t)3 * M * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaMemcpyAsync(emit_out, Q.emit_out, (size_t)M * sizeof(int), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaMemcpyAsync(scal1, Q.scal_out, sizeof(scal1), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaStreamSynchronize(st));
    *mh_std_io = scal1[0];
    *a_rate_io = scal1[1];
    return RB2_OK;
}

int rb2_launch_mh_planar(Rb2Ctx &ctx, const rb2_mh_config *cfg, const double *w_theta_host, int M, unsigned long long seed,
                         double *df_out, double *F_out, double *pos_out, double *a_rate_io, double *mh_std_io)
{
    if (M < 1) return RB2_OK;
    const int nw = cfg->y_num * cfg->x_num;
    if (nw < 1 || nw > 96 * 96) return rb2_fail(RB2_ERR_ARG, "work function table must have 1..9216 cells");
    const rb2_config &gc = ctx.cfg;
    const int n = ctx.n, max_init = 10000;
    const int NIC = !gc.image_charge ? -1 : (gc.N_ic_max >= 2 ? 2 : gc.N_ic_max);
    if (ctx.mh_small && M <= ctx.mh_small_max) {
        const int rc_small = launch_mh_small(ctx, cfg, w_theta_host, M, seed, NIC, max_init, df_out, F_out, pos_out, a_rate_io, mh_std_io);
        if (rc_small != RB2_ERR_ARG - 1000) return rc_small;  // else: too many particles for the resident scheme
    }
    void *kern = NIC < 0 ? (void *)k_mh_persistent<-1> : NIC == 0 ? (void *)k_mh_persistent<0>
               : NIC == 1 ? (void *)k_mh_persistent<1> : (void *)k_mh_persistent<2>;
    int occ = 0;
    RB2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, MHB, 0));
    if (occ < 1) return rb2_fail(RB2_ERR_CUDA, "sampler kernel does not fit on an SM");
    const int G_max = occ * ctx.sm_count;
    MhPlan L = make_plan(ctx, M, G_max);
    L.S = std::max(1, (n + MHB - 1) / MHB);  // an empty store still has one (empty) sub-tile per tile: its owner runs the accept step
    L.W = (long long)L.n_tiles * L.S;
    // the barrier costs grow with the number of CTAs and the field sums are throughput bound anyway: two CTAs per SM
    const int G = (int)std::max<long long>(1, std::min<long long>(std::min(G_max, ctx.mh_ctas_per_sm * ctx.sm_count), L.W));
    // the even split and, per tile, who contributes in which order
    std::vector<long long> h_wstart((size_t)G + 1);
    std::vector<int> h_tab((size_t)2 * G + L.n_tiles, 0);  // tfirst[G], kfirst[G], tcount[n_tiles]
    int *h_tfirst = h_tab.data(), *h_kfirst = h_tfirst + G, *h_tcount = h_kfirst + G;
    for (int g = 0; g <= G; ++g) h_wstart[g] = L.W * g / G;
    L.maxslots = 1;
    for (int g = 0; g < G; ++g) {
        const long long a = h_wstart[g], b = h_wstart[g + 1];
        if (b <= a) continue;
        const int t0 = (int)(a / L.S), t1 = (int)((b - 1) / L.S);
        h_tfirst[g] = t0;
        h_kfirst[g] = h_tcount[t0];
        for (int t = t0; t <= t1; ++t) L.maxslots = std::max(L.maxslots, ++h_tcount[t]);
    }
    const size_t n_partial = (size_t)L.n_tiles * L.maxslots * 32;
    // records resident in shared memory when every range is at most 6 sub-tiles (48 KB per CTA, two CTAs per SM)
    const long long max_range = (L.W + G - 1) / G;
    L.resident = (max_range <= 6) ? (int)max_range : 0;
    const size_t dyn_smem = (size_t)L.resident * MHB * sizeof(SurfRec);
    if (dyn_smem > 0) {
        RB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
        int occ2 = 0;
        RB2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, kern, MHB, dyn_smem));
        if ((long long)occ2 * ctx.sm_count < G) L.resident = 0;  // would not be co-resident: stream instead
    }
    const size_t dyn_smem_used = (size_t)L.resident * MHB * sizeof(SurfRec);
    L.max_init = max_init;
    L.seed = seed;
    L.mh_std0 = *mh_std_io; L.a_rate0 = *a_rate_io;
    // scratch (doubles): 4M state + partial sums + 5M outputs + table + 2 scalars + 8n records
    const size_t nd = (size_t)9 * M + n_partial + nw + 4 + (size_t)8 * n + 2 + (size_t)G + 1;
    const size_t ni = (size_t)M + 2 * ((size_t)cfg->ndim + 1) + max_init + h_tab.size() + (size_t)L.n_tiles + 1;
    int rc = rb2_ensure_stage(ctx, nd, ni);
    if (rc) return rc;
    cudaStream_t st = ctx.stream;
    double *d = ctx.d_stage_d;
    MhState S{};
    S.cur_x = d; S.cur_y = d + M; S.sup_cur = d + 2 * (size_t)M; S.F_cur = d + 3 * (size_t)M;
    S.df_out = d + 4 * (size_t)M; S.F_out = d + 5 * (size_t)M; S.pos_out = d + 6 * (size_t)M;
    S.partial = d + 9 * (size_t)M;
    double *d_w = S.partial + n_partial;
    S.scal_out = d_w + nw;
    long long *d_wstart = reinterpret_cast<long long *>(S.scal_out + 2);
    S.wstart = d_wstart;
    size_t off = (size_t)(S.scal_out + 2 - d) + (size_t)G + 1;
    off = (off + 1) & ~(size_t)1;  // 16-byte alignment for the records
    SurfRec *d_recs = reinterpret_cast<SurfRec *>(d + off);
    S.recs = d_recs;
    S.ok = ctx.d_stage_i;
    S.cnt = ctx.d_stage_i + M;
    S.bad = S.cnt + 2 * ((size_t)cfg->ndim + 1);
    int *d_tab = S.bad + max_init;
    S.tfirst = d_tab; S.kfirst = d_tab + G; S.tcount = d_tab + 2 * (size_t)G;
    S.tile_arrive = reinterpret_cast<unsigned *>(d_tab + h_tab.size());  // zeroed with the rest of the int scratch below
    S.bar = S.tile_arrive + L.n_tiles;
    MhParams P;
    P.c = *cfg;
    P.w_theta = d_w;
    const double pi = RB2_PI, h_bar = 6.62607015e-34 / (2.0 * pi);
    P.b_FN = -4.0 / (3.0 * h_bar) * sqrt(2.0 * rb2k::m_0 * rb2k::q_0);
    P.l_const = rb2k::q_0 / (4.0 * pi * rb2k::epsilon_0);
    RB2_CUDA(cudaMemcpyAsync(d_w, w_theta_host, (size_t)nw * sizeof(double), cudaMemcpyHostToDevice, st));
    RB2_CUDA(cudaMemsetAsync(ctx.d_stage_i, 0, ni * sizeof(int), st));
    RB2_CUDA(cudaMemsetAsync(d, 0, (size_t)4 * M * sizeof(double), st));
    RB2_CUDA(cudaMemcpyAsync(d_wstart, h_wstart.data(), ((size_t)G + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    RB2_CUDA(cudaMemcpyAsync(d_tab, h_tab.data(), h_tab.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    int launches = 1;
    if (n > 0) {
        k_surf_pack<<<(n + 255) / 256, 256, 0, st>>>(ctx.a.pq, n, L.two_d, gc.image_charge ? gc.N_ic_max : 0, d_recs);
        launches++;
    }
    void *args[] = {&P, &S, &L};
    RB2_CUDA(cudaLaunchCooperativeKernel(kern, dim3(G), dim3(MHB), args, dyn_smem_used, st));
    RB2_LAUNCHED(launches);
    double scal1[2];
    RB2_CUDA(cudaMemcpyAsync(df_out, S.df_out, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaMemcpyAsync(F_out, S.F_out, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaMemcpyAsync(pos_out, S.pos_out, (size_t)3 * M * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaMemcpyAsync(scal1, S.scal_out, sizeof(scal1), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaStreamSynchronize(st));
    *mh_std_io = scal1[0];
    *a_rate_io = scal1[1];
    return RB2_OK;
}

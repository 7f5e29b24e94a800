// rb2_pair_sym.cu -- pair-symmetric all-pairs acceleration for the planar geometry.
//
// The reference's CPU pair loop (src/mod_verlet.F90:763-884) evaluates every unordered pair
// ONCE and applies the result to both particles: Coulomb with the opposite sign, the image
// series with x, y mirrored and z unchanged (:862-871).  Its OpenACC kernel gives that up and
// gathers N(N-1) ordered pairs (:1187-1200).  This kernel keeps the halved arithmetic on the GPU
// without atomics and with a fixed summation order:
//
//   * particles are grouped in superblocks of 128 (one CTA = 4 warps); a CTA owns a target
//     superblock I and sweeps a group of source superblocks J >= I;
//   * inside a 128 x 128 tile each warp pairs its 32 targets (registers) with 32 "visitors"
//     that ROTATE through the lanes with warp shuffles, so every lane sees every visitor once;
//     the lane accumulates the force on its own particle and, in the visitor's travelling
//     accumulators, the reaction on the visitor.  Four rounds (warp w takes source warp-block
//     (w + r) mod 4) cover the tile; after each round the visitors' sums are added to
//     shared-memory slots that no other warp touches in that round;
//   * per tile the 128 source sums are stored once to a partial buffer indexed by
//     (source superblock, target superblock); the target sums are stored once per
//     (group, target superblock).  A reduce kernel adds both in ascending order.  To bound the
//     buffer the source superblocks are processed in bands (one launch pair per band);
//   * the diagonal tile (J == I) uses the gather form with per-element self mask and role sign.
//
// Per unordered pair at N_ic_max = 1: 81 FP64-pipe instructions + 6 MUFU + 14 SHFL, i.e. about
// 40 FP64 instructions per ordered pair interaction against 74 in the gather kernel.  Image roles
// (F8-ii) need no test here: J > I implies i < j for every pair of the tile.
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "rb2_internal.cuh"
#include <chrono>
#include "rb2_planar_math.cuh"

namespace {

#ifndef RB2_SYM_UNROLL
#define RB2_SYM_UNROLL 4
#endif
#ifndef RB2_SYM_LDSVIS
#define RB2_SYM_LDSVIS 1
#endif
#ifndef RB2_SYM_MINB
#define RB2_SYM_MINB 3
#endif
#ifndef RB2_EXACT_INLINE
#define RB2_EXACT_INLINE 0
#endif
#ifndef RB2_SYM_MINB2
#define RB2_SYM_MINB2 2
#endif
constexpr int SB = 128;  // particles per superblock = threads per CTA
constexpr int SYM_UNROLL = RB2_SYM_UNROLL;

struct SymGeom {
    int n, nsb, n_pad;         // nsb: 128-particle source tiles; n_pad: multiple of the target superblock size
    int nIb;                   // target superblocks (T * 128 particles each)
    int band_start, band_len;  // source tiles [band_start, band_start + band_len)
    int G, ngroups;            // source tiles per CTA group, groups in this band
    int K, nIsets;             // target superblocks a CTA takes one after the other (a "set"), sets in all
    int rank, world;           // CTA (set, grp) is owned by rank owner[set * ngroups + grp] (world > 1)
    const unsigned char *owner;  // this band's slice of the cost-balanced deal (null: one rank)
    const int2 *units;           // (set, grp) of this rank's units in this band, in the order they were dealt
};

__device__ __forceinline__ double rot1(double v, int src_lane) { return __shfl_sync(0xffffffffu, v, src_lane); }

// One round of a symmetric tile: this warp's 32 T targets against the 32 visitors of block wb0 (see k_pair_sym).
template <int NIC, int T, bool EXACT, bool FAR = false>
__device__ __forceinline__ void sym_round(const double *__restrict__ X, const double *__restrict__ Y, const double *__restrict__ Z,
                                          const double *__restrict__ Q, int wb0, int lane, int src_lane, const double (&xi)[T],
                                          const double (&yi)[T], const double (&zi)[T], const double (&qe)[T], const PlanarParams &P,
                                          double (&tx)[T], double (&ty)[T], double (&tz)[T], double &bx, double &by, double &bz,
                                          bool &close)
{
    const int home = wb0 + lane;
    double vx = X[home], vy = Y[home], vz = Z[home], vq = Q[home];
    bx = 0.0; by = 0.0; bz = 0.0;
#pragma unroll
    for (int s = 0; s < T; ++s) { tx[s] = 0.0; ty[s] = 0.0; tz[s] = 0.0; }
#pragma unroll (EXACT ? 1 : SYM_UNROLL)
    for (int k = 0; k < 32; ++k) {
#if RB2_SYM_LDSVIS
        // visitor coordinates straight from shared memory (conflict-free rotated index);
        // only the travelling accumulators go through the shuffle unit
        const int vi = wb0 + ((lane + k) & 31);
        vx = X[vi]; vy = Y[vi]; vz = Z[vi]; vq = Q[vi];
#endif
#pragma unroll
        for (int s = 0; s < T; ++s) {
            if (NIC == 1 && !EXACT) {
                // the hot path: N_ic_max = 1, i < j: one FMA chain for the image z-sum, far partners without softening
                const PairS w = FAR ? planar_weights_sym1<true>(xi[s], yi[s], zi[s], vx, vy, vz, P, close)
                                    : planar_weights_sym1<false>(xi[s], yi[s], zi[s], vx, vy, vz, P, close);
                const double ti = vq * w.U, tj = qe[s] * w.U;
                tx[s] = fma(w.dx, ti, tx[s]);
                ty[s] = fma(w.dy, ti, ty[s]);
                bx = fma(-w.dx, tj, bx);
                by = fma(-w.dy, tj, by);
                const double czz = w.dz * w.wc;
                tz[s] = fma(vq, w.icz + czz, tz[s]);
                bz = fma(qe[s], w.icz - czz, bz);
                continue;
            }
            const PairW w = EXACT ? planar_weights_exact<NIC>(xi[s], yi[s], zi[s], vx, vy, vz, P, false)
                                  : planar_weights<NIC>(xi[s], yi[s], zi[s], vx, vy, vz, P, close);
            // i < j here: evaluation at (z_i, z_j); reaction mirrored in x, y (src/mod_verlet.F90:862-871)
            const double ti = vq * w.U, tj = qe[s] * w.U;
            tx[s] = fma(w.dx, ti, tx[s]);
            ty[s] = fma(w.dy, ti, ty[s]);
            bx = fma(-w.dx, tj, bx);
            by = fma(-w.dy, tj, by);
            if (NIC < 0) {
                tz[s] = fma(w.dz, ti, tz[s]);
                bz = fma(-w.dz, tj, bz);
            } else {
                const double czz = w.dz * w.wc;
                const double icz = w.Zsame - w.Zopp;
                tz[s] = fma(vq, icz + czz, tz[s]);
                bz = fma(qe[s], icz - czz, bz);
            }
        }
#if !RB2_SYM_LDSVIS
        vx = rot1(vx, src_lane); vy = rot1(vy, src_lane); vz = rot1(vz, src_lane); vq = rot1(vq, src_lane);
#endif
        bx = rot1(bx, src_lane); by = rot1(by, src_lane); bz = rot1(bz, src_lane);
    }
}

// a laterally close pair on the diagonal tile: this thread's row of the tile again, reference arithmetic
template <int NIC>
__device__ __noinline__ Acc4 sym_diag_exact(const double *X, const double *Y, const double *Z, const double *Q, int tid, double xi,
                                            double yi, double zi, PlanarParams P)
{
    Acc4 a = {0.0, 0.0, 0.0, 0.0};
    for (int jj = 0; jj < SB; ++jj) {
        const double4 sj = make_double4(X[jj], Y[jj], Z[jj], Q[jj]);
        const double qe = (jj == tid) ? 0.0 : sj.w;
        const double qsg = (jj > tid) ? qe : -qe;
        planar_term_exact<NIC>(xi, yi, zi, sj, qe, qsg, P, a, jj < tid);
    }
    return a;
}

// The slow path of a round (a laterally close pair somewhere in it) out of line, so that the hot loop stays compact:
// at one CTA per SM (N ~ 1e4) the inlined copy cost 30 % in instruction fetch.  Arguments and results travel by value.
template <int T>
struct RoundIO {
    double xi[T], yi[T], zi[T], qe[T], tx[T], ty[T], tz[T], bx, by, bz;
};
template <int NIC, int T>
__device__ __noinline__ RoundIO<T> sym_round_exact(const double *X, const double *Y, const double *Z, const double *Q, int wb0, int lane,
                                                   int src_lane, RoundIO<T> io, PlanarParams P)
{
    bool dummy = false;
    sym_round<NIC, T, true>(X, Y, Z, Q, wb0, lane, src_lane, io.xi, io.yi, io.zi, io.qe, P, io.tx, io.ty, io.tz, io.bx, io.by, io.bz, dummy);
    return io;
}

// T targets per lane: the CTA's target superblock I holds the T source tiles T*I .. T*I+T-1 (sub-set s of the
// targets = tile T*I + s, one particle of each sub-set per thread).  Against source tile J a sub-set is evaluated
// symmetrically when its tile index is < J (then i < j for every pair), in the gather form when it IS tile J, and
// not at all when its tile index is > J (those pairs belong to the sweep of the other sub-set).  T = 2 halves the
// shared-memory reads, the shuffles and the partial-sum traffic per pair and doubles the independent work per warp.
template <int NIC, int T, bool FAR = false>
__global__ void __launch_bounds__(SB, T == 1 ? RB2_SYM_MINB : RB2_SYM_MINB2)
k_pair_sym(const double4 *__restrict__ pq, SymGeom g, PlanarParams P, double *__restrict__ bufI, double *__restrict__ bufJ)
{
    const int2 unit = g.units[blockIdx.x];  // this rank's units of the band, largest first
    const int iset = unit.x;
    const int grp = unit.y;
    const int I0 = iset * g.K;
    const int J0 = g.band_start + grp * g.G;
    const int J1 = min(J0 + g.G, min(g.band_start + g.band_len, g.nsb));
    if (max(J0, T * I0) >= J1) return;

    // source tile, double buffered (one CTA barrier per tile), and the visitors' reaction sums: one private copy per
    // warp (warp w meets each 32-particle block of the tile exactly once, so it just stores), again double buffered
    // because the sums of tile J are read after the barrier while the sweep of tile J+1 has started
    __shared__ double xs[2][SB], ys[2][SB], zs[2][SB], qs[2][SB];
    __shared__ double jacc[2][4][3][SB];
    // reaction sums of the group's source tiles, accumulated over the K target superblocks of this CTA (ascending I:
    // a fixed order) and stored ONCE at the end: [G][3][128]; element (.., tid) is only ever touched by thread tid
    extern __shared__ double accJ[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool direct = (g.K == 1);  // one superblock per CTA: the tile sums go straight to bufJ
    if (!direct)
        for (int k = tid; k < (J1 - J0) * 3 * SB; k += SB) accJ[k] = 0.0;
    // slots beyond the last particle: charge 0 and a position of their own, metres away from everything (at a shared
    // position -- e.g. the last particle's -- every pair among them would look "laterally close" and take the slow path)
    auto load = [&](int k) { return k < g.n ? pq[k] : make_double4(1.0 + (double)(k - g.n), 0.0, 1.0, 0.0); };
    const int src_lane = (lane + 1) & 31;

    for (int pass = 0; pass < g.K; ++pass) {
        const int I = I0 + pass;
        const int Jbeg = max(J0, T * I);
        if (I >= g.nIb || Jbeg >= J1) break;  // the triangle: later superblocks start even further right
        double xi[T], yi[T], zi[T], qi[T], ax[T], ay[T], az[T];
#pragma unroll
        for (int s = 0; s < T; ++s) {
            const int i = (I * T + s) * SB + tid;
            const double4 p = load(i);
            xi[s] = p.x; yi[s] = p.y; zi[s] = p.z;
            qi[s] = (i < g.n) ? p.w : 0.0;  // padding lanes: charge 0
            ax[s] = 0.0; ay[s] = 0.0; az[s] = 0.0;
        }
        double4 pj_next;
        __syncthreads();  // the previous pass is done with the staging buffers
        {
            const int j = Jbeg * SB + tid;
            const double4 pj = load(j);
            xs[0][tid] = pj.x; ys[0][tid] = pj.y; zs[0][tid] = pj.z;
            qs[0][tid] = (j < g.n) ? pj.w : 0.0;
            const int jn = j + SB;
            pj_next = load(jn);
        }
        __syncthreads();
        int cur = 0;
        for (int J = Jbeg; J < J1; ++J, cur ^= 1) {
            const double *__restrict__ X = xs[cur], *__restrict__ Y = ys[cur], *__restrict__ Z = zs[cur], *__restrict__ Q = qs[cur];
            const int rel = J - T * I;  // >= 0; sub-set s is symmetric iff s < rel, diagonal iff s == rel
            if (rel < T) {
                // gather form of tile J against itself (ordered pairs, per-element self mask and role sign)
#pragma unroll
                for (int s = 0; s < T; ++s) {
                    if (s != rel) continue;
                    Acc4 a = {0.0, 0.0, 0.0, 0.0};
                    bool close = false;
                    for (int jj = 0; jj < SB; ++jj) {
                        const double4 sj = make_double4(X[jj], Y[jj], Z[jj], Q[jj]);
                        const double qe = (jj == tid) ? 0.0 : sj.w;
                        const double qsg = (jj > tid) ? qe : -qe;
                        bool c1 = false;  // the self pair (offset 0) is not a close pair: it is masked out by qe = 0
                        planar_term<NIC>(xi[s], yi[s], zi[s], sj, qe, qsg, P, a, c1);
                        close = close || (c1 && jj != tid);
                    }
                    if (close) a = sym_diag_exact<NIC>(X, Y, Z, Q, tid, xi[s], yi[s], zi[s], P);
                    ax[s] += a.x; ay[s] += a.y; az[s] += a.z + a.t;
                }
            }
            if (rel >= 1) {
                double qe[T];  // a sub-set that is not symmetric against this tile neither pushes nor collects
#pragma unroll
                for (int s = 0; s < T; ++s) qe[s] = (s < rel) ? qi[s] : 0.0;
                for (int r = 0; r < 4; ++r) {
                    const int wb0 = ((warp + r) & 3) * 32;
                    const int home = wb0 + lane;
                    double bx, by, bz;          // reaction on the visitor, travels with it
                    double tx[T], ty[T], tz[T]; // force on my particles from this round
                    bool close = false;
                    sym_round<NIC, T, false, FAR>(X, Y, Z, Q, wb0, lane, src_lane, xi, yi, zi, qe, P, tx, ty, tz, bx, by, bz, close);
                    // a laterally close pair anywhere in this warp's 32 x 32T block: the round again with the reference's
                    // sqrt / divide (warp-uniform branch: the shuffles inside need every lane)
#if RB2_EXACT_INLINE
                    if (__any_sync(0xffffffffu, close))
                        sym_round<NIC, T, true>(X, Y, Z, Q, wb0, lane, src_lane, xi, yi, zi, qe, P, tx, ty, tz, bx, by, bz, close);
#else
                    if (__any_sync(0xffffffffu, close)) {
                        RoundIO<T> io;
#pragma unroll
                        for (int s = 0; s < T; ++s) { io.xi[s] = xi[s]; io.yi[s] = yi[s]; io.zi[s] = zi[s]; io.qe[s] = qe[s]; }
                        io = sym_round_exact<NIC, T>(X, Y, Z, Q, wb0, lane, src_lane, io, P);
#pragma unroll
                        for (int s = 0; s < T; ++s) { tx[s] = io.tx[s]; ty[s] = io.ty[s]; tz[s] = io.tz[s]; }
                        bx = io.bx; by = io.by; bz = io.bz;
                    }
#endif
                    // 32 rotations by one lane: every visitor is back at its home lane
#pragma unroll
                    for (int s = 0; s < T; ++s)
                        if (s < rel) { ax[s] += tx[s]; ay[s] += ty[s]; az[s] += tz[s]; }
                    jacc[cur][warp][0][home] = bx; jacc[cur][warp][1][home] = by; jacc[cur][warp][2][home] = bz;
                }
            }
            // stage the next tile into the other buffer (its last readers passed the previous barrier)
            if (J + 1 < J1) {
                const int jn = (J + 1) * SB + tid;
                xs[cur ^ 1][tid] = pj_next.x; ys[cur ^ 1][tid] = pj_next.y; zs[cur ^ 1][tid] = pj_next.z;
                qs[cur ^ 1][tid] = (jn < g.n) ? pj_next.w : 0.0;
                const int jnn = jn + SB;
                if (J + 2 < J1) pj_next = load(jnn);
            }
            __syncthreads();
            if (rel >= 1) {
                // block b of the tile was met by warp (b - r) & 3 in round r: add in round order, then onto the tile's
                // running sum over this CTA's target superblocks
                const int b = warp;
                double *aj = accJ + (size_t)(J - J0) * 3 * SB + tid;
                const size_t base = (((size_t)iset * g.band_len + (J - g.band_start)) * 3) * SB + tid;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double sum = ((jacc[cur][b][c][tid] + jacc[cur][(b + 3) & 3][c][tid]) + jacc[cur][(b + 2) & 3][c][tid]) +
                                       jacc[cur][(b + 1) & 3][c][tid];
                    if (direct) bufJ[base + c * SB] = sum;
                    else aj[c * SB] += sum;
                }
            }
        }
#pragma unroll
        for (int s = 0; s < T; ++s) {
            const size_t ib = (size_t)grp * 3 * g.n_pad + (size_t)(I * T + s) * SB + tid;
            bufI[ib] = ax[s];
            bufI[ib + g.n_pad] = ay[s];
            bufI[ib + 2 * (size_t)g.n_pad] = az[s];
        }
    }
    // the group's source sums over this CTA's superblocks, once: bufJ[set][tile of the band][3][128]
    for (int J = max(J0, T * I0); J < J1 && !direct; ++J) {
        const size_t base = (((size_t)iset * g.band_len + (J - g.band_start)) * 3) * SB + tid;
        const double *aj = accJ + (size_t)(J - J0) * 3 * SB + tid;
#pragma unroll
        for (int c = 0; c < 3; ++c) bufJ[base + c * SB] = aj[c * SB];
    }
}

// raw[c][p] += (sum over this band's groups of the target sums) + (sum over the sets of target superblocks that met
// tile J(p) of the source sums); only slots written by this rank are read.  One CTA per 128-particle
// tile, RQ threads per particle: thread (q, tid) adds every RQ-th term in ascending order, the RQ sums are joined
// in ascending q -- a fixed order, and short dependent chains (at N = 1e4 a single thread per particle spent 55 us
// on ~160 sequential loads next to a 300 us pair kernel).
constexpr int RQ = 4;
template <int T>
__global__ void __launch_bounds__(SB * RQ)
k_sym_reduce(SymGeom g, const double *__restrict__ bufI, const double *__restrict__ bufJ, double *__restrict__ raw)
{
    __shared__ double part[RQ][SB];
    const int Jp = blockIdx.x;  // 128-particle tile of this particle
    const int c = blockIdx.y;   // component (x, y, z): one CTA each, so that small systems still fill the SMs
    const int It = Jp / T;      // its target superblock
    const int Iset_t = It / g.K;
    const int tid = threadIdx.x & (SB - 1), q = threadIdx.x / SB;
    const int p = Jp * SB + tid;
    double s0 = 0.0;
    // as a target (superblock It, taken by the CTAs of its set)
    for (int grp = q; grp < g.ngroups; grp += RQ) {
        const int J0 = g.band_start + grp * g.G;
        const int J1 = min(J0 + g.G, min(g.band_start + g.band_len, g.nsb));
        if (max(J0, T * It) >= J1) continue;
        if (g.owner && g.owner[Iset_t * g.ngroups + grp] != g.rank) continue;
        s0 += bufI[((size_t)grp * 3 + c) * g.n_pad + p];
    }
    // as a source (tile Jp), when Jp lies in this band: the sets whose first superblock starts left of Jp
    if (Jp >= g.band_start && Jp < g.band_start + g.band_len) {
        const int grp = (Jp - g.band_start) / g.G;
        const int nset = min(g.nIsets, (Jp + T * g.K - 1) / (T * g.K));  // T * K * set < Jp: the set's first superblock lies left of the tile
        double t0 = 0.0;
#pragma unroll 4
        for (int st = q; st < nset; st += RQ) {
            if (g.owner && g.owner[st * g.ngroups + grp] != g.rank) continue;
            t0 += bufJ[(((size_t)st * g.band_len + (Jp - g.band_start)) * 3 + c) * SB + tid];
        }
        s0 += t0;
    }
    part[q][tid] = s0;
    __syncthreads();
    if (q == 0) {
#pragma unroll
        for (int k = 1; k < RQ; ++k) s0 += part[k][tid];
        raw[(size_t)c * g.n_pad + p] += s0;
    }
}


// a_i = ( q_i/(4 pi eps0) * raw_i + q_i * E_z zhat ) / m_i   (src/mod_verlet.F90:1333-1338)
__global__ void k_sym_finalize(int n, int n_pad, const double *__restrict__ raw, const double4 *__restrict__ pq,
                               const double *__restrict__ mass, PlanarParams P, double *__restrict__ acc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double q_1 = pq[i].w;
    const double qd_1 = q_1 * rb2k::div_fac_c;
    const double im_1 = 1.0 / mass[i];
    acc[3 * i] = (qd_1 * raw[i]) * im_1;
    acc[3 * i + 1] = (qd_1 * raw[(size_t)n_pad + i]) * im_1;
    acc[3 * i + 2] = (qd_1 * raw[2 * (size_t)n_pad + i] + q_1 * P.E_z) * im_1;
}

// Grow-only scratch.  The particle count of a real run changes every step (emission, absorption), so grow by half
// again: with exact sizes a slowly filling diode paid a cudaFree + cudaMalloc (milliseconds) on most steps.
int ensure_bytes(double **p, size_t *have, size_t want, size_t limit = 0)
{
    if (want > *have) {
        if (*p) RB2_CUDA(cudaFree(*p));
        *p = nullptr;
        *have = 0;
        size_t ask = want + want / 2;
        if (limit && ask > limit) ask = limit > want ? limit : want;
        if (cudaMalloc(p, ask) != cudaSuccess) {
            cudaGetLastError();
            ask = want;
            RB2_CUDA(cudaMalloc(p, ask));
        }
        *have = ask;
    }
    return RB2_OK;
}

}  // namespace

// Partial raw sums of this rank (rank 0 of 1: everything) into ctx.sym_raw[3][n_pad].
// K x G of the work units (see rb2_launch_accel_sym_partial): host arithmetic only, shared with rb2_sym_plan_probe.
static void sym_unit_shape(int nsb, int nIb, int T, int world, int sm_count, double waves, int kmax, int gmax, int &K_out, int &G_out)
{
    const int Gmax = (T == 1) ? 12 : 24;   // shared memory: 32 KB static + 3 KB x G, 3 (T = 1) or 2 CTAs per SM
    // unit size from the whole triangle: about sym_waves waves of units per rank and evaluation (not per band: tying the
    // unit to the band width shrank the units whenever the budget shrank the bands, which costs more scratch per tile,
    // which shrinks the bands ...)
    const double tri = 0.5 * (double)nIb * (double)nsb, slots = (double)sm_count * (4 / T) * world;
    double KG = tri / (slots * waves);
    // single-tile units pay ~5 % for their fixed cost (prologue, 6 KB of sums stored per 128T x 128 pairs): when the work
    // per rank is small, prefer K G = 4 over the full number of waves, down to 16 waves (8 ranks at N = 1e5: the slowest
    // rank's kernel at 95 % of 1/8 of the undivided one instead of 90 %, profiles/rank_emulation_r02.log)
    if (KG < 4.0 && waves > 16.0) KG = std::min(4.0, tri / (slots * 16.0));
    int G = (int)sqrt((double)T * KG);     // traffic per pair ~ T / G + 1 / K: G = T K at the optimum
    if (G > Gmax) G = Gmax;
    if (G > gmax) G = gmax;
    if (G > nsb) G = nsb;
    if (G < 1) G = 1;
    int K = (int)(KG / G);
    if (K > kmax) K = kmax;
    if (K < 1) K = 1;
    K_out = K; G_out = G;
}

// Source tiles per band: everything when it fits the budget, else whole groups of G tiles.
static int sym_band_width(int nsb, double tile_bytes, int G, size_t budget)
{
    int Wb = nsb;
    if (tile_bytes * nsb > (double)budget) {
        Wb = (int)((double)budget / tile_bytes);
        Wb = std::max(G, (Wb / G) * G);  // whole groups per band
        if (Wb > nsb) Wb = nsb;
    }
    return Wb;
}

// The deal of the work units of every band to the ranks: by cost (tile pairs in the unit; the triangle clips the ones
// near the diagonal), largest first to the least loaded rank, ties in group-major order.  tab: owner of unit (set, group)
// per band (world > 1 only), mine: this rank's units in the order they were dealt.  Pure host arithmetic: every rank
// computes the same table (tests/test_partition_gloo.py checks that through rb2_sym_plan_probe).
static void sym_deal_units(int nsb, int nIb, int T, int K, int G, int Wb, int world, int rank, std::vector<unsigned char> &tab,
                           std::vector<int2> &mine, std::vector<size_t> &unit_off, std::vector<int> &unit_cnt,
                           std::vector<long long> *rank_cost)
{
    std::vector<size_t> owner_off;
    size_t total = 0;
    for (int b0 = 0; b0 < nsb; b0 += Wb) {
        const int blen = (b0 + Wb <= nsb) ? Wb : (nsb - b0);
        const int nI = (b0 + blen - 1) / T + 1;
        owner_off.push_back(total);
        total += (size_t)((nI + K - 1) / K) * ((blen + G - 1) / G);
    }
    tab.assign(world > 1 ? total : 0, 0);
    mine.clear(); unit_off.clear(); unit_cnt.clear();
    if (rank_cost) rank_cost->assign((size_t)world, 0);
    constexpr bool order_gr_major = true;  // ties in group-major order: neighbouring CTAs share their source tiles (0.4 % at 8 ranks)
    std::vector<std::pair<long long, int>> units;  // (-cost, index): ascending sort = largest first, index order on ties
    size_t bi = 0;
    for (int b0 = 0; b0 < nsb; b0 += Wb, ++bi) {
        const int blen = (b0 + Wb <= nsb) ? Wb : (nsb - b0);
        const int nI = (b0 + blen - 1) / T + 1, nsets = (nI + K - 1) / K, ngr = (blen + G - 1) / G;
        units.clear();
        for (int is = 0; is < nsets; ++is)
            for (int gr = 0; gr < ngr; ++gr) {
                const int J0 = b0 + gr * G, J1 = std::min(J0 + G, std::min(b0 + blen, nsb));
                long long cost = 0;
                for (int I = is * K; I < std::min(is * K + K, nIb); ++I) {
                    const int Jb = std::max(J0, T * I);
                    if (Jb >= J1) break;
                    cost += (long long)T * (J1 - Jb);          // T sub-sets x tiles ...
                    for (int J = Jb; J < std::min(J1, T * I + T); ++J) cost -= (T - 1 - (J - T * I));  // ... less what the diagonal clips
                }
                if (cost > 0) units.emplace_back(-cost, order_gr_major ? gr * nsets + is : is * ngr + gr);
            }
        std::sort(units.begin(), units.end());
        std::vector<long long> load((size_t)world, 0);
        unit_off.push_back(mine.size());
        for (const auto &u : units) {
            int best = 0;
            for (int r = 1; r < world; ++r) if (load[(size_t)r] < load[(size_t)best]) best = r;
            load[(size_t)best] += -u.first;
            const int is = order_gr_major ? u.second % nsets : u.second / ngr, gr = order_gr_major ? u.second / nsets : u.second % ngr;
            if (world > 1) tab[owner_off[bi] + (size_t)is * ngr + gr] = (unsigned char)best;
            if (best == rank) mine.push_back(make_int2(is, gr));
        }
        unit_cnt.push_back((int)(mine.size() - unit_off.back()));
        if (rank_cost) for (int r = 0; r < world; ++r) (*rank_cost)[(size_t)r] += load[(size_t)r];
    }
}

int rb2_launch_accel_sym_partial(Rb2Ctx &ctx, const double4 *pq, int n)
{
    if (n < 1) return RB2_OK;
    const rb2_config &c = ctx.cfg;
    if (c.geometry != RB2_GEOM_PLANAR) return rb2_fail(RB2_ERR_GEOMETRY, "the pair-symmetric kernel implements the planar geometry only");
    // targets per lane: two from 25000 particles on (tools/sym_tpl_sweep.py, profiles/sym_tpl_sweep_r02.log: 1.521 vs
    // 1.528 ms at 24000, 2.046 vs 2.022 at 28000, 0.690 vs 0.723 at 16000; 5.5 % for T = 2 at 1e6), one below (half as
    // many, twice as long CTAs do not fill the machine at small N)
    const int T = ctx.sym_tpl == 1 ? 1 : (ctx.sym_tpl == 2 ? 2 : (n >= 25000 ? 2 : 1));
    SymGeom g{};
    g.n = n;
    g.nsb = (n + SB - 1) / SB;
    g.nIb = (g.nsb + T - 1) / T;
    g.n_pad = g.nIb * T * SB;
    g.rank = ctx.pair_rank;
    g.world = ctx.pair_world < 1 ? 1 : ctx.pair_world;
    // Work units: a CTA takes a SET of K consecutive target superblocks one after the other against a GROUP of G source
    // tiles whose reaction sums it keeps in shared memory (3 KB per tile) and stores once.  Per (superblock, tile) pair
    // that is 3 KB T / G of target sums + 3 KB / K of source sums written (and read back by the reduce kernel) instead
    // of 3 KB per pair with K = 1 (round 1: 96 GB per evaluation at N = 1e6).  K x G is sized for about sym_waves waves
    // of CTAs per launch and rank.  Over several ranks the units of a band are dealt out by COST (tile pairs in the
    // unit: the triangle clips the ones near the diagonal), largest first to the least loaded rank -- every rank computes
    // the same table, cached until the geometry changes.  (Round 1 dealt them round-robin, (I + grp) % world, and needed
    // world^2 times more, i.e. smaller, units to balance: at N = 1e5 on 8 GPUs single-tile CTAs whose fixed cost was 10 %.)
    // The source tiles are processed in bands that fit the scratch budget (default 2 GiB).
    int K = 1, G = 1;
    sym_unit_shape(g.nsb, g.nIb, T, g.world, ctx.sm_count, ctx.sym_waves, ctx.sym_kmax, ctx.sym_gmax, K, G);
    // band width from the scratch budget.  The layout is not compacted by owner: over `world` ranks every rank fills
    // 1 / world of the slots it allocates, so the budget (2 GiB of partial sums a rank actually writes) scales with it.
    size_t budget = ctx.sym_budget_bytes * (size_t)g.world;
    const size_t nsets = (size_t)(g.nIb + K - 1) / K;
    const double tile_bytes = (double)nsets * 3 * SB * sizeof(double) + 3.0 * g.n_pad * sizeof(double) / G;
    if (tile_bytes * g.nsb > (double)(ctx.sym_bufJ_bytes + ctx.sym_bufI_bytes)) {
        // the whole triangle does not fit what we hold: never ask for more than half of what is free (plus what we
        // already hold).  cudaMemGetInfo costs a fraction of a millisecond, so only look when it can matter.
        size_t fr = 0, tot = 0;
        if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) {
            const size_t cap = (fr + ctx.sym_bufJ_bytes + ctx.sym_bufI_bytes) / 2;
            if (budget > cap) budget = cap;
        }
    }
    const int Wb = sym_band_width(g.nsb, tile_bytes, G, budget);
    g.K = K;
    g.nIsets = (g.nIb + K - 1) / K;
    const int ngroups_max = (Wb + G - 1) / G;
    // banded: the sizes sit at the budget already -- allocate exactly (no head room for a growing particle count)
    const size_t wantJ = (size_t)g.nIsets * Wb * 3 * SB * sizeof(double), wantI = (size_t)ngroups_max * 3 * g.n_pad * sizeof(double);
    int rc = ensure_bytes(&ctx.sym_bufJ, &ctx.sym_bufJ_bytes, wantJ, Wb < g.nsb ? wantJ : 0);
    if (rc) return rc;
    rc = ensure_bytes(&ctx.sym_bufI, &ctx.sym_bufI_bytes, wantI, Wb < g.nsb ? wantI : 0);
    if (rc) return rc;
    if (ctx.p2p_world > 1) {
        // peers read the partial sums in place: they live in the exported exchange block (rb2_p2p.cu)
        ctx.sym_raw_cur = rb2_p2p_begin_evaluation(ctx, g.n_pad);
        if (!ctx.sym_raw_cur) return RB2_ERR_CAPACITY;
    } else {
        rc = ensure_bytes(&ctx.sym_raw, &ctx.sym_raw_bytes, (size_t)3 * g.n_pad * sizeof(double));
        if (rc) return rc;
        ctx.sym_raw_cur = ctx.sym_raw;
    }
    ctx.sym_n_pad = g.n_pad;
    cudaStream_t st = ctx.stream;
    const StepParams SP = rb2_make_step_params(c);
    // the work units of every band: dealt to the ranks by cost (one table for the reduce kernel, which must know whose
    // slots were written), and THIS rank's units as a list in the order they were dealt (largest first) -- the launch
    // grid is that list, so no CTA starts just to find that its unit is empty or somebody else's (a full 2-D grid at
    // N = 1e5 on 8 ranks launched 305 000 CTAs for 19 000 units), and the clipped units near the diagonal run last
    std::vector<size_t> owner_off;
    {
        const unsigned long long key[8] = {(unsigned long long)g.nsb, (unsigned long long)g.nIb, (unsigned long long)T, (unsigned long long)K,
                                           (unsigned long long)G, (unsigned long long)Wb, (unsigned long long)g.world, (unsigned long long)g.rank};
        size_t total = 0;
        for (int b0 = 0; b0 < g.nsb; b0 += Wb) {
            const int blen = (b0 + Wb <= g.nsb) ? Wb : (g.nsb - b0);
            const int nI = (b0 + blen - 1) / T + 1;
            owner_off.push_back(total);
            total += (size_t)((nI + K - 1) / K) * ((blen + G - 1) / G);
        }
        if (memcmp(key, ctx.sym_owner_key, sizeof(key)) != 0 || !ctx.sym_units) {
            if (ctx.capturing) return rb2_fail(RB2_ERR_CUDA, "pair-symmetric work units changed inside a graph capture");
            const auto t_plan0 = std::chrono::steady_clock::now();
            std::vector<unsigned char> tab;
            std::vector<int2> mine;
            sym_deal_units(g.nsb, g.nIb, T, K, G, Wb, g.world, g.rank, tab, mine, ctx.sym_unit_off, ctx.sym_unit_cnt, nullptr);
            if (g.world > 1 && total > ctx.sym_owner_cap) {
                if (ctx.sym_owner) RB2_CUDA(cudaFree(ctx.sym_owner));
                ctx.sym_owner = nullptr; ctx.sym_owner_cap = 0;
                RB2_CUDA(cudaMalloc(&ctx.sym_owner, total + total / 2 + 256));
                ctx.sym_owner_cap = total + total / 2 + 256;
            }
            if (mine.size() + 1 > ctx.sym_units_cap) {
                if (ctx.sym_units) RB2_CUDA(cudaFree(ctx.sym_units));
                ctx.sym_units = nullptr; ctx.sym_units_cap = 0;
                const size_t cap = mine.size() + mine.size() / 2 + 256;
                RB2_CUDA(cudaMalloc(&ctx.sym_units, cap * sizeof(int2)));
                ctx.sym_units_cap = cap;
            }
            if (g.world > 1) RB2_CUDA(cudaMemcpyAsync(ctx.sym_owner, tab.data(), total, cudaMemcpyHostToDevice, st));
            if (!mine.empty()) RB2_CUDA(cudaMemcpyAsync(ctx.sym_units, mine.data(), mine.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
            RB2_CUDA(cudaStreamSynchronize(st));  // the host vectors go out of scope
            memcpy(ctx.sym_owner_key, key, sizeof(key));
            ctx.sym_plans += 1;
            ctx.sym_plan_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_plan0).count();
        }
    }
    RB2_CUDA(rb2_event_record(ctx, ctx.ev_a0));
    RB2_CUDA(cudaMemsetAsync(ctx.sym_raw_cur, 0, (size_t)3 * g.n_pad * sizeof(double), st));
    int launches = 0;
    for (int b0 = 0; b0 < g.nsb; b0 += Wb) {
        g.band_start = b0;
        g.band_len = (b0 + Wb <= g.nsb) ? Wb : (g.nsb - b0);
        g.G = G;
        g.ngroups = (g.band_len + G - 1) / G;
        g.owner = (g.world > 1) ? ctx.sym_owner + owner_off[(size_t)(b0 / Wb)] : nullptr;
        const size_t bi = (size_t)(b0 / Wb);
        g.units = ctx.sym_units + ctx.sym_unit_off[bi];
        dim3 grid((unsigned)ctx.sym_unit_cnt[bi]), block(SB);
        const size_t dyn = (size_t)G * 3 * SB * sizeof(double);
#define RB2_GO(N, F)                                                                                     \
    do {                                                                                                 \
        const unsigned bit = 1u << ((((N) + 1) * 2 + (T - 1)) * 2 + ((F) ? 1 : 0));                      \
        if (T == 1) {                                                                                    \
            if (!(ctx.sym_attr_mask & bit)) { RB2_CUDA(cudaFuncSetAttribute(k_pair_sym<N, 1, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 3 * SB * 8)); ctx.sym_attr_mask |= bit; } \
            k_pair_sym<N, 1, F><<<grid, block, dyn, st>>>(pq, g, SP.pl, ctx.sym_bufI, ctx.sym_bufJ);       \
        } else {                                                                                         \
            if (!(ctx.sym_attr_mask & bit)) { RB2_CUDA(cudaFuncSetAttribute(k_pair_sym<N, 2, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 3 * SB * 8)); ctx.sym_attr_mask |= bit; } \
            k_pair_sym<N, 2, F><<<grid, block, dyn, st>>>(pq, g, SP.pl, ctx.sym_bufI, ctx.sym_bufJ);       \
        }                                                                                                \
    } while (0)
        if (grid.x == 0) { /* nothing of this band is ours */ }
        else if (!c.image_charge) RB2_GO(-1, false);
        else if (c.N_ic_max == 0) RB2_GO(0, false);
        else if (c.N_ic_max == 1) { if (SP.pl.far_ok && rb2_far_allowed(ctx)) RB2_GO(1, true); else RB2_GO(1, false); }
        else RB2_GO(2, false);
#undef RB2_GO
        RB2_CUDA(cudaGetLastError());
        if (T == 1) k_sym_reduce<1><<<dim3(g.nsb, 3), SB * RQ, 0, st>>>(g, ctx.sym_bufI, ctx.sym_bufJ, ctx.sym_raw_cur);
        else k_sym_reduce<2><<<dim3(g.nIb * 2, 3), SB * RQ, 0, st>>>(g, ctx.sym_bufI, ctx.sym_bufJ, ctx.sym_raw_cur);
        RB2_CUDA(cudaGetLastError());
        launches += 2;
    }
    RB2_LAUNCHED(launches);
    ctx.last_grid_x = g.nsb; ctx.last_grid_y = (g.nsb + Wb - 1) / Wb; ctx.last_block = SB; ctx.last_split = G * 1000 + K;
    return RB2_OK;
}

int rb2_launch_accel_sym_finalize(Rb2Ctx &ctx, const double4 *pq, const double *mass, int n, double *acc_out)
{
    if (n < 1) return RB2_OK;
    if (ctx.p2p_world > 1) return rb2_launch_accel_sym_exchange_finalize(ctx, pq, mass, n, acc_out);
    const StepParams SP = rb2_make_step_params(ctx.cfg);
    k_sym_finalize<<<(n + 255) / 256, 256, 0, ctx.stream>>>(n, ctx.sym_n_pad, ctx.sym_raw_cur, pq, mass, SP.pl, acc_out);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(1);
    RB2_CUDA(rb2_event_record(ctx, ctx.ev_a1));
    return RB2_OK;
}


// Host-only view of the planning above (no device needed, no context): the unit shape, the band width and THIS rank's
// work units for a system of n particles dealt to `world` ranks.  The multi-process tests run it on every rank and check
// that the lists are disjoint, cover the triangle and balance the cost (tests/test_partition_gloo.py).
extern "C" int rb2_sym_plan_probe(int n, int tpl, int world, int rank, int sm_count, double waves, int kmax, int gmax, double budget_mb,
                                  int *shape_out, long long *rank_cost_out, int *units_out, int cap, int *n_units_out,
                                  unsigned long long *table_hash_out)
{
    if (n < 1 || world < 1 || world > 255 || rank < 0 || rank >= world || sm_count < 1 || waves < 1.0 || kmax < 1 || gmax < 1 || budget_mb <= 0.0)
        return rb2_fail(RB2_ERR_ARG, "rb2_sym_plan_probe: bad argument");
    const int T = tpl == 1 ? 1 : (tpl == 2 ? 2 : (n >= 25000 ? 2 : 1));
    const int nsb = (n + SB - 1) / SB, nIb = (nsb + T - 1) / T, n_pad = nIb * T * SB;
    int K = 1, G = 1;
    sym_unit_shape(nsb, nIb, T, world, sm_count, waves, kmax, gmax, K, G);
    const size_t budget = (size_t)(budget_mb * 1048576.0) * (size_t)world;
    const size_t nsets = (size_t)(nIb + K - 1) / K;
    const double tile_bytes = (double)nsets * 3 * SB * sizeof(double) + 3.0 * n_pad * sizeof(double) / G;
    const int Wb = sym_band_width(nsb, tile_bytes, G, budget);
    std::vector<unsigned char> tab;
    std::vector<int2> mine;
    std::vector<size_t> unit_off;
    std::vector<int> unit_cnt;
    std::vector<long long> cost;
    sym_deal_units(nsb, nIb, T, K, G, Wb, world, rank, tab, mine, unit_off, unit_cnt, &cost);
    if (shape_out) { shape_out[0] = T; shape_out[1] = K; shape_out[2] = G; shape_out[3] = Wb; shape_out[4] = nsb; shape_out[5] = nIb; }
    if (rank_cost_out) for (int r = 0; r < world; ++r) rank_cost_out[r] = cost[(size_t)r];
    if (n_units_out) *n_units_out = (int)mine.size();
    if (units_out) {
        size_t k = 0;
        for (size_t b = 0; b < unit_off.size(); ++b)
            for (int u = 0; u < unit_cnt[b] && (int)k < cap; ++u, ++k) {
                const int2 v = mine[unit_off[b] + (size_t)u];
                units_out[3 * k] = (int)b; units_out[3 * k + 1] = v.x; units_out[3 * k + 2] = v.y;
            }
    }
    if (table_hash_out) {
        unsigned long long h = 1469598103934665603ull;
        for (unsigned char c : tab) { h ^= c; h *= 1099511628211ull; }
        *table_hash_out = h;
    }
    return RB2_OK;
}

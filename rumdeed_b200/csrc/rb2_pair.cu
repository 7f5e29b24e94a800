// rb2_pair.cu -- the O(N^2) kernels: all-pairs acceleration (K1 planar, K2 tip) and the
// batched surface-field evaluation (K4 planar, K5 tip).
//
// Replaces reference src/mod_verlet.F90:1217-1429 (Calculate_Acceleration_Particles_ACC)
// and :1635-1911 (Calc_Field_at_Batch), including src/acc_ic_planar_series.inc and
// src/acc_tip_*.inc.  Not a translation: the reference runs one thread per i over all j
// from global memory.  Here
//   * j-particles {x,y,z,q} (32 B) stream through a 3-stage shared-memory ring filled by
//     TMA bulk copies (cp.async.bulk + mbarrier), one elected thread issues;
//   * the j-range is split over blockIdx.y so that every launch has >= 8 waves of
//     equal-cost CTAs on 148 SMs (no tail), partial sums are written once and reduced in
//     fixed order by a finalize kernel (deterministic, no atomics);
//   * the planar image series shares dx, dy, dx^2+dy^2 over all partners, sums the lateral
//     weights first (2 FMAs per pair instead of 2 per partner) and evaluates
//     1/(sqrt(s)+1e-18)^3 with one MUFU.RSQ64H seed + 7 FP64 instructions
//     (rb2_inv_r3_soft) instead of sqrt + divide: 74 FP64-pipe instructions + 6 MUFU per
//     ordered pair at N_ic_max = 1 against 99 algorithmic flops;
//   * the reference's index-ordered image roles (j>i evaluates at (z_i,z_j), j<i at
//     (z_j,z_i)) reduce to a sign on the same-charge z-sum, resolved per tile except in
//     the tiles that overlap the CTA's own i-range.
#include "rb2_internal.cuh"
#include "rb2_planar_math.cuh"
#include "rb2_tip_math.cuh"

namespace {

#ifndef RB2_BLOCK
#define RB2_BLOCK 128
#endif
#ifndef RB2_MINB
#define RB2_MINB 4
#endif
#ifndef RB2_UNROLL
#define RB2_UNROLL 4
#endif
#ifndef RB2_TJ
#define RB2_TJ 128
#endif
constexpr int TJ = RB2_TJ;        // j-particles per shared-memory tile (32 B each)
constexpr int STAGES = 3;         // ring depth
constexpr int UNROLL = RB2_UNROLL;
constexpr int BLOCK = RB2_BLOCK;  // i-particles (or field points) per CTA, one per thread

// ---- TMA bulk copy + mbarrier (sm_90+ PTX, UBLKCP / SYNCS in SASS) -----------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "RB2_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra RB2_DONE;\n"
        "bra RB2_WAIT;\n"
        "RB2_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Slow path of a (target, source tile) whose fast sweep flagged a laterally close pair (rb2_is_close): the whole
// tile again with the reference's sqrt / divide.  i_self: global index of the target when the tile may hold it or
// lower-indexed particles (per-element self mask and role sign, j = j0 + jj); below -cnt: plain sources.
template <int NIC>
__device__ __noinline__ Acc4 planar_tile_exact(double xi, double yi, double zi, const double4 *tile, int cnt, int j0, int i_self,
                                               PlanarParams P)  // i_self < 0: no roles; INT_MAX: every source has the lower index
{
    Acc4 a = {0.0, 0.0, 0.0, 0.0};
    const bool roles = i_self >= 0;
    for (int jj = 0; jj < cnt; ++jj) {
        const double4 pj = tile[jj];
        const int j = j0 + jj;
        const double qe = (roles && j == i_self) ? 0.0 : pj.w;
        const bool swapped = roles && j < i_self;
        planar_term_exact<NIC>(xi, yi, zi, pj, qe, swapped ? -qe : qe, P, a, swapped);
    }
    return a;
}

// Slow path of the tip tiles: the literal arithmetic (IEEE sqrt / divide, the source image recomputed per element).
template <int NIC, bool FIELD>
__device__ __noinline__ Acc4 tip_tile_exact(TipParams T, double xi, double yi, double zi, const double4 *tile, int cnt, int j0, int i)
{
    Acc4 a = {0.0, 0.0, 0.0, 0.0};
    TipImage im_i{};
    if (NIC >= 0) im_i = tip_image_point(T, xi, yi, zi);
    for (int jj = 0; jj < cnt; ++jj) {
        const double4 pj = tile[jj];
        const int j = j0 + jj;
        const double qe = (!FIELD && (j == i)) ? 0.0 : pj.w;
        // Coulomb, src/mod_verlet.F90:1371-1379
        const double dx = xi - pj.x, dy = yi - pj.y, dz = zi - pj.z;
        const double r = sqrt(dx * dx + dy * dy + dz * dz) + rb2k::soft;
        const double inv_r3 = 1.0 / (r * r * r);
        double fx = inv_r3 * dx, fy = inv_r3 * dy, fz = inv_r3 * dz;
        if (NIC >= 0 && !(!FIELD && j == i)) {
            double ic_x, ic_y, ic_z;
            if (FIELD || (j > i)) {
                // a = target (the field point, or the lower-indexed particle i), b = source j
                tip_ic_force(T, im_i, xi, yi, zi, pj.x, pj.y, pj.z, ic_x, ic_y, ic_z);
                fx += ic_x; fy += ic_y; fz += ic_z;
            } else {
                // a = particle j (lower index), b = particle i; x, y mirrored (sgn_xy = -1)
                const TipImage im_j = tip_image_point(T, pj.x, pj.y, pj.z);
                tip_ic_force(T, im_j, pj.x, pj.y, pj.z, xi, yi, zi, ic_x, ic_y, ic_z);
                fx -= ic_x; fy -= ic_y; fz += ic_z;
            }
        }
        a.x = fma(qe, fx, a.x);
        a.y = fma(qe, fy, a.y);
        a.z = fma(qe, fz, a.z);
    }
    return a;
}

// ---- the tiled pair kernel --------------------------------------------------------------------
// GEOM 1 planar / 2 tip.  NIC: -1 image charge off, 0, 1, 2 (= runtime N_ic_max loop); tip uses
// NIC -1 / 0 for image charge off / on.  FIELD: targets are the M field points (no self
// term, no index roles); otherwise targets are particles i_begin..i_end-1 of pq itself.
// partial[(slot*3 + c)*n_tgt + (i - i_begin)] receives  sum_j q_j * (Coulomb + image)_c  over
// this CTA's j-chunk.
template <int GEOM, int NIC, bool FIELD>
__global__ void __launch_bounds__(BLOCK, RB2_MINB)
k_pair(const double4 *__restrict__ src, int n_src, const double4 *__restrict__ tgt_pq, const double *__restrict__ tgt_pts,
       int i_begin, int i_end, int j_chunk, int slot0, PlanarParams P, TipParams T, double *__restrict__ partial,
       const double4 *__restrict__ src_img)  // tip accelerations: the sources' sphere images (tip_image_packed)
{
    __shared__ __align__(128) double4 tiles[STAGES * TJ];
    __shared__ __align__(8) uint64_t full[STAGES];

    const int tid = threadIdx.x;
    const int ib = i_begin + blockIdx.x * BLOCK;  // first target of this CTA
    const int ie = min(ib + BLOCK, i_end);
    const int i = ib + tid;
    const bool active = i < i_end;
    const bool warp_active = (ib + (tid & ~31)) < i_end;
    const int ii = active ? i : (i_end - 1);

    double xi, yi, zi;
    if (FIELD) {
        xi = tgt_pts[3 * ii];
        yi = tgt_pts[3 * ii + 1];
        zi = tgt_pts[3 * ii + 2];
    } else {
        const double4 pi = tgt_pq[ii];
        xi = pi.x; yi = pi.y; zi = pi.z;
    }

    // idle lanes: metres away from every source (never "laterally close"); planar: at a height inside the gap, so that
    // they take the same branch as a field point between the plates (no divergence in a partly filled warp)
    if (!active) { xi = 1.0 + (double)tid; yi = 0.0; zi = (GEOM == 1) ? 0.25 * P.two_d : 1.0; }

    const int j_begin = blockIdx.y * j_chunk;
    const int j_end = min(n_src, j_begin + j_chunk);
    const int ntiles = (j_end - j_begin + TJ - 1) / TJ;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int t) {
        if (t < ntiles) {
            const int j0 = j_begin + t * TJ;
            const int cnt = min(TJ, j_end - j0);
            const int s = t % STAGES;
            mbar_expect_tx(&full[s], (uint32_t)cnt * 32u);
            bulk_g2s(&tiles[s * TJ], src + j0, (uint32_t)cnt * 32u, &full[s]);
        }
    };
    if (tid == 0) {
        for (int t = 0; t < STAGES - 1; ++t) issue(t);
    }

    double4 img_i = make_double4(0.0, 0.0, 0.0, 0.0);
    if (GEOM == 2 && NIC >= 0) img_i = tip_image_packed(T, xi, yi, zi);

    double ax = 0.0, ay = 0.0, az = 0.0;
    for (int t = 0; t < ntiles; ++t) {
        if (tid == 0) issue(t + STAGES - 1);  // refills the stage released by the barrier of iteration t-1
        const int s = t % STAGES;
        mbar_wait(&full[s], (uint32_t)((t / STAGES) & 1));
        const int j0 = j_begin + t * TJ;
        const int cnt = min(TJ, j_end - j0);
        const double4 *tile = &tiles[s * TJ];

        if (warp_active) {
            Acc4 a = {0.0, 0.0, 0.0, 0.0};
            if (GEOM == 1) {
                bool close = false;
                if (!FIELD && (j0 < ie) && (j0 + cnt > ib)) {
                    // tile overlaps this CTA's own particles: per-element self mask and role sign
                    for (int jj = 0; jj < cnt; ++jj) {
                        const double4 pj = tile[jj];
                        const int j = j0 + jj;
                        const double qe = (j == i) ? 0.0 : pj.w;
                        const double qs = (j > i) ? qe : -qe;
                        bool c1 = false;  // the self pair (offset 0) is not a close pair: it is masked out by qe = 0
                        planar_term<NIC>(xi, yi, zi, pj, qe, qs, P, a, c1);
                        close = close || (c1 && j != i);
                    }
                    if (close) a = planar_tile_exact<NIC>(xi, yi, zi, tile, cnt, j0, i, P);
                    az += a.t;
                } else {
                    // particles between the plates, d >= 1 um: the three far image partners without the softening term;
                    // field points qualify one by one (a point's value must not depend on its neighbours in the batch)
                    if (NIC == 1 && P.far_ok && (!FIELD || (zi >= 0.0 && zi <= 0.5 * P.two_d))) {
#pragma unroll UNROLL
                        for (int jj = 0; jj < cnt; ++jj) {
                            const double4 pj = tile[jj];
                            planar_term<NIC, true>(xi, yi, zi, pj, pj.w, pj.w, P, a, close);
                        }
                    } else {
#pragma unroll UNROLL
                        for (int jj = 0; jj < cnt; ++jj) {
                            const double4 pj = tile[jj];
                            planar_term<NIC>(xi, yi, zi, pj, pj.w, pj.w, P, a, close);
                        }
                    }
                    if (close) {
                        // sources all above (or field points: no roles) / all below this CTA's rows; the slow path
                        // resolves the roles per element, so its a.t is already signed
                        const bool above = FIELD || (j0 >= ie);
                        a = planar_tile_exact<NIC>(xi, yi, zi, tile, cnt, j0, above ? -1 : 0x7fffffff, P);
                        az += a.t;
                    } else {
                        const double sg = (FIELD || (j0 >= ie)) ? 1.0 : -1.0;
                        az = fma(sg, a.t, az);
                    }
                }
            } else {
                // tip: the fast pair forms of rb2_tip_math.cuh; a pair closer than 1e-11 m sends this thread's row of the
                // tile to the literal arithmetic
                bool close = false;
                for (int jj = 0; jj < cnt; ++jj) {
                    const double4 pj = tile[jj];
                    const int j = j0 + jj;
                    const bool self = !FIELD && (j == i);
                    const double qe = self ? 0.0 : pj.w;
                    double fx, fy, fz;
                    bool c1 = false;
                    if (FIELD || j >= i) tip_pair_fast_upper(img_i, NIC >= 0, xi, yi, zi, pj, fx, fy, fz, c1);
                    else tip_pair_fast_lower(NIC >= 0 ? src_img[j] : make_double4(0.0, 0.0, 0.0, 0.0), NIC >= 0, xi, yi, zi, pj, fx, fy, fz, c1);
                    close = close || (c1 && !self);
                    a.x = fma(qe, fx, a.x);
                    a.y = fma(qe, fy, a.y);
                    a.z = fma(qe, fz, a.z);
                }
                if (close) a = tip_tile_exact<NIC, FIELD>(T, xi, yi, zi, tile, cnt, j0, i);
            }
            ax += a.x; ay += a.y; az += a.z;
        }
        __syncthreads();  // every warp is done with stage s before it is refilled
    }

    if (active) {
        const int n_tgt = i_end - i_begin;
        const size_t base = (size_t)(slot0 + blockIdx.y) * 3 * (size_t)n_tgt + (size_t)(i - i_begin);
        partial[base] = ax;
        partial[base + (size_t)n_tgt] = ay;
        partial[base + 2 * (size_t)n_tgt] = az;
    }
}

// a_i = ( q_i/(4 pi eps0) * sum + q_i * E_vac(r_i) ) / m_i   (src/mod_verlet.F90:1333-1338, :1416-1424)
__global__ void k_accel_finalize(const double *__restrict__ partial, int nslots, int n_tgt, int i_begin,
                                 const double4 *__restrict__ pq, const double *__restrict__ mass, int geometry,
                                 PlanarParams P, TipParams T, double *__restrict__ acc)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_tgt) return;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int s = 0; s < nslots; ++s) {
        const size_t base = (size_t)s * 3 * (size_t)n_tgt + (size_t)k;
        sx += partial[base];
        sy += partial[base + (size_t)n_tgt];
        sz += partial[base + 2 * (size_t)n_tgt];
    }
    const int i = i_begin + k;
    const double4 pi = pq[i];
    const double q_1 = pi.w;
    const double qd_1 = q_1 * rb2k::div_fac_c;
    const double im_1 = 1.0 / mass[i];
    if (geometry == RB2_GEOM_PLANAR) {
        acc[3 * i] = (qd_1 * sx) * im_1;
        acc[3 * i + 1] = (qd_1 * sy) * im_1;
        acc[3 * i + 2] = (qd_1 * sz + q_1 * P.E_z) * im_1;
    } else {
        double fE_x, fE_y, fE_z;
        rb2_tip_field_E(T, pi.x, pi.y, pi.z, fE_x, fE_y, fE_z);
        acc[3 * i] = (qd_1 * sx + q_1 * fE_x) * im_1;
        acc[3 * i + 1] = (qd_1 * sy + q_1 * fE_y) * im_1;
        acc[3 * i + 2] = (qd_1 * sz + q_1 * fE_z) * im_1;
    }
}

// E(p) = E_vac(p) + 1/(4 pi eps0) * sum   (src/mod_verlet.F90:1737-1743 / :1808-1820 seed + accumulate)
__global__ void k_field_finalize(const double *__restrict__ partial, int nslots, int M, const double *__restrict__ pts,
                                 int geometry, PlanarParams P, TipParams T, double *__restrict__ fld)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int s = 0; s < nslots; ++s) {
        const size_t base = (size_t)s * 3 * (size_t)M + (size_t)k;
        sx += partial[base];
        sy += partial[base + (size_t)M];
        sz += partial[base + 2 * (size_t)M];
    }
    double fE_x = 0.0, fE_y = 0.0, fE_z = P.E_z;
    if (geometry == RB2_GEOM_TIP) rb2_tip_field_E(T, pts[3 * k], pts[3 * k + 1], pts[3 * k + 2], fE_x, fE_y, fE_z);
    fld[3 * k] = fE_x + rb2k::div_fac_c * sx;
    fld[3 * k + 1] = fE_y + rb2k::div_fac_c * sy;
    fld[3 * k + 2] = fE_z + rb2k::div_fac_c * sz;
}

// Few points against many j-chunks (the samplers' small batches): one CTA per point, strided slot sums
// and a fixed-shape tree, so the result does not depend on scheduling.
__global__ void __launch_bounds__(128) k_field_finalize_wide(const double *__restrict__ partial, int nslots, int M,
                                                             const double *__restrict__ pts, int geometry, PlanarParams P,
                                                             TipParams T, double *__restrict__ fld)
{
    __shared__ double sh[3][128];
    const int k = blockIdx.x, t = threadIdx.x;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int s = t; s < nslots; s += 128) {
        const size_t base = (size_t)s * 3 * (size_t)M + (size_t)k;
        sx += partial[base];
        sy += partial[base + (size_t)M];
        sz += partial[base + 2 * (size_t)M];
    }
    sh[0][t] = sx; sh[1][t] = sy; sh[2][t] = sz;
    __syncthreads();
    for (int w = 64; w > 0; w >>= 1) {
        if (t < w) { sh[0][t] += sh[0][t + w]; sh[1][t] += sh[1][t + w]; sh[2][t] += sh[2][t + w]; }
        __syncthreads();
    }
    if (t == 0) {
        double fE_x = 0.0, fE_y = 0.0, fE_z = P.E_z;
        if (geometry == RB2_GEOM_TIP) rb2_tip_field_E(T, pts[3 * k], pts[3 * k + 1], pts[3 * k + 2], fE_x, fE_y, fE_z);
        fld[3 * k] = fE_x + rb2k::div_fac_c * sh[0][0];
        fld[3 * k + 1] = fE_y + rb2k::div_fac_c * sh[1][0];
        fld[3 * k + 2] = fE_z + rb2k::div_fac_c * sh[2][0];
    }
}

__global__ void k_tip_images(const double4 *__restrict__ pq, int n, TipParams T, double4 *__restrict__ img)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) { const double4 p = pq[j]; img[j] = tip_image_packed(T, p.x, p.y, p.z); }
}

struct Split {
    int nblk, nsplit, j_chunk;
};

// Equal-cost CTAs: (target blocks) x (j-chunks).  Enough chunks for >= 8 waves at 4 CTAs/SM,
// each chunk a whole number of tiles.
Split choose_split(int n_tgt, int n_src, int sm_count)
{
    Split s;
    s.nblk = (n_tgt + BLOCK - 1) / BLOCK;
    const int want_units = sm_count * RB2_MINB * 8;
    int ns = (want_units + s.nblk - 1) / s.nblk;
    // chunks are whole tiles -- unless that leaves less than one wave of CTAs (a thousand particles: 8 x 8 CTAs, each
    // thread a serial chain over 128 sources; the tip deck spent 120 us per step there): then quarter tiles
    int gran = TJ;
    if ((long long)s.nblk * ((n_src + TJ - 1) / TJ) < (long long)sm_count * RB2_MINB) gran = 32;
    const int max_ns = (n_src + gran - 1) / gran;
    if (ns > max_ns) ns = max_ns;
    if (ns < 1) ns = 1;
    int chunk = (n_src + ns - 1) / ns;
    chunk = ((chunk + gran - 1) / gran) * gran;
    s.j_chunk = chunk;
    s.nsplit = (n_src + chunk - 1) / chunk;
    if (s.nsplit > 65535) {
        s.nsplit = 65535;
        chunk = (n_src + s.nsplit - 1) / s.nsplit;
        s.j_chunk = ((chunk + TJ - 1) / TJ) * TJ;
        s.nsplit = (n_src + s.j_chunk - 1) / s.j_chunk;
    }
    return s;
}

int ensure_partial(Rb2Ctx &ctx, size_t bytes)
{
    if (bytes > ctx.partial_bytes) {
        if (ctx.partial) RB2_CUDA(cudaFree(ctx.partial));
        ctx.partial = nullptr;
        ctx.partial_bytes = 0;
        size_t want = bytes + bytes / 4;
        RB2_CUDA(cudaMalloc(&ctx.partial, want));
        ctx.partial_bytes = want;
    }
    return RB2_OK;
}

template <bool FIELD>
int launch_pair(Rb2Ctx &ctx, const double4 *src, int n_src, const double4 *tgt_pq, const double *tgt_pts, int i_begin,
                int i_end, const Split &sp, int slot0, double *partial)
{
    const double4 *src_img = nullptr;
    if (!FIELD && ctx.cfg.geometry == RB2_GEOM_TIP && ctx.cfg.image_charge) {
        // the sphere image of every particle, once per evaluation (the pairs with a lower-indexed source need the SOURCE's)
        if ((size_t)n_src > ctx.tip_img_cap) {
            if (ctx.d_tip_img) RB2_CUDA(cudaFree(ctx.d_tip_img));
            ctx.d_tip_img = nullptr; ctx.tip_img_cap = 0;
            const size_t want = std::max((size_t)ctx.cap, (size_t)n_src);
            RB2_CUDA(cudaMalloc(&ctx.d_tip_img, want * sizeof(double4)));
            ctx.tip_img_cap = want;
        }
        k_tip_images<<<(n_src + 255) / 256, 256, 0, ctx.stream>>>(src, n_src, rb2_make_step_params(ctx.cfg).tip, ctx.d_tip_img);
        RB2_CUDA(cudaGetLastError());
        RB2_LAUNCHED(1);
        src_img = ctx.d_tip_img;
    }
    const rb2_config &c = ctx.cfg;
    StepParams P = rb2_make_step_params(c);
    P.pl.far_ok = P.pl.far_ok && rb2_far_allowed(ctx);  // option "sym_far" (on by default), sources between the plates
    dim3 grid(sp.nblk, sp.nsplit), block(BLOCK);
    cudaStream_t st = ctx.stream;
#define RB2_GO(G, N) k_pair<G, N, FIELD><<<grid, block, 0, st>>>(src, n_src, tgt_pq, tgt_pts, i_begin, i_end, sp.j_chunk, slot0, P.pl, P.tip, partial, src_img)
    if (c.geometry == RB2_GEOM_PLANAR) {
        if (!c.image_charge) RB2_GO(1, -1);
        else if (c.N_ic_max == 0) RB2_GO(1, 0);
        else if (c.N_ic_max == 1) RB2_GO(1, 1);
        else RB2_GO(1, 2);
    } else if (c.geometry == RB2_GEOM_TIP) {
        if (!c.image_charge) RB2_GO(2, -1);
        else RB2_GO(2, 0);
    } else {
        return rb2_fail(RB2_ERR_GEOMETRY, "geometry %d is not implemented on the device (ACC_GEOM_OTHER)", c.geometry);
    }
#undef RB2_GO
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(1);
    return RB2_OK;
}

}  // namespace

// Calculate_Acceleration_Particles for targets i_begin..i_end-1 (overwrites acc(3,i)).
int rb2_launch_accel(Rb2Ctx &ctx, const double4 *pq, const double *mass, int n, int i_begin, int i_end, double *acc_out)
{
    if (n < 1 || i_end <= i_begin) return RB2_OK;
    const int n_tgt = i_end - i_begin;
    const Split sp = choose_split(n_tgt, n, ctx.sm_count);
    int rc = ensure_partial(ctx, (size_t)sp.nsplit * 3 * (size_t)n_tgt * sizeof(double));
    if (rc != RB2_OK) return rc;
    RB2_CUDA(rb2_event_record(ctx, ctx.ev_a0));
    rc = launch_pair<false>(ctx, pq, n, pq, nullptr, i_begin, i_end, sp, 0, ctx.partial);
    if (rc != RB2_OK) return rc;
    const StepParams P = rb2_make_step_params(ctx.cfg);
    k_accel_finalize<<<(n_tgt + 255) / 256, 256, 0, ctx.stream>>>(ctx.partial, sp.nsplit, n_tgt, i_begin, pq, mass,
                                                                  ctx.cfg.geometry, P.pl, P.tip, acc_out);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(1);
    RB2_CUDA(rb2_event_record(ctx, ctx.ev_a1));
    ctx.last_grid_x = sp.nblk; ctx.last_grid_y = sp.nsplit; ctx.last_block = BLOCK; ctx.last_split = sp.j_chunk;
    return RB2_OK;
}

// Calc_Field_at_Batch: M points in d_pts (3,M) -> d_fld (3,M); sources pq[0..n) plus extra[0..n_extra).
// Tip geometry, few points against few particles (the tip sampler: ~200 proposals x ~1e3 electrons per jump): the tiled
// kernel above has one thread per point and at most n / 128 source chunks -- 16 CTAs of 128 sequential heavy (IEEE sqrt /
// divide) pair evaluations each, ~90 us.  Here a CTA owns ONE point, its 128 threads stride over the sources, and the
// sums are joined in a fixed order (shuffle tree, then the four warps): M CTAs, a few us.  Same arithmetic per pair.
__global__ void __launch_bounds__(128) k_tip_field_point(const double4 *__restrict__ src, int n, const double *__restrict__ pts,
                                                          int do_ic, TipParams T, double *__restrict__ fld)
{
    __shared__ double sh[3][4];
    const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double xi = pts[3 * k], yi = pts[3 * k + 1], zi = pts[3 * k + 2];
    TipImage im_i{};
    double4 img_i = make_double4(0.0, 0.0, 0.0, 0.0);
    if (do_ic) { im_i = tip_image_point(T, xi, yi, zi); img_i = tip_image_packed(T, xi, yi, zi); }
    double ax = 0.0, ay = 0.0, az = 0.0;
    for (int j = tid; j < n; j += 128) {
        const double4 pj = src[j];
        double fx, fy, fz;
        bool close = false;
        // Coulomb (src/mod_verlet.F90:1511-1518) + Sphere_IC_field(p, r_j) (:1520, the field point is imaged), fast form;
        // a source within 1e-11 m: the literal arithmetic
        tip_pair_fast_upper(img_i, do_ic != 0, xi, yi, zi, pj, fx, fy, fz, close);
        if (close) tip_point_field(T, im_i, do_ic != 0, xi, yi, zi, pj, fx, fy, fz);
        ax = fma(pj.w, fx, ax);
        ay = fma(pj.w, fy, ay);
        az = fma(pj.w, fz, az);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ax += __shfl_xor_sync(0xffffffffu, ax, o);
        ay += __shfl_xor_sync(0xffffffffu, ay, o);
        az += __shfl_xor_sync(0xffffffffu, az, o);
    }
    if (lane == 0) { sh[0][warp] = ax; sh[1][warp] = ay; sh[2][warp] = az; }
    __syncthreads();
    if (tid == 0) {
        const double sx = ((sh[0][0] + sh[0][1]) + sh[0][2]) + sh[0][3];
        const double sy = ((sh[1][0] + sh[1][1]) + sh[1][2]) + sh[1][3];
        const double sz = ((sh[2][0] + sh[2][1]) + sh[2][2]) + sh[2][3];
        double fE_x, fE_y, fE_z;
        rb2_tip_field_E(T, xi, yi, zi, fE_x, fE_y, fE_z);
        fld[3 * k] = fE_x + rb2k::div_fac_c * sx;
        fld[3 * k + 1] = fE_y + rb2k::div_fac_c * sy;
        fld[3 * k + 2] = fE_z + rb2k::div_fac_c * sz;
    }
}

int rb2_launch_field(Rb2Ctx &ctx, const double4 *pq, int n, const double4 *extra, int n_extra, const double *d_pts, int M,
                     double *d_fld)
{
    if (M < 1) return RB2_OK;
    if (ctx.cfg.geometry == RB2_GEOM_TIP && ctx.tip_field_small && n_extra == 0 && n >= 1 && n <= 8192 && M <= 2048) {
        const StepParams P = rb2_make_step_params(ctx.cfg);
        k_tip_field_point<<<M, 128, 0, ctx.stream>>>(pq, n, d_pts, P.tip.do_ic, P.tip, d_fld);
        RB2_CUDA(cudaGetLastError());
        RB2_LAUNCHED(1);
        return RB2_OK;
    }
    Split sp{}, spx{};
    int nslots = 0;
    if (n > 0) { sp = choose_split(M, n, ctx.sm_count); nslots += sp.nsplit; }
    if (n_extra > 0) { spx = choose_split(M, n_extra, ctx.sm_count); nslots += spx.nsplit; }
    if (nslots > 0) {
        int rc = ensure_partial(ctx, (size_t)nslots * 3 * (size_t)M * sizeof(double));
        if (rc != RB2_OK) return rc;
        if (n > 0) {
            rc = launch_pair<true>(ctx, pq, n, nullptr, d_pts, 0, M, sp, 0, ctx.partial);
            if (rc != RB2_OK) return rc;
        }
        if (n_extra > 0) {
            rc = launch_pair<true>(ctx, extra, n_extra, nullptr, d_pts, 0, M, spx, sp.nsplit, ctx.partial);
            if (rc != RB2_OK) return rc;
        }
    }
    const StepParams P = rb2_make_step_params(ctx.cfg);
    if (ctx.cfg.geometry != RB2_GEOM_PLANAR && ctx.cfg.geometry != RB2_GEOM_TIP)
        return rb2_fail(RB2_ERR_GEOMETRY, "geometry %d is not implemented on the device", ctx.cfg.geometry);
    if (nslots >= 32)
        k_field_finalize_wide<<<M, 128, 0, ctx.stream>>>(ctx.partial, nslots, M, d_pts, ctx.cfg.geometry, P.pl, P.tip, d_fld);
    else
        k_field_finalize<<<(M + 127) / 128, 128, 0, ctx.stream>>>(ctx.partial, nslots, M, d_pts, ctx.cfg.geometry, P.pl, P.tip, d_fld);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(1);
    return RB2_OK;
}

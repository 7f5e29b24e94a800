// rb2_nearest.cu -- Sample_Elec_Position (src/mod_pair.F90:975-1037): for every electron the distance to the
// nearest other electron and that electron's index.  The reference documents this O(N^2) sweep as a 4x cost when it
// runs every step (:981-984); it has the access pattern of the pair kernel with a min-reduction instead of a sum.
//
// One thread per target row, 256 per CTA; the sources stream through shared memory in tiles of 256 {x, y, z,
// electron flag}; the j-range is split into chunks so that small systems still fill the machine, and a second
// kernel joins the chunk minima in ascending chunk order.  Bit-exact with the serial scan of the reference:
//   * dist = sqrt(dx*dx + dy*dy + dz*dz) with every product and sum rounded separately (no FMA contraction) and
//     an IEEE square root -- but the root is only taken when the squared distance beats the running one, which
//     is rare after the first few sources (sqrt is monotone, so a row that loses on d^2 loses on d);
//   * strict `<` over ascending j inside a chunk and over ascending chunks: the lowest index wins a tie, also
//     when two different squared distances round to the same distance;
//   * the species test only (electrons already marked for removal still count, like in the reference).
#include "rb2_internal.cuh"

namespace {

constexpr int NB = 256;

struct Best {
    double d2, dist;
    int id;
};

__device__ __forceinline__ void consider(Best &b, double d2, int j)
{
    if (d2 < b.d2) {
        const double dist = __dsqrt_rn(d2);
        if (dist < b.dist) { b.d2 = d2; b.dist = dist; b.id = j; }
    }
}

__global__ void __launch_bounds__(NB)
k_nearest(int n, const double4 *__restrict__ pq, const int *__restrict__ species, int j_chunk, double *__restrict__ pdist,
          int *__restrict__ pid)
{
    __shared__ double xs[NB], ys[NB], zs[NB];
    __shared__ int es[NB];
    const int i = blockIdx.x * NB + threadIdx.x;
    const int j0 = blockIdx.y * j_chunk, j1 = min(n, j0 + j_chunk);
    double xi = 0.0, yi = 0.0, zi = 0.0;
    bool mine = false;
    if (i < n) {
        const double4 p = pq[i];
        xi = p.x; yi = p.y; zi = p.z;
        mine = species[i] == RB2_SPECIES_ELEC;
    }
    Best b{1.0e6, 1000.0, -1};  // particles_nearest_dist = 1000.0d0 (:988); 1000^2 is exact
    for (int t0 = j0; t0 < j1; t0 += NB) {
        const int j = t0 + threadIdx.x;
        __syncthreads();
        if (j < j1) {
            const double4 p = pq[j];
            xs[threadIdx.x] = p.x; ys[threadIdx.x] = p.y; zs[threadIdx.x] = p.z;
            es[threadIdx.x] = species[j] == RB2_SPECIES_ELEC;
        } else {
            es[threadIdx.x] = 0;
        }
        __syncthreads();
        if (!mine) continue;
        const int cnt = min(NB, j1 - t0);
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
            const double dx = __dsub_rn(xi, xs[k]), dy = __dsub_rn(yi, ys[k]), dz = __dsub_rn(zi, zs[k]);
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            if (es[k] && (t0 + k) != i) consider(b, d2, t0 + k);
        }
    }
    if (i < n) {
        pdist[(size_t)blockIdx.y * n + i] = b.dist;
        pid[(size_t)blockIdx.y * n + i] = b.id;
    }
}

__global__ void k_nearest_join(int n, int nchunks, const double *__restrict__ pdist, const int *__restrict__ pid,
                               double *__restrict__ dist, int *__restrict__ id)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double best = 1000.0;
    int bid = -1;
    for (int c = 0; c < nchunks; ++c) {
        const double d = pdist[(size_t)c * n + i];
        if (d < best) { best = d; bid = pid[(size_t)c * n + i]; }
    }
    dist[i] = best;
    id[i] = bid;
}

}  // namespace

int rb2_launch_nearest(Rb2Ctx &ctx, double *d_dist, int *d_id)
{
    const int n = ctx.n;
    if (n < 1) return RB2_OK;
    const int iblocks = (n + NB - 1) / NB;
    // enough CTAs for 8 per SM; chunks are whole tiles
    int nchunks = (8 * ctx.sm_count + iblocks - 1) / iblocks;
    const int max_chunks = (n + NB - 1) / NB;
    if (nchunks > max_chunks) nchunks = max_chunks;
    if (nchunks > 64) nchunks = 64;
    if (nchunks < 1) nchunks = 1;
    int j_chunk = (n + nchunks - 1) / nchunks;
    j_chunk = ((j_chunk + NB - 1) / NB) * NB;
    nchunks = (n + j_chunk - 1) / j_chunk;
    // partial results: nchunks * n doubles + nchunks * n ints in the staging buffers, behind the outputs
    int rc = rb2_ensure_stage(ctx, (size_t)(nchunks + 1) * n, (size_t)(nchunks + 1) * n);
    if (rc) return rc;
    double *pdist = ctx.d_stage_d + n;
    int *pid = ctx.d_stage_i + n;
    if (!d_dist) d_dist = ctx.d_stage_d;
    if (!d_id) d_id = ctx.d_stage_i;
    k_nearest<<<dim3(iblocks, nchunks), NB, 0, ctx.stream>>>(n, ctx.a.pq, ctx.a.species, j_chunk, pdist, pid);
    RB2_CUDA(cudaGetLastError());
    k_nearest_join<<<(n + 255) / 256, 256, 0, ctx.stream>>>(n, nchunks, pdist, pid, d_dist, d_id);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(2);
    return RB2_OK;
}

extern "C" int rb2_nearest_electron(double *dist_out, int *id_out)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    const int n = c.n;
    if (n < 1) return RB2_OK;
    int rc = rb2_launch_nearest(c, nullptr, nullptr);
    if (rc) return rc;
    if (dist_out) RB2_CUDA(cudaMemcpyAsync(dist_out, c.d_stage_d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    if (id_out) RB2_CUDA(cudaMemcpyAsync(id_out, c.d_stage_i, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    RB2_CUDA(cudaStreamSynchronize(c.stream));
    return RB2_OK;
}

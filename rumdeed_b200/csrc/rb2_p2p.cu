// rb2_p2p.cu -- the one exchange step of the multi-GPU path over NVLink peer memory, fused with the
// finalisation of the accelerations.
//
// With the pair work split over `world` processes (one per GPU, rb2_set_pair_rank) every rank holds partial
// sums raw_r[3][n_pad] of the forces on ALL particles.  Instead of handing that buffer to an NCCL all-reduce
// and finalising afterwards, every rank maps the exchange block of every peer (CUDA IPC) and ONE kernel
//   - waits until every peer has published its partial sums of this evaluation (flags written straight into
//     this GPU's memory by the peers' signal kernels),
//   - reads raw_r[i] of every rank r through the peer mappings (NVLink loads), adds them in rank order --
//     the same order on every rank, so the replicated O(N) state stays bit-identical everywhere --
//   - applies q_i/(4 pi eps0), the vacuum field and 1/m_i (src/mod_verlet.F90:1333-1338) and writes acc.
// The partial-sum buffer is double buffered by evaluation parity, which makes one flag round per evaluation
// enough: a rank overwrites buffer k&1 in evaluation k+2, after its own wait of evaluation k+1, and a peer
// signals k+1 only behind its reads of evaluation k in stream order.
//
// Exchange block (one cudaMalloc, exported with cudaIpcGetMemHandle):
//   [ 4 KB: flags, one u64 per peer; header at byte 2048 ] [ raw parity 0: 3 * npad_max doubles ] [ raw parity 1: the same ]
// A flag holds (evaluation number << 32 | padded particle count of that evaluation).  The header {magic, npad_max} is
// checked when the peers are attached (all ranks must have exported for the same n_max: the slot offsets depend on
// it); the padded count is checked against the reader's own in the finalise kernel.  A peer that does not arrive within
// 20 s, or arrives with another padded count (the ranks did not make the same sequence of evaluations on the same
// particle store), is REPORTED: the kernel sets an error word in mapped host memory and returns, the call that waited
// for it fails with RB2_ERR_CUDA -- the context survives.
//
// Every pair-symmetric evaluation on an attached context is COLLECTIVE: rb2_step, rb2_accel_only, rb2_accel_host and
// rb2_accel_partial / rb2_accel_finalize must be called by all ranks in the same order with the same particle count.
#include "rb2_internal.cuh"

namespace {

constexpr size_t P2P_FLAG_BYTES = 4096;
constexpr size_t P2P_HEADER_OFFSET = 2048;
constexpr unsigned long long P2P_MAGIC = 0x7262325F70327031ull;  // "rb2_p2p1"
struct P2PHeader { unsigned long long magic, npad_max; };
enum { P2P_ERR_TIMEOUT = 1, P2P_ERR_MISMATCH = 2 };
constexpr unsigned long long P2P_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;  // a peer that never signals: trap

struct P2PPeers {
    const double *raw[RB2_P2P_MAX];          // this evaluation's partial sums of rank r
    unsigned long long *flags[RB2_P2P_MAX];  // flag array of rank r
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double2 ld_peer2(const double *p)
{
    double2 v;  // read at system scope: the line lives in the peer's L2, never in a stale local L1
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Runs behind this rank's partial-sum kernels in stream order: tell every peer (and myself) that the partial sums
// of evaluation `epoch` are complete.
__global__ void k_p2p_signal(P2PPeers pp, int rank, int world, unsigned long long epoch, int n_pad)
{
    const int p = threadIdx.x;
    if (p < world) {
        __threadfence_system();
        st_release_sys(pp.flags[p] + rank, (epoch << 32) | (unsigned long long)(unsigned)n_pad);
    }
}

// Two particles per thread (n_pad is even, the rows are 16-byte aligned).
__global__ void __launch_bounds__(256)
k_sym_finalize_p2p(int n, int n_pad, P2PPeers pp, int rank, int world, unsigned long long epoch,
                   const double4 *__restrict__ pq, const double *__restrict__ mass, PlanarParams P, double *__restrict__ acc,
                   volatile int *__restrict__ err)
{
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    if (threadIdx.x < world) {
        const unsigned long long *f = pp.flags[rank] + threadIdx.x;  // local memory, written by peer threadIdx.x
        const unsigned long long t0 = globaltimer_ns();
        unsigned long long v;
        while (((v = ld_acquire_sys(f)) >> 32) < epoch) {
            if (*err != 0) { bad = 1; break; }  // another CTA has already given up
            __nanosleep(200);
            if (globaltimer_ns() - t0 > P2P_TIMEOUT_NS) { *err = P2P_ERR_TIMEOUT; bad = 1; break; }
        }
        if (!bad && ((v >> 32) != epoch || (int)(v & 0xffffffffu) != n_pad)) { *err = P2P_ERR_MISMATCH; bad = 1; }
    }
    __syncthreads();
    if (bad) return;  // reported by the host side of the call (rb2_p2p_check)
    const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n) return;
    double2 s0 = make_double2(0.0, 0.0), s1 = s0, s2 = s0;
    for (int r = 0; r < world; ++r) {
        const double *raw = pp.raw[r];
        const double2 a = ld_peer2(raw + i), b = ld_peer2(raw + (size_t)n_pad + i), c = ld_peer2(raw + 2 * (size_t)n_pad + i);
        s0.x += a.x; s0.y += a.y; s1.x += b.x; s1.y += b.y; s2.x += c.x; s2.y += c.y;
    }
    {
        const double q_1 = pq[i].w, qd_1 = q_1 * rb2k::div_fac_c, im_1 = 1.0 / mass[i];
        acc[3 * i] = (qd_1 * s0.x) * im_1;
        acc[3 * i + 1] = (qd_1 * s1.x) * im_1;
        acc[3 * i + 2] = (qd_1 * s2.x + q_1 * P.E_z) * im_1;
    }
    if (i + 1 < n) {
        const double q_1 = pq[i + 1].w, qd_1 = q_1 * rb2k::div_fac_c, im_1 = 1.0 / mass[i + 1];
        acc[3 * i + 3] = (qd_1 * s0.y) * im_1;
        acc[3 * i + 4] = (qd_1 * s1.y) * im_1;
        acc[3 * i + 5] = (qd_1 * s2.y + q_1 * P.E_z) * im_1;
    }
}

double *raw_of(void *block, int npad_max, int parity)
{
    return reinterpret_cast<double *>(static_cast<char *>(block) + P2P_FLAG_BYTES) + (size_t)parity * 3 * (size_t)npad_max;
}

}  // namespace

// Partial-sum buffer of the evaluation that starts now (called by rb2_launch_accel_sym_partial when attached).
double *rb2_p2p_begin_evaluation(Rb2Ctx &ctx, int n_pad)
{
    if (n_pad > ctx.p2p_npad_max) {
        rb2_fail(RB2_ERR_CAPACITY, "the peer exchange block holds %d padded particles, %d needed", ctx.p2p_npad_max, n_pad);
        return nullptr;
    }
    // the evaluation number is only committed when its signal is queued (rb2_launch_accel_sym_exchange_finalize): an
    // error on the way there leaves the protocol where it was
    return raw_of(ctx.p2p_local, ctx.p2p_npad_max, (int)((ctx.p2p_epoch + 1) & 1));
}

// After a stream synchronisation: did the last exchange fail?
int rb2_p2p_check(Rb2Ctx &ctx)
{
    if (!ctx.p2p_err || *ctx.p2p_err == 0) return RB2_OK;
    const int e = *ctx.p2p_err;
    *ctx.p2p_err = 0;
    if (e == P2P_ERR_TIMEOUT)
        return rb2_fail(RB2_ERR_CUDA, "peer exchange: a rank did not publish its partial sums of evaluation %llu within 20 s "
                                      "(every pair-symmetric evaluation is collective once the peers are attached)", ctx.p2p_epoch);
    return rb2_fail(RB2_ERR_CUDA, "peer exchange: a rank published evaluation %llu for a different particle count / sequence "
                                  "of evaluations than this rank", ctx.p2p_epoch);
}

int rb2_launch_accel_sym_exchange_finalize(Rb2Ctx &ctx, const double4 *pq, const double *mass, int n, double *acc_out)
{
    if (n < 1) return RB2_OK;
    ctx.p2p_epoch += 1;
    P2PPeers pp{};
    for (int r = 0; r < ctx.p2p_world; ++r) {
        pp.raw[r] = raw_of(ctx.p2p_peer[r], ctx.p2p_npad_max, (int)(ctx.p2p_epoch & 1));
        pp.flags[r] = static_cast<unsigned long long *>(ctx.p2p_peer[r]);
    }
    const StepParams SP = rb2_make_step_params(ctx.cfg);
    k_p2p_signal<<<1, 32, 0, ctx.stream>>>(pp, ctx.pair_rank, ctx.p2p_world, ctx.p2p_epoch, ctx.sym_n_pad);
    RB2_CUDA(cudaGetLastError());
    const int threads = 256, per_block = 2 * threads;
    k_sym_finalize_p2p<<<(n + per_block - 1) / per_block, threads, 0, ctx.stream>>>(
        n, ctx.sym_n_pad, pp, ctx.pair_rank, ctx.p2p_world, ctx.p2p_epoch, pq, mass, SP.pl, acc_out, ctx.p2p_err_dev);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(2);
    RB2_CUDA(rb2_event_record(ctx, ctx.ev_a1));
    return RB2_OK;
}

// One process, several devices (rb2_set_devices): every context gets an exchange block and the plain device pointers of
// all the others (peer access is on); from there the protocol is the one of the attached processes.
int rb2_p2p_link_local(Rb2Ctx *all, int n)
{
    const int npad_max = ((all[0].cap + 255) / 256) * 256;
    const size_t bytes = P2P_FLAG_BYTES + 2 * 3 * (size_t)npad_max * sizeof(double);
    for (int d = 0; d < n; ++d) {
        Rb2Ctx &c = all[d];
        RB2_CUDA(cudaSetDevice(c.dev));
        RB2_CUDA(cudaMalloc(&c.p2p_local, bytes));
        RB2_CUDA(cudaMemset(c.p2p_local, 0, bytes));
        const P2PHeader hdr = {P2P_MAGIC, (unsigned long long)npad_max};
        RB2_CUDA(cudaMemcpy(static_cast<char *>(c.p2p_local) + P2P_HEADER_OFFSET, &hdr, sizeof(hdr), cudaMemcpyHostToDevice));
        c.p2p_npad_max = npad_max;
        RB2_CUDA(cudaHostAlloc((void **)&c.p2p_err, sizeof(int), cudaHostAllocMapped));
        *c.p2p_err = 0;
        RB2_CUDA(cudaHostGetDevicePointer((void **)&c.p2p_err_dev, (void *)c.p2p_err, 0));
    }
    for (int d = 0; d < n; ++d) {
        Rb2Ctx &c = all[d];
        for (int r = 0; r < n; ++r) c.p2p_peer[r] = all[r].p2p_local;
        c.p2p_world = n;
        c.p2p_ipc = false;
        c.pair_rank = d;
        c.pair_world = n;
        c.p2p_epoch = 0;
    }
    return RB2_OK;
}

int rb2_p2p_release(Rb2Ctx &c)
{
    for (int r = 0; r < c.p2p_world; ++r)
        if (c.p2p_ipc && c.p2p_peer[r] && c.p2p_peer[r] != c.p2p_local) cudaIpcCloseMemHandle(c.p2p_peer[r]);
    for (int r = 0; r < RB2_P2P_MAX; ++r) c.p2p_peer[r] = nullptr;
    c.p2p_world = 0;
    if (c.p2p_local) cudaFree(c.p2p_local);
    c.p2p_local = nullptr;
    c.p2p_npad_max = 0;
    c.p2p_epoch = 0;
    if (c.p2p_err) cudaFreeHost((void *)c.p2p_err);
    c.p2p_err = nullptr;
    c.p2p_err_dev = nullptr;
    return RB2_OK;
}

extern "C" {

int rb2_p2p_export(int n_max, void *handle_out)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    if (g_rb2_ndev > 1) return rb2_fail(RB2_ERR_ARG, "rb2_p2p_export: this process already drives several devices (rb2_set_devices)");
    if (n_max < 1 || !handle_out) return rb2_fail(RB2_ERR_ARG, "rb2_p2p_export: n_max >= 1 and a handle buffer are required");
    static_assert(sizeof(cudaIpcMemHandle_t) == RB2_P2P_HANDLE_BYTES, "handle size");
    RB2_CUDA(cudaStreamSynchronize(c.stream));
    rb2_p2p_release(c);
    const int npad_max = ((n_max + 255) / 256) * 256;
    const size_t bytes = P2P_FLAG_BYTES + 2 * 3 * (size_t)npad_max * sizeof(double);
    RB2_CUDA(cudaMalloc(&c.p2p_local, bytes));
    RB2_CUDA(cudaMemset(c.p2p_local, 0, bytes));
    const P2PHeader hdr = {P2P_MAGIC, (unsigned long long)npad_max};
    RB2_CUDA(cudaMemcpy(static_cast<char *>(c.p2p_local) + P2P_HEADER_OFFSET, &hdr, sizeof(hdr), cudaMemcpyHostToDevice));
    c.p2p_npad_max = npad_max;
    RB2_CUDA(cudaHostAlloc((void **)&c.p2p_err, sizeof(int), cudaHostAllocMapped));
    *c.p2p_err = 0;
    RB2_CUDA(cudaHostGetDevicePointer((void **)&c.p2p_err_dev, (void *)c.p2p_err, 0));
    cudaIpcMemHandle_t h;
    RB2_CUDA(cudaIpcGetMemHandle(&h, c.p2p_local));
    memcpy(handle_out, &h, sizeof(h));
    return RB2_OK;
}

int rb2_p2p_attach(int world, int rank, const void *handles)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    if (g_rb2_ndev > 1) return rb2_fail(RB2_ERR_ARG, "rb2_p2p_attach: this process already drives several devices (rb2_set_devices)");
    if (!c.p2p_local) return rb2_fail(RB2_ERR_ARG, "rb2_p2p_attach before rb2_p2p_export");
    if (world < 1 || world > RB2_P2P_MAX || rank < 0 || rank >= world || !handles)
        return rb2_fail(RB2_ERR_ARG, "rb2_p2p_attach: bad rank %d of %d (at most %d)", rank, world, RB2_P2P_MAX);
    for (int r = 0; r < world; ++r) {
        if (r == rank) { c.p2p_peer[r] = c.p2p_local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + (size_t)r * RB2_P2P_HANDLE_BYTES, sizeof(h));
        void *p = nullptr;
        RB2_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c.p2p_peer[r] = p;
        // the slot offsets inside a block depend on n_max: every rank must have exported for the same one
        P2PHeader hdr{};
        RB2_CUDA(cudaMemcpy(&hdr, static_cast<char *>(p) + P2P_HEADER_OFFSET, sizeof(hdr), cudaMemcpyDeviceToHost));
        if (hdr.magic != P2P_MAGIC || hdr.npad_max != (unsigned long long)c.p2p_npad_max) {
            const unsigned long long got = hdr.npad_max;
            c.p2p_world = r + 1;  // so that the release below closes what has been opened
            rb2_p2p_release(c);
            return rb2_fail(RB2_ERR_ARG, "rb2_p2p_attach: rank %d exported an exchange block for %llu padded particles, this rank "
                                         "for another size (rb2_p2p_export must be called with the same n_max everywhere)", r, got);
        }
    }
    c.p2p_world = world;
    c.p2p_ipc = true;
    c.pair_rank = rank;
    c.pair_world = world;
    c.p2p_epoch = 0;
    return RB2_OK;
}

int rb2_p2p_detach(void)
{
    RB2_REQUIRE_INIT();
    RB2_CUDA(cudaStreamSynchronize(g_rb2.stream));
    return rb2_p2p_release(g_rb2);
}

}  // extern "C"

// rb2_api.cu -- the C ABI declared in include/rumdeed_b200.h: state management, host<->device
// plumbing and the per-step orchestration.  All compute is in rb2_pair.cu / rb2_integrate.cu;
// there is no CPU fallback anywhere in this library.
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include "rb2_internal.cuh"

Rb2Ctx  g_rb2_all[RB2_P2P_MAX];
Rb2Ctx *g_rb2_cur = &g_rb2_all[0];
int     g_rb2_ndev = 1;
char    g_rb2_err[512] = "no error";

// Run f on the context of every device of this process (rb2_set_devices), the first device last-selected again.
template <class F>
static int each_device(F f)
{
    if (g_rb2_ndev <= 1) return f();
    int rc0 = RB2_OK;
    for (int d = 0; d < g_rb2_ndev; ++d) {
        g_rb2_cur = &g_rb2_all[d];
        cudaSetDevice(g_rb2_cur->dev);
        const int rc = f();
        if (rc != RB2_OK && rc0 == RB2_OK) rc0 = rc;
    }
    g_rb2_cur = &g_rb2_all[0];
    cudaSetDevice(g_rb2_cur->dev);
    return rc0;
}

int rb2_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_rb2_err, sizeof(g_rb2_err), fmt, ap);
    va_end(ap);
    return code;
}

StepParams rb2_make_step_params(const rb2_config &c)
{
    StepParams P{};
    P.geometry = c.geometry;
    P.dt = c.time_step;
    P.dt2 = c.time_step * c.time_step;  // time_step2 = time_step**2
    P.box_z = c.box_dim[2];
    P.d = c.d;
    P.planes_N = c.planes_N < 0 ? 0 : (c.planes_N > RB2_PLANES_MAX ? RB2_PLANES_MAX : c.planes_N);
    for (int k = 0; k < RB2_PLANES_MAX; ++k) P.planes_z[k] = c.planes_z[k];
    P.pl.two_d = 2.0 * c.d;
    P.pl.E_z = c.E_z;
    P.pl.nic = c.N_ic_max;
    P.pl.do_ic = c.image_charge;
    // far image partners without the softening term: gap of at least 1 um, and no charged particle survives outside it
    P.pl.far_ok = (c.d >= 1.0e-6 && c.box_dim[2] <= c.d * (1.0 + 1.0e-9)) ? 1 : 0;
    P.tip.a_foci = c.a_foci;
    P.tip.shift_z = c.shift_z;
    P.tip.pre_fac_E_tip = c.pre_fac_E_tip;
    P.tip.eta_1 = c.eta_1;
    P.tip.z_0 = c.h_tip - c.r_tip;
    P.tip.r_tip = c.r_tip;
    P.tip.unit_scale_num = c.pre_fac_E_tip_unit_voltage;
    P.tip.unit_scale_den = c.pre_fac_E_tip;
    P.tip.do_ic = c.image_charge;
    return P;
}

namespace {

template <class T>
int dev_alloc(T **p, size_t count)
{
    RB2_CUDA(cudaMalloc((void **)p, (count > 0 ? count : 1) * sizeof(T)));
    return RB2_OK;
}

int alloc_arrays(DevArrays &A, size_t cap)
{
    int rc;
    if ((rc = dev_alloc(&A.pq, cap))) return rc;
    if ((rc = dev_alloc(&A.prev_pos, 3 * cap))) return rc;
    if ((rc = dev_alloc(&A.vel, 3 * cap))) return rc;
    if ((rc = dev_alloc(&A.acc, 3 * cap))) return rc;
    if ((rc = dev_alloc(&A.acc_prev, 3 * cap))) return rc;
    if ((rc = dev_alloc(&A.acc_prev2, 3 * cap))) return rc;
    if ((rc = dev_alloc(&A.mass, cap))) return rc;
    if ((rc = dev_alloc(&A.species, cap))) return rc;
    if ((rc = dev_alloc(&A.step, cap))) return rc;
    if ((rc = dev_alloc(&A.emitter, cap))) return rc;
    if ((rc = dev_alloc(&A.section, cap))) return rc;
    if ((rc = dev_alloc(&A.life, cap))) return rc;
    if ((rc = dev_alloc(&A.id, cap))) return rc;
    return RB2_OK;
}
void free_arrays(DevArrays &A)
{
    cudaFree(A.pq); cudaFree(A.prev_pos); cudaFree(A.vel); cudaFree(A.acc); cudaFree(A.acc_prev); cudaFree(A.acc_prev2);
    cudaFree(A.mass); cudaFree(A.species); cudaFree(A.step); cudaFree(A.emitter); cudaFree(A.section); cudaFree(A.life);
    cudaFree(A.id);
    A = DevArrays{};
}

int check_config(const rb2_config *cfg)
{
    if (!cfg) return rb2_fail(RB2_ERR_ARG, "config is NULL");
    if (cfg->geometry != RB2_GEOM_PLANAR && cfg->geometry != RB2_GEOM_TIP)
        return rb2_fail(RB2_ERR_GEOMETRY, "geometry %d: only planar (1) and hyperboloid tip (2) run on the device", cfg->geometry);
    if (cfg->planes_N < 0 || cfg->planes_N > RB2_PLANES_MAX) return rb2_fail(RB2_ERR_ARG, "planes_N out of range");
    if (cfg->N_ic_max < 0) return rb2_fail(RB2_ERR_ARG, "N_ic_max < 0");
    return RB2_OK;
}

void fill_counts_from_device(Rb2Ctx &c)
{
    const DevCounters &h = *c.h_counters;
    c.counts.nrPart_remove = h.mark_part;
    c.counts.nrElec_remove = h.mark_elec;
    c.counts.nrIon_remove = h.mark_ion;
    c.counts.nrAtom_remove = h.mark_atom;
    c.counts.nrPart_remove_top = h.top_part;
    c.counts.nrPart_remove_bot = h.bot_part;
    c.counts.nrElec_remove_top = h.top_elec;
    c.counts.nrElec_remove_bot = h.bot_elec;
    c.counts.nrIon_remove_top = h.top_ion;
    c.counts.nrIon_remove_bot = h.bot_ion;
    c.counts.nrPart_remove_ion = h.ion_part;
    c.counts.nrElec_remove_ion = h.ion_elec;
    c.counts.nrAtom_remove_ion = h.ion_atom;
}

int fetch_counters(Rb2Ctx &c)
{
    RB2_CUDA(cudaMemcpyAsync(c.h_counters, c.d_counters, sizeof(DevCounters), cudaMemcpyDeviceToHost, c.stream));
    RB2_CUDA(cudaStreamSynchronize(c.stream));
    fill_counts_from_device(c);
    return RB2_OK;
}

void fill_velocity_result(Rb2Ctx &c, rb2_step_result *out)
{
    const double *r = c.h_red;
    for (int k = 0; k < 4; ++k) out->ramo_current[k] = r[k];
    for (int k = 0; k < 3; ++k) {
        // Average_Velocities, src/mod_verlet.F90:428-447
        out->avg_part_vel[k] = (c.counts.nrPart != 0) ? r[4 + k] / c.counts.nrPart : r[4 + k];
        out->avg_elec_vel[k] = (c.counts.nrElec != 0) ? r[7 + k] / c.counts.nrElec : r[7 + k];
        out->avg_ion_vel[k] = (c.counts.nrIon != 0) ? r[10 + k] / c.counts.nrIon : r[10 + k];
    }
}

// Which pair kernel evaluates Calculate_Acceleration_Particles: the pair-symmetric kernel for the planar
// geometry when the whole i-range is evaluated here (pair_mode 2, or auto from sym_min_n particles on),
// the gather kernel otherwise (tip geometry, i-partition in effect, small N).
bool use_sym_kernel(const Rb2Ctx &c, int n, int i0, int i1)
{
    if (c.cfg.geometry != RB2_GEOM_PLANAR || c.pair_mode == 1) return false;
    if (i0 != 0 || i1 != n) return false;
    if (c.pair_mode == 2) return true;
    return n >= c.sym_min_n;
}

int launch_accel_any(Rb2Ctx &c, const double4 *pq, const double *mass, int n, int i0, int i1, double *acc)
{
    if (use_sym_kernel(c, n, i0, i1)) {
        if (c.pair_world > 1 && c.p2p_world != c.pair_world)  // attached: the finalisation below exchanges over peer memory
            return rb2_fail(RB2_ERR_ARG, "pair work is split over %d processes: attach the peers (rb2_p2p_attach) or use "
                                         "rb2_accel_partial / rb2_accel_finalize around your own all-reduce", c.pair_world);
        int rc = rb2_launch_accel_sym_partial(c, pq, n);
        if (rc) return rc;
        c.last_pair_kernel = 2;
        return rb2_launch_accel_sym_finalize(c, pq, mass, n, acc);
    }
    c.last_pair_kernel = 1;
    return rb2_launch_accel(c, pq, mass, n, i0, i1, acc);
}

// Queues the position update, the (device-guarded) record passes, the counter copy and -- for the fused step -- the
// acceleration evaluation.  Nothing here waits for the device: finish_position() reads the results after the
// caller's stream synchronisation.
int do_update_position(Rb2Ctx &c, bool overlap_accel, int *rc_accel)
{
    c.host_events.clear();
    if (c.n < 1) {
        if (overlap_accel && rc_accel) *rc_accel = RB2_OK;
        return RB2_OK;
    }
    RB2_CUDA(cudaMemsetAsync(&c.d_counters->n_events, 0, sizeof(int), c.stream));
    int rc = rb2_launch_update_position(c);
    if (rc) return rc;
    rc = rb2_launch_events(c);  // must see the pre-update velocities: before the velocity update in stream order
    if (rc) return rc;
    RB2_CUDA(cudaMemcpyAsync(c.h_counters, c.d_counters, sizeof(DevCounters), cudaMemcpyDeviceToHost, c.stream));
    if (overlap_accel) {
        int i0 = c.part_begin, i1 = (c.part_end < 0 || c.part_end > c.n) ? c.n : c.part_end;
        if (i0 < 0) i0 = 0;
        *rc_accel = launch_accel_any(c, c.a.pq, c.a.mass, c.n, i0, i1, c.a.acc);
        if (*rc_accel) return *rc_accel;
        c.accel_timed = true;
    }
    return RB2_OK;
}

// After the stream has been synchronised: counters of the position update and its record list on the host.
int finish_position(Rb2Ctx &c, bool after_velocity_update)
{
    if (c.n < 1) return RB2_OK;
    fill_counts_from_device(c);
    const int nev = c.h_counters->n_events;
    if (nev > 0) {
        if (nev > c.ev_cap) {
            int rc = rb2_rebuild_events(c, nev, after_velocity_update);
            if (rc) return rc;
        }
        c.host_events.resize((size_t)nev);
        RB2_CUDA(cudaMemcpyAsync(c.host_events.data(), c.d_events, (size_t)nev * sizeof(rb2_event), cudaMemcpyDeviceToHost, c.stream));
        RB2_CUDA(cudaStreamSynchronize(c.stream));
    }
    return RB2_OK;
}

}  // namespace

int rb2_fetch_counters(Rb2Ctx &c) { return fetch_counters(c); }

int rb2_ensure_stage(Rb2Ctx &c, size_t n_doubles, size_t n_ints)
{
    if (n_doubles > c.stage_d_cap) {
        if (c.d_stage_d) RB2_CUDA(cudaFree(c.d_stage_d));
        c.d_stage_d = nullptr; c.stage_d_cap = 0;
        const size_t want = n_doubles + n_doubles / 2 + 1024;
        RB2_CUDA(cudaMalloc(&c.d_stage_d, want * sizeof(double)));
        c.stage_d_cap = want;
    }
    if (n_ints > c.stage_i_cap) {
        if (c.d_stage_i) RB2_CUDA(cudaFree(c.d_stage_i));
        c.d_stage_i = nullptr; c.stage_i_cap = 0;
        const size_t want = n_ints + n_ints / 2 + 1024;
        RB2_CUDA(cudaMalloc(&c.d_stage_i, want * sizeof(int)));
        c.stage_i_cap = want;
    }
    return RB2_OK;
}

extern "C" {

const char *rb2_last_error_string(void) { return g_rb2_err; }

int rb2_device_available(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) { cudaGetLastError(); return 0; }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
    return prop.major >= 10 ? 1 : 0;
}

// Frees whatever the context holds (also after a failed rb2_init: every pointer that was allocated is non-null, the
// rest is null) and resets it.
static void release_all(Rb2Ctx &c)
{
    if (c.stream) cudaStreamSynchronize(c.stream);
    rb2_p2p_release(c);
    rb2_collisions_release(c);
    free_arrays(c.a);
    free_arrays(c.b);
    cudaFree(c.mask); cudaFree(c.evcnt); cudaFree(c.evbits); cudaFree(c.prefix); cudaFree(c.blocksum); cudaFree(c.life_hist);
    cudaFree(c.d_counters); cudaFree(c.d_red); cudaFree(c.d_redpart); cudaFree(c.d_total); cudaFree(c.partial);
    cudaFree(c.d_ramo_part); cudaFree(c.d_ramo_sec);
    if (c.h_ramo_sec) cudaFreeHost(c.h_ramo_sec);
    cudaFree(c.sym_bufI); cudaFree(c.sym_bufJ); cudaFree(c.sym_raw); cudaFree(c.sym_owner); cudaFree(c.sym_units);
    cudaFree(c.d_events); cudaFree(c.d_pts); cudaFree(c.d_fld); cudaFree(c.d_extra); cudaFree(c.d_stage_d); cudaFree(c.d_stage_i);
    if (c.h_counters) cudaFreeHost(c.h_counters);
    if (c.h_red) cudaFreeHost(c.h_red);
    if (c.h_total) cudaFreeHost(c.h_total);
    if (c.h_pts) cudaFreeHost(c.h_pts);
    if (c.h_fld) cudaFreeHost(c.h_fld);
    if (c.h_stage) cudaFreeHost(c.h_stage);
    cudaFree(c.d_sup_grid); cudaFree(c.d_tip_img); cudaFree(c.d_supq);
    if (c.h_supq) cudaFreeHost(c.h_supq);
    if (c.h_sup) cudaFreeHost(c.h_sup);
    if (c.graph_exec) cudaGraphExecDestroy(c.graph_exec);
    cudaEvent_t evs[] = {c.ev_a0, c.ev_a1, c.ev_s0, c.ev_s1, c.ev_c};
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    if (c.stream) cudaStreamDestroy(c.stream);
    cudaGetLastError();
    c = Rb2Ctx{};
}

static int init_impl(Rb2Ctx &c, const rb2_config *cfg)
{
    int rc = RB2_OK;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) {
        cudaGetLastError();
        return rb2_fail(RB2_ERR_CUDA, "no CUDA device: librumdeed_b200 has no CPU fallback");
    }
    if (cfg->device >= 0) RB2_CUDA(cudaSetDevice(cfg->device));
    RB2_CUDA(cudaGetDevice(&c.dev));
    cudaDeviceProp prop;
    RB2_CUDA(cudaGetDeviceProperties(&prop, c.dev));
    if (prop.major < 10)
        return rb2_fail(RB2_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", c.dev, prop.major, prop.minor);
    c.sm_count = prop.multiProcessorCount;
    c.cfg = *cfg;
    c.cap = cfg->capacity;
    c.n = 0;
    c.counts = rb2_counts{};
    c.part_begin = 0;
    c.part_end = -1;
    c.launches = 0;
    RB2_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    RB2_CUDA(cudaEventCreate(&c.ev_a0));
    RB2_CUDA(cudaEventCreate(&c.ev_a1));
    RB2_CUDA(cudaEventCreate(&c.ev_s0));
    RB2_CUDA(cudaEventCreate(&c.ev_s1));
    RB2_CUDA(cudaEventCreateWithFlags(&c.ev_c, cudaEventDisableTiming));
    const size_t cap = (size_t)c.cap;
    if ((rc = alloc_arrays(c.a, cap))) return rc;
    if ((rc = alloc_arrays(c.b, cap))) return rc;
    if ((rc = dev_alloc(&c.mask, cap))) return rc;
    if ((rc = dev_alloc(&c.evcnt, cap))) return rc;
    if ((rc = dev_alloc(&c.evbits, cap))) return rc;
    if ((rc = dev_alloc(&c.prefix, cap))) return rc;
    if ((rc = dev_alloc(&c.blocksum, cap / 1024 + 8))) return rc;
    if ((rc = dev_alloc(&c.life_hist, (size_t)(RB2_MAX_LIFE_TIME + 1) * 4))) return rc;
    if ((rc = dev_alloc(&c.d_counters, 1))) return rc;
    if ((rc = dev_alloc(&c.d_red, 16))) return rc;
    if ((rc = dev_alloc(&c.d_total, 4))) return rc;
    RB2_CUDA(cudaMallocHost(&c.h_counters, sizeof(DevCounters)));
    RB2_CUDA(cudaMallocHost(&c.h_red, 16 * sizeof(double)));
    RB2_CUDA(cudaMallocHost(&c.h_total, 4 * sizeof(int)));
    memset(c.h_counters, 0, sizeof(DevCounters));
    memset(c.h_red, 0, 16 * sizeof(double));
    RB2_CUDA(cudaMemsetAsync(c.d_counters, 0, sizeof(DevCounters), c.stream));
    RB2_CUDA(cudaMemsetAsync(c.d_red, 0, 16 * sizeof(double), c.stream));
    RB2_CUDA(cudaMemsetAsync(c.life_hist, 0, (size_t)(RB2_MAX_LIFE_TIME + 1) * 4 * sizeof(unsigned long long), c.stream));
    c.init = true;
    if ((rc = rb2_launch_fill_mask(c, c.cap))) return rc;
    RB2_CUDA(cudaStreamSynchronize(c.stream));
    if (const char *e = getenv("RB2_NO_GRAPH")) c.use_graph = atoi(e) != 0 ? 0 : 1;       // same as rb2_set_option("step_graph", 0)
    if (const char *e = getenv("RB2_SYM_WAVES")) { const double v = atof(e); if (v >= 1.0) c.sym_waves = v; }                   // measurement scripts
    if (const char *e = getenv("RB2_SYM_FAR")) c.sym_far = atoi(e) != 0;
    if (const char *e = getenv("RB2_PAIR_MODE")) { const int v = atoi(e); if (v >= 0 && v <= 2) c.pair_mode = v; }
    if (const char *e = getenv("RB2_MH_SMALL")) c.mh_small = atoi(e) != 0;
    if (const char *e = getenv("RB2_MH_SMALL_MAX")) { const int v = atoi(e); if (v >= 1 && v <= 512) c.mh_small_max = v; }
    if (const char *e = getenv("RB2_MH_CTAS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v <= 4) c.mh_ctas_per_sm = v; }  // same as rb2_set_option("mh_small", ..)
    return RB2_OK;
}

int rb2_init(const rb2_config *cfg)
{
    rb2_finalize();  // every device context of a previous set-up
    Rb2Ctx &c = g_rb2;
    int rc = check_config(cfg);
    if (rc) return rc;
    if (cfg->capacity < 1) return rb2_fail(RB2_ERR_ARG, "capacity must be >= 1");
    rc = init_impl(c, cfg);
    if (rc) release_all(c);  // a failed allocation half way leaves nothing behind (the error string is kept)
    return rc;
}

int rb2_finalize(void)
{
    for (int d = g_rb2_ndev - 1; d >= 0; --d) {
        Rb2Ctx &c = g_rb2_all[d];
        if (!c.init) continue;
        g_rb2_cur = &c;
        cudaSetDevice(c.dev);
        release_all(c);
    }
    g_rb2_cur = &g_rb2_all[0];
    g_rb2_ndev = 1;
    return RB2_OK;
}

// One process, several GPUs.  devices[0] must be the device of rb2_init; a replica of the (still empty) particle store is
// set up on each further device, peer access is enabled between all of them and every context gets the exchange block
// of every other one (plain peer pointers -- no IPC), so that the pair work of rb2_step / rb2_accel_only /
// rb2_accel_host is split over the devices exactly as it is over the ranks of the one-process-per-GPU set-up.
int rb2_set_devices(int n_devices, const int *devices)
{
    RB2_REQUIRE_INIT();
    if (n_devices < 1 || n_devices > RB2_P2P_MAX || !devices) return rb2_fail(RB2_ERR_ARG, "rb2_set_devices: 1..%d devices", RB2_P2P_MAX);
    if (g_rb2_ndev > 1) return rb2_fail(RB2_ERR_ARG, "rb2_set_devices: already set (call rb2_init again first)");
    Rb2Ctx &c0 = g_rb2_all[0];
    if (devices[0] != c0.dev) return rb2_fail(RB2_ERR_ARG, "rb2_set_devices: devices[0] must be the device of rb2_init (%d)", c0.dev);
    if (c0.n != 0) return rb2_fail(RB2_ERR_ARG, "rb2_set_devices must be called before particles are uploaded or added");
    if (c0.p2p_world > 1) return rb2_fail(RB2_ERR_ARG, "rb2_set_devices: this context is attached to other processes (rb2_p2p_attach)");
    if (c0.cfg.geometry != RB2_GEOM_PLANAR) return rb2_fail(RB2_ERR_GEOMETRY, "rb2_set_devices: the pair work is split for the planar geometry only");
    if (n_devices == 1) return RB2_OK;
    int count = 0;
    RB2_CUDA(cudaGetDeviceCount(&count));
    for (int d = 0; d < n_devices; ++d) {
        if (devices[d] < 0 || devices[d] >= count) return rb2_fail(RB2_ERR_ARG, "rb2_set_devices: no device %d", devices[d]);
        for (int e = 0; e < d; ++e)
            if (devices[e] == devices[d]) return rb2_fail(RB2_ERR_ARG, "rb2_set_devices: device %d listed twice", devices[d]);
    }
    int rc = RB2_OK;
    for (int d = 1; d < n_devices && rc == RB2_OK; ++d) {
        rb2_config cfg = c0.cfg;
        cfg.device = devices[d];
        g_rb2_cur = &g_rb2_all[d];
        rc = init_impl(g_rb2_all[d], &cfg);
        if (rc == RB2_OK) {
            Rb2Ctx &cd = g_rb2_all[d];
            cd.pair_mode = c0.pair_mode; cd.sym_min_n = c0.sym_min_n; cd.sym_tpl = c0.sym_tpl; cd.sym_waves = c0.sym_waves;
            cd.sym_budget_bytes = c0.sym_budget_bytes; cd.sym_kmax = c0.sym_kmax; cd.sym_gmax = c0.sym_gmax; cd.ev_min = c0.ev_min;
            cd.ramo_n_sec = c0.ramo_n_sec; cd.ramo_n_emit = c0.ramo_n_emit;
        }
        g_rb2_ndev = d + 1;
    }
    for (int a = 0; a < n_devices && rc == RB2_OK; ++a)
        for (int b = 0; b < n_devices && rc == RB2_OK; ++b) {
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
            if (!can) { rc = rb2_fail(RB2_ERR_CUDA, "rb2_set_devices: device %d cannot access the memory of device %d", devices[a], devices[b]); break; }
            cudaSetDevice(devices[a]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = rb2_fail(RB2_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", devices[a], devices[b], cudaGetErrorString(e));
            cudaGetLastError();
        }
    if (rc == RB2_OK) rc = rb2_p2p_link_local(g_rb2_all, n_devices);
    g_rb2_cur = &g_rb2_all[0];
    cudaSetDevice(c0.dev);
    if (rc != RB2_OK) {  // back to one device
        for (int d = g_rb2_ndev - 1; d >= 1; --d) {
            g_rb2_cur = &g_rb2_all[d];
            cudaSetDevice(g_rb2_all[d].dev);
            release_all(g_rb2_all[d]);
        }
        g_rb2_cur = &g_rb2_all[0];
        g_rb2_ndev = 1;
        cudaSetDevice(c0.dev);
        rb2_p2p_release(c0);
        c0.pair_rank = 0; c0.pair_world = 1;
    }
    return rc;
}

static int update_config_one(const rb2_config *cfg)
{
    RB2_REQUIRE_INIT();
    int rc = check_config(cfg);
    if (rc) return rc;
    const int cap = g_rb2.cfg.capacity, dev = g_rb2.cfg.device;
    g_rb2.cfg = *cfg;
    g_rb2.cfg.capacity = cap;  // capacity and device are fixed at init
    g_rb2.cfg.device = dev;
    return RB2_OK;
}
int rb2_update_config(const rb2_config *cfg)
{
    return each_device([&]() -> int { return update_config_one(cfg); });
}

static int upload_particles_one(int n, const double *pos, const double *prev_pos, const double *vel, const double *acc,
                         const double *acc_prev, const double *acc_prev2, const double *charge, const double *mass,
                         const int *species, const int *step, const int *emitter, const int *section, const int *life,
                         const int *id, int nrID)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    if (n < 0 || n > c.cap) return rb2_fail(RB2_ERR_CAPACITY, "n = %d exceeds capacity %d", n, c.cap);
    if (n > 0 && (!pos || !charge || !mass)) return rb2_fail(RB2_ERR_ARG, "pos, charge and mass are required");
    const size_t b3 = (size_t)n * 3 * sizeof(double), b1 = (size_t)n * sizeof(double), bi = (size_t)n * sizeof(int);
    cudaStream_t st = c.stream;
    c.far_state_ok = true;
    if (n > 0) {
        // pos + charge travel through the spare set and are packed into pq on the device
        RB2_CUDA(cudaMemcpyAsync(c.b.prev_pos, pos, b3, cudaMemcpyHostToDevice, st));
        if (c.cfg.geometry == RB2_GEOM_PLANAR) c.far_state_ok = rb2_all_between_plates(pos, n, c.cfg.d);  // (while the copy runs)
        RB2_CUDA(cudaMemcpyAsync(c.b.mass, charge, b1, cudaMemcpyHostToDevice, st));
        int rc = rb2_launch_pack(c, c.b.prev_pos, c.b.mass, n, c.a.pq);
        if (rc) return rc;
        RB2_CUDA(cudaMemcpyAsync(c.a.mass, mass, b1, cudaMemcpyHostToDevice, st));
#define RB2_UP3(dst, src)                                                          \
    if (src) RB2_CUDA(cudaMemcpyAsync(dst, src, b3, cudaMemcpyHostToDevice, st)); \
    else RB2_CUDA(cudaMemsetAsync(dst, 0, b3, st))
        RB2_UP3(c.a.prev_pos, prev_pos);
        RB2_UP3(c.a.vel, vel);
        RB2_UP3(c.a.acc, acc);
        RB2_UP3(c.a.acc_prev, acc_prev);
        RB2_UP3(c.a.acc_prev2, acc_prev2);
#undef RB2_UP3
        if (species) RB2_CUDA(cudaMemcpyAsync(c.a.species, species, bi, cudaMemcpyHostToDevice, st));
        if (step) RB2_CUDA(cudaMemcpyAsync(c.a.step, step, bi, cudaMemcpyHostToDevice, st));
        if (emitter) RB2_CUDA(cudaMemcpyAsync(c.a.emitter, emitter, bi, cudaMemcpyHostToDevice, st));
        if (section) RB2_CUDA(cudaMemcpyAsync(c.a.section, section, bi, cudaMemcpyHostToDevice, st));
        if (life) RB2_CUDA(cudaMemcpyAsync(c.a.life, life, bi, cudaMemcpyHostToDevice, st));
        if (id) RB2_CUDA(cudaMemcpyAsync(c.a.id, id, bi, cudaMemcpyHostToDevice, st));
        rc = rb2_launch_fill_defaults(c, n, !species, !step, !emitter, !section, !life, !id);
        if (rc) return rc;
    }
    int rc = rb2_launch_fill_mask(c, c.n > n ? c.n : n);
    if (rc) return rc;
    RB2_CUDA(cudaMemsetAsync(c.d_counters, 0, sizeof(DevCounters), st));
    RB2_CUDA(cudaStreamSynchronize(st));
    memset(c.h_counters, 0, sizeof(DevCounters));
    c.n = n;
    rb2_counts k{};
    k.nrPart = n;
    if (species) {
        for (int i = 0; i < n; ++i) {
            if (species[i] == RB2_SPECIES_ELEC) k.nrElec++;
            else if (species[i] == RB2_SPECIES_ION) k.nrIon++;
            else if (species[i] == RB2_SPECIES_ATOM) k.nrAtom++;
        }
    } else {
        k.nrElec = n;
    }
    k.nrID = nrID >= 0 ? nrID : n;
    k.nrPart_dropped = c.counts.nrPart_dropped;
    c.counts = k;
    c.host_events.clear();
    return RB2_OK;
}
int rb2_upload_particles(int n, const double *pos, const double *prev_pos, const double *vel, const double *acc,
                         const double *acc_prev, const double *acc_prev2, const double *charge, const double *mass,
                         const int *species, const int *step, const int *emitter, const int *section, const int *life,
                         const int *id, int nrID)
{
    return each_device([&]() -> int { return upload_particles_one(n, pos, prev_pos, vel, acc, acc_prev, acc_prev2, charge, mass, species, step, emitter, section, life, id, nrID); });
}

int rb2_download_particles(double *pos, double *prev_pos, double *vel, double *acc, double *acc_prev, double *acc_prev2,
                           double *charge, double *mass, int *species, int *step, int *emitter, int *section, int *life,
                           int *id, int *mask)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    const int n = c.n;
    if (n < 1) return RB2_OK;
    const size_t b3 = (size_t)n * 3 * sizeof(double), b1 = (size_t)n * sizeof(double), bi = (size_t)n * sizeof(int);
    cudaStream_t st = c.stream;
    if (pos || charge) {
        int rc = rb2_launch_unpack(c, c.a.pq, n, pos ? c.b.prev_pos : nullptr, charge ? c.b.mass : nullptr);
        if (rc) return rc;
        if (pos) RB2_CUDA(cudaMemcpyAsync(pos, c.b.prev_pos, b3, cudaMemcpyDeviceToHost, st));
        if (charge) RB2_CUDA(cudaMemcpyAsync(charge, c.b.mass, b1, cudaMemcpyDeviceToHost, st));
    }
    if (prev_pos) RB2_CUDA(cudaMemcpyAsync(prev_pos, c.a.prev_pos, b3, cudaMemcpyDeviceToHost, st));
    if (vel) RB2_CUDA(cudaMemcpyAsync(vel, c.a.vel, b3, cudaMemcpyDeviceToHost, st));
    if (acc) RB2_CUDA(cudaMemcpyAsync(acc, c.a.acc, b3, cudaMemcpyDeviceToHost, st));
    if (acc_prev) RB2_CUDA(cudaMemcpyAsync(acc_prev, c.a.acc_prev, b3, cudaMemcpyDeviceToHost, st));
    if (acc_prev2) RB2_CUDA(cudaMemcpyAsync(acc_prev2, c.a.acc_prev2, b3, cudaMemcpyDeviceToHost, st));
    if (mass) RB2_CUDA(cudaMemcpyAsync(mass, c.a.mass, b1, cudaMemcpyDeviceToHost, st));
    if (species) RB2_CUDA(cudaMemcpyAsync(species, c.a.species, bi, cudaMemcpyDeviceToHost, st));
    if (step) RB2_CUDA(cudaMemcpyAsync(step, c.a.step, bi, cudaMemcpyDeviceToHost, st));
    if (emitter) RB2_CUDA(cudaMemcpyAsync(emitter, c.a.emitter, bi, cudaMemcpyDeviceToHost, st));
    if (section) RB2_CUDA(cudaMemcpyAsync(section, c.a.section, bi, cudaMemcpyDeviceToHost, st));
    if (life) RB2_CUDA(cudaMemcpyAsync(life, c.a.life, bi, cudaMemcpyDeviceToHost, st));
    if (id) RB2_CUDA(cudaMemcpyAsync(id, c.a.id, bi, cudaMemcpyDeviceToHost, st));
    if (mask) RB2_CUDA(cudaMemcpyAsync(mask, c.mask, bi, cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaStreamSynchronize(st));
    return RB2_OK;
}

int rb2_get_counts(rb2_counts *out)
{
    RB2_REQUIRE_INIT();
    if (!out) return rb2_fail(RB2_ERR_ARG, "out is NULL");
    int rc = fetch_counters(g_rb2);
    if (rc) return rc;
    *out = g_rb2.counts;
    return RB2_OK;
}

static int add_particles_one(int k, const double *pos, const double *vel, const int *species, int step, const int *emit,
                      const int *sec, const int *life)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    if (k < 0) return rb2_fail(RB2_ERR_ARG, "k < 0");
    if (k == 0) return RB2_OK;
    if (!pos || !vel || !species) return rb2_fail(RB2_ERR_ARG, "pos, vel and species are required");
    for (int t = 0; t < k; ++t)
        if (species[t] < RB2_SPECIES_ELEC || species[t] > RB2_SPECIES_ATOM)
            return rb2_fail(RB2_ERR_ARG, "unknown particle species %d", species[t]);  // reference: ERROR UNKNOWN PARTICLE TYPE
    int room = c.cap - c.n;
    int kk = k <= room ? k : (room > 0 ? room : 0);
    c.counts.nrPart_dropped += (k - kk);  // src/mod_pair.F90:37-43
    if (kk == 0) return RB2_OK;
    int rc = rb2_ensure_stage(c, (size_t)6 * kk, (size_t)4 * kk);
    if (rc) return rc;
    std::vector<int> tmp((size_t)3 * kk);
    for (int t = 0; t < kk; ++t) {
        tmp[t] = emit ? emit[t] : 1;
        int s = sec ? sec[t] : 1;
        if (s > 96 * 96) s = 96 * 96;  // MAX_SECTIONS clamp, src/mod_pair.F90:46-51
        tmp[kk + t] = s;
        tmp[2 * kk + t] = life ? life[t] : -1;
    }
    cudaStream_t st = c.stream;
    if (c.cfg.geometry == RB2_GEOM_PLANAR && c.far_state_ok) c.far_state_ok = rb2_all_between_plates(pos, kk, c.cfg.d);
    double *d_pos = c.d_stage_d, *d_vel = c.d_stage_d + (size_t)3 * kk;
    int *d_sp = c.d_stage_i, *d_emit = d_sp + kk, *d_sec = d_sp + 2 * kk, *d_life = d_sp + 3 * kk;
    RB2_CUDA(cudaMemcpyAsync(d_pos, pos, (size_t)3 * kk * sizeof(double), cudaMemcpyHostToDevice, st));
    RB2_CUDA(cudaMemcpyAsync(d_vel, vel, (size_t)3 * kk * sizeof(double), cudaMemcpyHostToDevice, st));
    RB2_CUDA(cudaMemcpyAsync(d_sp, species, (size_t)kk * sizeof(int), cudaMemcpyHostToDevice, st));
    RB2_CUDA(cudaMemcpyAsync(d_emit, tmp.data(), (size_t)3 * kk * sizeof(int), cudaMemcpyHostToDevice, st));
    rc = rb2_launch_add(c, kk, c.n, c.counts.nrID, step, d_pos, d_vel, d_sp, d_emit, d_sec, d_life);
    if (rc) return rc;
    RB2_CUDA(cudaStreamSynchronize(st));  // tmp and the caller's arrays may go away
    for (int t = 0; t < kk; ++t) {
        if (species[t] == RB2_SPECIES_ELEC) c.counts.nrElec++;
        else if (species[t] == RB2_SPECIES_ION) c.counts.nrIon++;
        else c.counts.nrAtom++;
    }
    c.n += kk;
    c.counts.nrPart = c.n;
    c.counts.nrID += kk;
    return RB2_OK;
}
int rb2_add_particles(int k, const double *pos, const double *vel, const int *species, int step, const int *emit,
                      const int *sec, const int *life)
{
    return each_device([&]() -> int { return add_particles_one(k, pos, vel, species, step, emit, sec, life); });
}

int rb2_capacity_left(int *out)
{
    RB2_REQUIRE_INIT();
    if (!out) return rb2_fail(RB2_ERR_ARG, "out is NULL");
    *out = g_rb2.cap - g_rb2.n;
    return RB2_OK;
}

static int mark_remove_one(int k, const int *index, const int *reason)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    if (k < 0) return rb2_fail(RB2_ERR_ARG, "k < 0");
    if (k == 0) return RB2_OK;
    if (!index || !reason) return rb2_fail(RB2_ERR_ARG, "index and reason are required");
    for (int t = 0; t < k; ++t)
        if (index[t] < 0 || index[t] >= c.n) return rb2_fail(RB2_ERR_ARG, "rb2_mark_remove: index %d outside 0..%d", index[t], c.n - 1);
    for (int t = 0; t < k; ++t)  // reference: 'Error unknown remove case' (src/mod_pair.F90:280-282); species-specific cases are checked on the device
        if (reason[t] < RB2_REMOVE_TOP || reason[t] > RB2_REMOVE_ION) return rb2_fail(RB2_ERR_ARG, "rb2_mark_remove: unknown remove case %d", reason[t]);
    int rc = rb2_ensure_stage(c, 0, (size_t)2 * k);
    if (rc) return rc;
    RB2_CUDA(cudaMemcpyAsync(c.d_stage_i, index, (size_t)k * sizeof(int), cudaMemcpyHostToDevice, c.stream));
    RB2_CUDA(cudaMemcpyAsync(c.d_stage_i + k, reason, (size_t)k * sizeof(int), cudaMemcpyHostToDevice, c.stream));
    rc = rb2_launch_mark(c, k, c.d_stage_i, c.d_stage_i + k);
    if (rc) return rc;
    return fetch_counters(c);
}
int rb2_mark_remove(int k, const int *index, const int *reason)
{
    return each_device([&]() -> int { return mark_remove_one(k, index, reason); });
}

static int remove_marked_one(int step, rb2_counts *out)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    int rc = fetch_counters(c);
    if (rc) return rc;
    const DevCounters h = *c.h_counters;
    if (h.mark_part > 0 && c.n > 0) {  // src/mod_pair.F90:358
        const int n_old = c.n;
        if (n_old - h.mark_part > 0) {
            rc = rb2_launch_compact(c, step);
            if (rc) return rc;
        }
        c.counts.nrElec -= h.mark_elec;
        c.counts.nrIon -= h.mark_ion;
        c.counts.nrAtom -= h.mark_atom;
        if (c.counts.nrElec < 0) c.counts.nrElec = 0;
        if (c.counts.nrIon < 0) c.counts.nrIon = 0;
        c.counts.nrPart = c.counts.nrElec + c.counts.nrIon + c.counts.nrAtom;
        c.n = c.counts.nrPart;
        rc = rb2_launch_fill_mask(c, n_old);
        if (rc) return rc;
        RB2_CUDA(cudaMemsetAsync(c.d_counters, 0, sizeof(DevCounters), c.stream));
        RB2_CUDA(cudaStreamSynchronize(c.stream));
        memset(c.h_counters, 0, sizeof(DevCounters));
        fill_counts_from_device(c);
    }
    if (out) *out = c.counts;
    return RB2_OK;
}
int rb2_remove_marked(int step, rb2_counts *out)
{
    return each_device([&]() -> int { return remove_marked_one(step, out); });
}

int rb2_get_life_time(long long *out)
{
    RB2_REQUIRE_INIT();
    if (!out) return rb2_fail(RB2_ERR_ARG, "out is NULL");
    RB2_CUDA(cudaMemcpyAsync(out, g_rb2.life_hist, (size_t)(RB2_MAX_LIFE_TIME + 1) * 4 * sizeof(long long),
                             cudaMemcpyDeviceToHost, g_rb2.stream));
    RB2_CUDA(cudaStreamSynchronize(g_rb2.stream));
    return RB2_OK;
}

static int update_position_one(int step)
{
    (void)step;
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    int rc = do_update_position(c, false, nullptr);
    if (rc) return rc;
    RB2_CUDA(cudaStreamSynchronize(c.stream));
    return finish_position(c, false);
}
int rb2_update_position(int step)
{
    return each_device([&]() -> int { return update_position_one(step); });
}

int rb2_accel_only(void)
{
    RB2_REQUIRE_INIT();
    // queue on every device first, wait afterwards: the devices work side by side and exchange inside their kernels
    int rc = each_device([&]() -> int {
        Rb2Ctx &c = g_rb2;
        int i0 = c.part_begin < 0 ? 0 : c.part_begin;
        int i1 = (c.part_end < 0 || c.part_end > c.n) ? c.n : c.part_end;
        int rc1 = launch_accel_any(c, c.a.pq, c.a.mass, c.n, i0, i1, c.a.acc);
        if (rc1) return rc1;
        c.accel_timed = (c.n > 0 && i1 > i0);
        return RB2_OK;
    });
    const int rc2 = each_device([&]() -> int {
        Rb2Ctx &c = g_rb2;
        RB2_CUDA(cudaStreamSynchronize(c.stream));
        return rb2_p2p_check(c);
    });
    return rc ? rc : rc2;
}

static int update_velocity_one(rb2_step_result *out)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    int rc = rb2_launch_update_velocity(c);
    if (rc) return rc;
    if ((rc = rb2_launch_ramo_sections(c))) return rc;
    RB2_CUDA(cudaMemcpyAsync(c.h_red, c.d_red, 16 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    RB2_CUDA(cudaStreamSynchronize(c.stream));
    if (out) {
        memset(out, 0, sizeof(*out));
        fill_velocity_result(c, out);
        out->n_events = (int)c.host_events.size();
        out->counts = c.counts;
    }
    return RB2_OK;
}
int rb2_update_velocity(rb2_step_result *out)
{
    return each_device([&]() -> int { return update_velocity_one(out); });
}

// Everything rb2_step queues on the stream, from the first event record to the last.
static int queue_step(Rb2Ctx &c)
{
    RB2_CUDA(rb2_event_record(c, c.ev_s0));
    int rc_accel = RB2_OK;
    c.accel_timed = false;
    int rc = do_update_position(c, true, &rc_accel);
    if (rc) return rc;
    rc = rb2_launch_update_velocity(c);
    if (rc) return rc;
    if ((rc = rb2_launch_ramo_sections(c))) return rc;
    RB2_CUDA(cudaMemcpyAsync(c.h_red, c.d_red, 16 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    RB2_CUDA(rb2_event_record(c, c.ev_s1));
    return RB2_OK;
}

// Identity of the work a step would queue: particle count, every array and scratch pointer a kernel receives, the
// scalars passed by value and the scheduling options.  (FNV-1a over the bytes.)
static unsigned long long step_key(const Rb2Ctx &c)
{
    struct {
        int n, cap, pair_mode, sym_min_n, pair_rank, pair_world, sym_tpl, ramo_n_sec, ramo_n_emit, ramo_blocks, ev_cap, redpart_blocks,
            part_begin, part_end, sm_count, pad;
        double sym_waves;
        size_t sym_budget, partial_bytes, bufI_bytes, bufJ_bytes, raw_bytes;
        const void *p[34];
        rb2_config cfg;
    } k;
    memset(&k, 0, sizeof(k));
    k.n = c.n; k.cap = c.cap; k.pair_mode = c.pair_mode; k.sym_min_n = c.sym_min_n; k.pair_rank = c.pair_rank; k.pair_world = c.pair_world;
    k.sym_tpl = c.sym_tpl; k.ramo_n_sec = c.ramo_n_sec; k.ramo_n_emit = c.ramo_n_emit; k.ramo_blocks = c.ramo_blocks; k.ev_cap = c.ev_cap;
    k.redpart_blocks = c.redpart_blocks; k.part_begin = c.part_begin; k.part_end = c.part_end; k.sm_count = c.sm_count;
    k.pad = (c.sym_kmax * 64 + c.sym_gmax) * 2 + (rb2_far_allowed(c) ? 1 : 0);
    k.sym_waves = c.sym_waves; k.sym_budget = c.sym_budget_bytes; k.partial_bytes = c.partial_bytes;
    k.bufI_bytes = c.sym_bufI_bytes; k.bufJ_bytes = c.sym_bufJ_bytes; k.raw_bytes = c.sym_raw_bytes;
    const void *ptrs[] = {c.a.pq, c.a.prev_pos, c.a.vel, c.a.acc, c.a.acc_prev, c.a.acc_prev2, c.a.mass, c.a.species, c.a.step, c.a.emitter,
                          c.a.section, c.a.life, c.a.id, c.b.vel, c.mask, c.evcnt, c.evbits, c.prefix, c.blocksum, c.d_counters, c.d_red,
                          c.d_redpart, c.d_total, c.d_events, c.partial, c.sym_bufI, c.sym_bufJ, c.sym_raw, c.d_ramo_part, c.d_ramo_sec,
                          c.h_ramo_sec, c.h_red, c.sym_units, c.sym_owner};
    static_assert(sizeof(ptrs) / sizeof(ptrs[0]) == 34, "pointer table");
    for (int i = 0; i < 34; ++i) k.p[i] = ptrs[i];
    k.cfg = c.cfg;
    unsigned long long h = 1469598103934665603ull;
    const unsigned char *b = reinterpret_cast<const unsigned char *>(&k);
    for (size_t i = 0; i < sizeof(k); ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h ? h : 1;
}

static int step_finish(rb2_step_result *out);
static int step_queue(void);

int rb2_step(int step, rb2_step_result *out)
{
    (void)step;
    RB2_REQUIRE_INIT();
    // queue on every device first, wait afterwards (rb2_set_devices): the devices integrate their replicas side by side
    // and exchange the partial pair sums inside the finalise kernel
    const int rc = each_device([&]() -> int { return step_queue(); });
    if (rc) {  // nothing to read back; let whatever was queued drain
        char msg[sizeof(g_rb2_err)];
        memcpy(msg, g_rb2_err, sizeof(msg));
        each_device([&]() -> int { cudaStreamSynchronize(g_rb2.stream); return RB2_OK; });
        memcpy(g_rb2_err, msg, sizeof(msg));
        return rc;
    }
    return each_device([&]() -> int { return step_finish(out); });
}

static int step_queue(void)
{
    Rb2Ctx &c = g_rb2;
    {   // the fused step integrates EVERY row: with an i-partition in effect only rows [i_begin, i_end) would get an
        // acceleration and the rest would be integrated with a = 0.  Partitioned runs use the three phases
        // (rb2_update_position, rb2_accel_only + the caller's exchange of the acceleration slices, rb2_update_velocity).
        const int i1 = (c.part_end < 0 || c.part_end > c.n) ? c.n : c.part_end;
        if (c.n > 0 && (c.part_begin > 0 || i1 < c.n))
            return rb2_fail(RB2_ERR_ARG, "rb2_step with the i-partition [%d, %d) of %d particles: the fused step has no exchange; use "
                                         "rb2_update_position / rb2_accel_only + exchange / rb2_update_velocity", c.part_begin, i1, c.n);
    }
    int rc = RB2_OK;
    // The step is ~12 dependent launches; at a few thousand particles their launch gaps are a third of the step.  When
    // two consecutive steps would queue exactly the same work (step_key), the second one is captured into a CUDA graph
    // and replayed from then on.  A run whose particle count changes every step (emission) simply never captures.
    const bool can_graph = c.use_graph && c.p2p_world <= 1 && c.n > 0;
    const unsigned long long key = can_graph ? step_key(c) : 0;
    if (can_graph && c.graph_exec && key == c.graph_key) {
        c.host_events.clear();
        c.accel_timed = c.graph_accel_timed;
        RB2_CUDA(cudaGraphLaunch(c.graph_exec, c.stream));
        c.launches += c.graph_launches;
        c.graph_replays += 1;
    } else if (can_graph && key == c.prev_step_key) {
        if (c.graph_exec) { cudaGraphExecDestroy(c.graph_exec); c.graph_exec = nullptr; }
        const long long l0 = c.launches;
        cudaGraph_t graph = nullptr;
        RB2_CUDA(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
        c.capturing = true;
        rc = queue_step(c);
        c.capturing = false;
        const cudaError_t e = cudaStreamEndCapture(c.stream, &graph);
        if (rc || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            c.use_graph = 0;  // something in the step cannot be captured here: plain launches from now on
            c.launches = l0;
            if ((rc = queue_step(c))) return rc;
        } else {
            const cudaError_t ei = cudaGraphInstantiate(&c.graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ei != cudaSuccess) {
                cudaGetLastError();
                c.graph_exec = nullptr;
                c.use_graph = 0;
                c.launches = l0;
                if ((rc = queue_step(c))) return rc;
            } else {
                c.graph_key = key;
                c.graph_launches = c.launches - l0;
                c.graph_accel_timed = c.accel_timed;
                RB2_CUDA(cudaGraphLaunch(c.graph_exec, c.stream));
            }
        }
    } else {
        if ((rc = queue_step(c))) return rc;
    }
    c.prev_step_key = key;
    return RB2_OK;
}

static int step_finish(rb2_step_result *out)
{
    Rb2Ctx &c = g_rb2;
    int rc = RB2_OK;
    RB2_CUDA(cudaStreamSynchronize(c.stream));
    if ((rc = rb2_p2p_check(c))) return rc;
    rc = finish_position(c, true);
    if (rc) return rc;
    if (out) {
        memset(out, 0, sizeof(*out));
        fill_velocity_result(c, out);
        out->n_events = (int)c.host_events.size();
        out->counts = c.counts;
        if (c.accel_timed) RB2_CUDA(cudaEventElapsedTime(&out->accel_ms, c.ev_a0, c.ev_a1));
        RB2_CUDA(cudaEventElapsedTime(&out->step_ms, c.ev_s0, c.ev_s1));
    }
    return RB2_OK;
}

int rb2_get_ramo_sections(int n_sec, int n_emit, double *out)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    if (!out || n_sec < 1 || n_emit < 1) return rb2_fail(RB2_ERR_ARG, "rb2_get_ramo_sections: bad arguments");
    if (c.ramo_n_sec < 1) return rb2_fail(RB2_ERR_ARG, "per-section Ramo current is off: rb2_set_option(\"ramo_sections\", n) first");
    RB2_CUDA(cudaStreamSynchronize(c.stream));
    for (int e = 0; e < n_emit; ++e)
        for (int s = 0; s < n_sec; ++s)
            out[(size_t)e * n_sec + s] = (c.h_ramo_sec && e < c.ramo_n_emit && s < c.ramo_n_sec) ? c.h_ramo_sec[(size_t)e * c.ramo_n_sec + s] : 0.0;
    return RB2_OK;
}

int rb2_get_events(int max_events, rb2_event *out, int *n_out)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    RB2_CUDA(cudaStreamSynchronize(c.stream));
    const int n = (int)c.host_events.size();
    if (n_out) *n_out = n;
    if (out && max_events > 0) {
        const int m = n < max_events ? n : max_events;
        if (m > 0) memcpy(out, c.host_events.data(), (size_t)m * sizeof(rb2_event));
    }
    return RB2_OK;
}

int rb2_accel_host(int n, const double *pos, const double *charge, const double *mass, double *acc_out)
{
    RB2_REQUIRE_INIT();
    if (n < 0 || n > g_rb2.cap) return rb2_fail(RB2_ERR_CAPACITY, "n = %d exceeds capacity %d", n, g_rb2.cap);
    if (n == 0) return RB2_OK;
    if (!pos || !charge || !mass || !acc_out) return rb2_fail(RB2_ERR_ARG, "NULL argument");
    int inside = -1;  // this call's particle set between the plates?  (scanned once, while the first copies run)
    // every device gets the inputs and queues its share of the pair work; the first one returns the accelerations
    const int rc = each_device([&]() -> int {
        Rb2Ctx &c = g_rb2;
        // scratch = the spare array set; the resident particle state is untouched
        const size_t b3 = (size_t)n * 3 * sizeof(double), b1 = (size_t)n * sizeof(double);
        cudaStream_t st = c.stream;
        RB2_CUDA(cudaMemcpyAsync(c.b.prev_pos, pos, b3, cudaMemcpyHostToDevice, st));
        RB2_CUDA(cudaMemcpyAsync(c.b.mass, charge, b1, cudaMemcpyHostToDevice, st));
        RB2_CUDA(cudaMemcpyAsync(c.b.vel, mass, b1, cudaMemcpyHostToDevice, st));
        int rc1 = rb2_launch_pack(c, c.b.prev_pos, c.b.mass, n, c.b.pq);
        if (rc1) return rc1;
        if (inside < 0) inside = (c.cfg.geometry != RB2_GEOM_PLANAR || rb2_all_between_plates(pos, n, c.cfg.d)) ? 1 : 0;
        const bool state_ok = c.far_state_ok;
        c.far_state_ok = true;  // the resident particles do not take part in this evaluation
        c.far_call_ok = inside == 1;
        // honours the i-partition: this process evaluates and returns rows [i0, i1) only
        int i0 = c.part_begin < 0 ? 0 : c.part_begin;
        int i1 = (c.part_end < 0 || c.part_end > n) ? n : c.part_end;
        if (i0 > i1) i0 = i1;
        rc1 = launch_accel_any(c, c.b.pq, c.b.vel, n, i0, i1, c.b.acc);
        c.far_state_ok = state_ok;
        c.far_call_ok = true;
        if (rc1) return rc1;
        c.accel_timed = (i1 > i0);
        return RB2_OK;
    });
    // the copy-out only once every device has its work queued: into pageable memory it blocks the host until the
    // finalise kernel is through, and that kernel waits for the other devices' partial sums
    const int rc2 = each_device([&]() -> int {
        Rb2Ctx &c = g_rb2;
        int i0 = c.part_begin < 0 ? 0 : c.part_begin;
        int i1 = (c.part_end < 0 || c.part_end > n) ? n : c.part_end;
        if (i0 > i1) i0 = i1;
        if (rc == RB2_OK && i1 > i0 && &c == &g_rb2_all[0])
            RB2_CUDA(cudaMemcpyAsync(acc_out + (size_t)3 * i0, c.b.acc + (size_t)3 * i0, (size_t)(i1 - i0) * 3 * sizeof(double),
                                     cudaMemcpyDeviceToHost, c.stream));
        RB2_CUDA(cudaStreamSynchronize(c.stream));
        return rb2_p2p_check(c);
    });
    return rc ? rc : rc2;
}

static int ensure_field_buffers(Rb2Ctx &c, int M)
{
    if (M > c.fld_cap) {
        if (c.d_pts) RB2_CUDA(cudaFree(c.d_pts));
        if (c.d_fld) RB2_CUDA(cudaFree(c.d_fld));
        if (c.h_pts) RB2_CUDA(cudaFreeHost(c.h_pts));
        if (c.h_fld) RB2_CUDA(cudaFreeHost(c.h_fld));
        c.d_pts = c.d_fld = c.h_pts = c.h_fld = nullptr;
        c.fld_cap = 0;
        const int want = M < 1024 ? 1024 : M + M / 2;  // like the reference's max(M, 1024), src/mod_verlet.F90:1710
        RB2_CUDA(cudaMalloc(&c.d_pts, (size_t)3 * want * sizeof(double)));
        RB2_CUDA(cudaMalloc(&c.d_fld, (size_t)3 * want * sizeof(double)));
        RB2_CUDA(cudaMallocHost(&c.h_pts, (size_t)3 * want * sizeof(double)));
        RB2_CUDA(cudaMallocHost(&c.h_fld, (size_t)3 * want * sizeof(double)));
        c.fld_cap = want;
    }
    return RB2_OK;
}

// Several devices in this process (rb2_set_devices), many points: every device holds the whole particle store, so the
// POINTS are dealt out in contiguous slices -- no reduction (the result of a point agrees with the one-device result to
// rounding: the chunking of the particle range depends on the number of points of a launch).
// surface: E_z only (rb2_field_surface_z), else the three components.
static int field_batch_sharded(int M, const double *pos_in, double *out, bool surface)
{
    const int nd = g_rb2_ndev;
    const int ncomp = surface ? 1 : 3;
    int d = 0;
    int rc = each_device([&]() -> int {
        Rb2Ctx &c = g_rb2;
        const int m0 = (int)((long long)M * d / nd), m1 = (int)((long long)M * (d + 1) / nd);
        ++d;
        const int Md = m1 - m0;
        if (Md < 1) return RB2_OK;
        int rc1 = ensure_field_buffers(c, Md);
        if (rc1) return rc1;
        const size_t b3 = (size_t)3 * Md * sizeof(double);
        memcpy(c.h_pts, pos_in + (size_t)3 * m0, b3);
        RB2_CUDA(cudaMemcpyAsync(c.d_pts, c.h_pts, b3, cudaMemcpyHostToDevice, c.stream));
        rc1 = surface ? rb2_launch_surface_field(c, c.d_pts, Md, c.d_fld) : rb2_launch_field(c, c.a.pq, c.n, nullptr, 0, c.d_pts, Md, c.d_fld);
        if (rc1) return rc1;
        RB2_CUDA(cudaMemcpyAsync(c.h_fld, c.d_fld, (size_t)ncomp * Md * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        return RB2_OK;
    });
    d = 0;
    const int rc2 = each_device([&]() -> int {
        Rb2Ctx &c = g_rb2;
        const int m0 = (int)((long long)M * d / nd), m1 = (int)((long long)M * (d + 1) / nd);
        ++d;
        RB2_CUDA(cudaStreamSynchronize(c.stream));
        if (rc == RB2_OK && m1 > m0) memcpy(out + (size_t)ncomp * m0, c.h_fld, (size_t)ncomp * (m1 - m0) * sizeof(double));
        return RB2_OK;
    });
    return rc ? rc : rc2;
}

int rb2_field_batch_delta(int M, const double *pos_in, int n_new, const double *new_pos, const double *new_charge,
                          double *field_out)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    if (M < 1) return RB2_OK;  // src/mod_verlet.F90:1658
    if (!pos_in || !field_out) return rb2_fail(RB2_ERR_ARG, "NULL argument");
    if (n_new < 0 || (n_new > 0 && (!new_pos || !new_charge))) return rb2_fail(RB2_ERR_ARG, "bad pending-particle arguments");
    if (g_rb2_ndev > 1 && n_new == 0 && M >= 512 * g_rb2_ndev) return field_batch_sharded(M, pos_in, field_out, false);
    int rc = ensure_field_buffers(c, M);
    if (rc) return rc;
    cudaStream_t st = c.stream;
    const size_t b3 = (size_t)3 * M * sizeof(double);
    memcpy(c.h_pts, pos_in, b3);
    RB2_CUDA(cudaMemcpyAsync(c.d_pts, c.h_pts, b3, cudaMemcpyHostToDevice, st));
    if (n_new > 0) {
        if (n_new > c.extra_cap) {
            if (c.d_extra) RB2_CUDA(cudaFree(c.d_extra));
            c.d_extra = nullptr; c.extra_cap = 0;
            const int want = n_new + n_new / 2 + 256;
            RB2_CUDA(cudaMalloc(&c.d_extra, (size_t)want * sizeof(double4)));
            c.extra_cap = want;
        }
        rc = rb2_ensure_stage(c, (size_t)4 * n_new, 0);
        if (rc) return rc;
        RB2_CUDA(cudaMemcpyAsync(c.d_stage_d, new_pos, (size_t)3 * n_new * sizeof(double), cudaMemcpyHostToDevice, st));
        RB2_CUDA(cudaMemcpyAsync(c.d_stage_d + (size_t)3 * n_new, new_charge, (size_t)n_new * sizeof(double), cudaMemcpyHostToDevice, st));
        rc = rb2_launch_pack(c, c.d_stage_d, c.d_stage_d + (size_t)3 * n_new, n_new, c.d_extra);
        if (rc) return rc;
    }
    rc = rb2_launch_field(c, c.a.pq, c.n, c.d_extra, n_new, c.d_pts, M, c.d_fld);
    if (rc) return rc;
    RB2_CUDA(cudaMemcpyAsync(c.h_fld, c.d_fld, b3, cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaStreamSynchronize(st));
    memcpy(field_out, c.h_fld, b3);
    return RB2_OK;
}

int rb2_field_batch(int M, const double *pos_in, double *field_out)
{
    return rb2_field_batch_delta(M, pos_in, 0, nullptr, nullptr, field_out);
}

int rb2_field_surface_z(int M, const double *pos_in, double *Ez_out)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    if (M < 1) return RB2_OK;
    if (!pos_in || !Ez_out) return rb2_fail(RB2_ERR_ARG, "NULL argument");
    if (c.cfg.geometry != RB2_GEOM_PLANAR) return rb2_fail(RB2_ERR_GEOMETRY, "rb2_field_surface_z: planar geometry only");
    for (int k = 0; k < M; ++k)
        if (pos_in[3 * (size_t)k + 2] != 0.0) return rb2_fail(RB2_ERR_ARG, "rb2_field_surface_z: point %d is not on the cathode plane z = 0", k);
    if (g_rb2_ndev > 1 && M >= 512 * g_rb2_ndev) return field_batch_sharded(M, pos_in, Ez_out, true);
    int rc = ensure_field_buffers(c, M);
    if (rc) return rc;
    cudaStream_t st = c.stream;
    const size_t b3 = (size_t)3 * M * sizeof(double);
    memcpy(c.h_pts, pos_in, b3);
    RB2_CUDA(cudaMemcpyAsync(c.d_pts, c.h_pts, b3, cudaMemcpyHostToDevice, st));
    rc = rb2_launch_surface_field(c, c.d_pts, M, c.d_fld);
    if (rc) return rc;
    RB2_CUDA(cudaMemcpyAsync(c.h_fld, c.d_fld, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaStreamSynchronize(st));
    memcpy(Ez_out, c.h_fld, (size_t)M * sizeof(double));
    return RB2_OK;
}

int rb2_mh_planar(const rb2_mh_config *cfg, const double *w_theta, int M, unsigned long long seed, double *df_out,
                  double *F_out, double *pos_out, double *a_rate_io, double *mh_std_io)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    if (M < 1) return RB2_OK;
    if (!cfg || !w_theta || !df_out || !F_out || !pos_out || !a_rate_io || !mh_std_io) return rb2_fail(RB2_ERR_ARG, "NULL argument");
    if (c.cfg.geometry != RB2_GEOM_PLANAR) return rb2_fail(RB2_ERR_GEOMETRY, "rb2_mh_planar: planar geometry only");
    if (cfg->kind != 1 && cfg->kind != 2) return rb2_fail(RB2_ERR_ARG, "rb2_mh_planar: kind must be 1 or 2");
    if (cfg->ndim < 0 || cfg->emit_dim[0] <= 0.0 || cfg->emit_dim[1] <= 0.0) return rb2_fail(RB2_ERR_ARG, "rb2_mh_planar: bad chain setup");
    return rb2_launch_mh_planar(c, cfg, w_theta, M, seed, df_out, F_out, pos_out, a_rate_io, mh_std_io);
}

int rb2_mh_planar_serial(const rb2_mh_config *cfg, const double *w_theta, int M, unsigned long long seed, double *df_out,
                         double *F_out, double *pos_out, int *emit_out, double *a_rate_io, double *mh_std_io)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    if (M < 1) return RB2_OK;
    if (!cfg || !w_theta || !df_out || !F_out || !pos_out || !emit_out || !a_rate_io || !mh_std_io) return rb2_fail(RB2_ERR_ARG, "NULL argument");
    if (c.cfg.geometry != RB2_GEOM_PLANAR) return rb2_fail(RB2_ERR_GEOMETRY, "rb2_mh_planar_serial: planar geometry only");
    if (cfg->kind != 1 && cfg->kind != 2) return rb2_fail(RB2_ERR_ARG, "rb2_mh_planar_serial: kind must be 1 or 2");
    if (cfg->ndim < 0 || cfg->emit_dim[0] <= 0.0 || cfg->emit_dim[1] <= 0.0) return rb2_fail(RB2_ERR_ARG, "rb2_mh_planar_serial: bad chain setup");
    return rb2_launch_mh_planar_serial(c, cfg, w_theta, M, seed, df_out, F_out, pos_out, emit_out, a_rate_io, mh_std_io);
}

int rb2_mh_tip(int M, int ndim, unsigned long long seed, double *eta_f_out, double *df_out, double *pos_out, double *a_rate_io,
               double *mh_std_io)
{
    RB2_REQUIRE_INIT();
    if (M < 0) return rb2_fail(RB2_ERR_ARG, "M < 0");
    if (M == 0) return RB2_OK;
    if (!eta_f_out || !df_out || !pos_out || !a_rate_io || !mh_std_io) return rb2_fail(RB2_ERR_ARG, "NULL argument");
    return rb2_launch_mh_tip(g_rb2, M, ndim, seed, eta_f_out, df_out, pos_out, a_rate_io, mh_std_io);
}

int rb2_planar_supply_level(const rb2_mh_config *cfg, const double *w_theta, int kind, int K, const double *shifts, int n_done,
                            int n_new, double *sums_out, double *ez_sum_out)
{
    RB2_REQUIRE_INIT();
    if (!cfg || !w_theta || !shifts || !sums_out) return rb2_fail(RB2_ERR_ARG, "NULL argument");
    return rb2_planar_supply_level_impl(g_rb2, cfg, w_theta, kind, K, shifts, n_done, n_new, sums_out, ez_sum_out);
}

int rb2_tip_supply_set_grid(int M, const double *pts, const double *normals, const double *area)
{
    RB2_REQUIRE_INIT();
    if (M < 1) return rb2_fail(RB2_ERR_ARG, "rb2_tip_supply_set_grid: M < 1");
    if (!pts || !normals || !area) return rb2_fail(RB2_ERR_ARG, "NULL argument");
    return rb2_tip_supply_set_grid_impl(g_rb2, M, pts, normals, area);
}

int rb2_tip_supply(double *n_s_out, double *F_sum_out)
{
    RB2_REQUIRE_INIT();
    if (!n_s_out) return rb2_fail(RB2_ERR_ARG, "NULL argument");
    return rb2_tip_supply_impl(g_rb2, n_s_out, F_sum_out);
}

int rb2_field_window_open(void) { RB2_REQUIRE_INIT(); return RB2_OK; }
int rb2_field_window_close(void) { RB2_REQUIRE_INIT(); return RB2_OK; }

static int set_option_one(const char *name, double value)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    if (!name) return rb2_fail(RB2_ERR_ARG, "NULL option name");
    if (!strcmp(name, "pair_mode")) {
        if (value < 0 || value > 2) return rb2_fail(RB2_ERR_ARG, "pair_mode must be 0 (auto), 1 (gather) or 2 (pair-symmetric)");
        c.pair_mode = (int)value;
    } else if (!strcmp(name, "sym_min_n")) {
        c.sym_min_n = (int)value;
    } else if (!strcmp(name, "event_buffer")) {
        if (value < 1) return rb2_fail(RB2_ERR_ARG, "event_buffer must be >= 1 record");
        c.ev_min = (int)value;
        if (c.d_events) { RB2_CUDA(cudaStreamSynchronize(c.stream)); RB2_CUDA(cudaFree(c.d_events)); }
        c.d_events = nullptr;
        c.ev_cap = 0;
    } else if (!strcmp(name, "ramo_sections") || !strcmp(name, "ramo_emitters")) {
        const bool sec = !strcmp(name, "ramo_sections");
        if (value < (sec ? 0 : 1) || value > (sec ? 96 * 96 : 8)) return rb2_fail(RB2_ERR_ARG, "%s out of range", name);
        const int ns = sec ? (int)value : c.ramo_n_sec, ne = sec ? c.ramo_n_emit : (int)value;
        if ((size_t)ns * ne * sizeof(double) > 200 * 1024) return rb2_fail(RB2_ERR_ARG, "ramo_sections x ramo_emitters exceeds the shared-memory table (25600 entries)");
        RB2_CUDA(cudaStreamSynchronize(c.stream));
        cudaFree(c.d_ramo_part); cudaFree(c.d_ramo_sec); cudaFreeHost(c.h_ramo_sec);
        c.d_ramo_part = c.d_ramo_sec = c.h_ramo_sec = nullptr;
        c.ramo_n_sec = ns; c.ramo_n_emit = ne;
    } else if (!strcmp(name, "step_graph")) {
        c.use_graph = (value != 0.0) ? 1 : 0;
        if (c.graph_exec) { RB2_CUDA(cudaStreamSynchronize(c.stream)); cudaGraphExecDestroy(c.graph_exec); c.graph_exec = nullptr; }
        c.graph_key = c.prev_step_key = 0;
    } else if (!strcmp(name, "sym_tpl")) {
        if (value != 0 && value != 1 && value != 2) return rb2_fail(RB2_ERR_ARG, "sym_tpl (targets per lane) must be 0 (auto), 1 or 2");
        c.sym_tpl = (int)value;
    } else if (!strcmp(name, "mh_ctas_per_sm")) {
        if (value < 1 || value > 4) return rb2_fail(RB2_ERR_ARG, "mh_ctas_per_sm must be 1..4");
        c.mh_ctas_per_sm = (int)value;
    } else if (!strcmp(name, "tip_field_small")) {
        c.tip_field_small = (value != 0.0) ? 1 : 0;
    } else if (!strcmp(name, "mh_small_max")) {
        if (value < 1 || value > 512) return rb2_fail(RB2_ERR_ARG, "mh_small_max must be 1..512");
        c.mh_small_max = (int)value;
    } else if (!strcmp(name, "mh_small")) {
        c.mh_small = (value != 0.0) ? 1 : 0;
    } else if (!strcmp(name, "sym_far")) {
        c.sym_far = (value != 0.0) ? 1 : 0;
    } else if (!strcmp(name, "sym_kmax")) {
        if (value < 1 || value > 4096) return rb2_fail(RB2_ERR_ARG, "sym_kmax must be 1..4096");
        c.sym_kmax = (int)value;
    } else if (!strcmp(name, "sym_gmax")) {
        if (value < 1 || value > 24) return rb2_fail(RB2_ERR_ARG, "sym_gmax must be 1..24");
        c.sym_gmax = (int)value;
    } else if (!strcmp(name, "sym_waves")) {
        if (value < 1) return rb2_fail(RB2_ERR_ARG, "sym_waves must be >= 1");
        c.sym_waves = value;
    } else if (!strcmp(name, "sym_budget_mb")) {
        if (value <= 0) return rb2_fail(RB2_ERR_ARG, "sym_budget_mb must be > 0");
        c.sym_budget_bytes = (size_t)(value * 1048576.0);
    } else {
        return rb2_fail(RB2_ERR_ARG, "unknown option '%s'", name);
    }
    return RB2_OK;
}
int rb2_set_option(const char *name, double value)
{
    return each_device([&]() -> int { return set_option_one(name, value); });
}

int rb2_set_pair_rank(int rank, int world)
{
    RB2_REQUIRE_INIT();
    if (world < 1 || rank < 0 || rank >= world) return rb2_fail(RB2_ERR_ARG, "bad rank %d of %d", rank, world);
    if (g_rb2_ndev > 1) return rb2_fail(RB2_ERR_ARG, "rb2_set_pair_rank: the devices of this process already share the pair work (rb2_set_devices)");
    g_rb2.pair_rank = rank;
    g_rb2.pair_world = world;
    return RB2_OK;
}

int rb2_accel_partial(void)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    if (c.cfg.geometry != RB2_GEOM_PLANAR) return rb2_fail(RB2_ERR_GEOMETRY, "rb2_accel_partial: planar geometry only");
    if (g_rb2_ndev > 1) return rb2_fail(RB2_ERR_ARG, "rb2_accel_partial / rb2_accel_finalize are the multi-process plumbing: not with rb2_set_devices");
    int rc = rb2_launch_accel_sym_partial(c, c.a.pq, c.n);
    if (rc) return rc;
    c.last_pair_kernel = 2;
    RB2_CUDA(cudaStreamSynchronize(c.stream));
    return RB2_OK;
}

int rb2_accel_finalize(void)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    if (c.n > 0 && (!c.sym_raw_cur || c.sym_n_pad < c.n)) return rb2_fail(RB2_ERR_ARG, "rb2_accel_finalize without rb2_accel_partial");
    int rc = rb2_launch_accel_sym_finalize(c, c.a.pq, c.a.mass, c.n, c.a.acc);
    if (rc) return rc;
    c.accel_timed = c.n > 0;
    RB2_CUDA(cudaStreamSynchronize(c.stream));
    return rb2_p2p_check(c);
}

static int set_partition_one(int i_begin, int i_end)
{
    RB2_REQUIRE_INIT();
    if (i_begin < 0 || (i_end >= 0 && i_end < i_begin)) return rb2_fail(RB2_ERR_ARG, "bad partition [%d, %d)", i_begin, i_end);
    g_rb2.part_begin = i_begin;
    g_rb2.part_end = i_end;
    return RB2_OK;
}
int rb2_set_partition(int i_begin, int i_end)
{
    return each_device([&]() -> int { return set_partition_one(i_begin, i_end); });
}

int rb2_device_buffer(const char *name, void **dev_ptr, size_t *bytes)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    if (!name || !dev_ptr) return rb2_fail(RB2_ERR_ARG, "NULL argument");
    const size_t cap = (size_t)c.cap;
    void *p = nullptr;
    size_t b = 0;
    if (!strcmp(name, "acc")) { p = c.a.acc; b = 3 * cap * sizeof(double); }
    else if (!strcmp(name, "pq")) { p = c.a.pq; b = cap * sizeof(double4); }
    else if (!strcmp(name, "vel")) { p = c.a.vel; b = 3 * cap * sizeof(double); }
    else if (!strcmp(name, "acc_prev")) { p = c.a.acc_prev; b = 3 * cap * sizeof(double); }
    else if (!strcmp(name, "acc_prev2")) { p = c.a.acc_prev2; b = 3 * cap * sizeof(double); }
    else if (!strcmp(name, "mass")) { p = c.a.mass; b = cap * sizeof(double); }
    else if (!strcmp(name, "raw")) { p = c.sym_raw_cur; b = (size_t)3 * c.sym_n_pad * sizeof(double); }
    else return rb2_fail(RB2_ERR_ARG, "unknown buffer '%s'", name);
    *dev_ptr = p;
    if (bytes) *bytes = b;
    return RB2_OK;
}

static int synchronize_one(void)
{
    RB2_REQUIRE_INIT();
    RB2_CUDA(cudaStreamSynchronize(g_rb2.stream));
    return RB2_OK;
}
int rb2_synchronize(void)
{
    return each_device([&]() -> int { return synchronize_one(); });
}

int rb2_stream(void **stream_out)
{
    RB2_REQUIRE_INIT();
    if (!stream_out) return rb2_fail(RB2_ERR_ARG, "NULL argument");
    *stream_out = (void *)g_rb2.stream;
    return RB2_OK;
}

int rb2_fp64_peak(double ms_target, double *tflops_out, float *ms_out)
{
    RB2_REQUIRE_INIT();
    if (!tflops_out) return rb2_fail(RB2_ERR_ARG, "NULL argument");
    if (ms_target <= 0.0) ms_target = 50.0;
    return rb2_launch_fp64_peak(g_rb2, ms_target, tflops_out, ms_out);
}

int rb2_launch_count(long long *out, int reset)
{
    RB2_REQUIRE_INIT();
    if (out) *out = g_rb2.launches;
    if (reset) g_rb2.launches = 0;
    return RB2_OK;
}

int rb2_get_stat(const char *name, double *out)
{
    RB2_REQUIRE_INIT();
    if (!name || !out) return rb2_fail(RB2_ERR_ARG, "NULL argument");
    if (!strcmp(name, "graph_replays")) *out = (double)g_rb2.graph_replays;
    else if (!strcmp(name, "graph_launches")) *out = (double)g_rb2.graph_launches;
    else if (!strcmp(name, "launches")) *out = (double)g_rb2.launches;
    else if (!strcmp(name, "sym_plans")) *out = (double)g_rb2.sym_plans;
    else if (!strcmp(name, "sym_plan_ms")) *out = g_rb2.sym_plan_ms;
    else return rb2_fail(RB2_ERR_ARG, "unknown counter '%s'", name);
    return RB2_OK;
}

int rb2_last_accel_info(float *ms, int *grid_x, int *grid_y, int *block, int *j_split)
{
    RB2_REQUIRE_INIT();
    Rb2Ctx &c = g_rb2;
    if (ms) {
        *ms = 0.f;
        if (c.accel_timed) {
            RB2_CUDA(cudaEventSynchronize(c.ev_a1));
            RB2_CUDA(cudaEventElapsedTime(ms, c.ev_a0, c.ev_a1));
        }
    }
    if (grid_x) *grid_x = c.last_grid_x;
    if (grid_y) *grid_y = c.last_grid_y;
    if (block) *block = c.last_block;
    if (j_split) *j_split = c.last_split;
    return RB2_OK;
}

}  // extern "C"

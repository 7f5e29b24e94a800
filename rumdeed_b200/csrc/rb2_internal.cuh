// rb2_internal.cuh -- shared declarations of librumdeed_b200.so (not installed).
//
// Data layout in HBM (one allocation per array, capacity = rb2_config.capacity; every array
// exists twice so that the stable compaction of Remove_Particles is one gather pass into
// the spare set followed by a pointer swap):
//   pq        double4[cap]   {x, y, z, q}: the only thing the O(N^2) kernels read
//                            (32 B/particle, 16 B aligned rows for TMA bulk copies)
//   prev_pos, vel, acc, acc_prev, acc_prev2   double[3*cap]  Fortran (3,N) order
//   mass      double[cap]
//   species, step, emitter, section, life, id   int[cap]
//   mask      int[cap]       1 = live, 0 = marked for removal
//   evcnt     uint8[cap]     records this particle produced in the position update
//   evbits    uint16[cap]    bit 0/1: absorbed top/bot record, bit 2+k: plane k crossed
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <vector>

#include "rumdeed_b200.h"

// ---- physical constants (reference src/mod_global.F90:26-75, :333) -------------------
#define RB2_PI 3.141592653589793238462643383279502884197169399375105820974944592307816406286
namespace rb2k {
constexpr double mu_0 = 1.25663706212e-6;
constexpr double c = 299792458.0;
constexpr double epsilon_0 = 1.0 / (mu_0 * (c * c));  // derived, like the reference
constexpr double div_fac_c = 1.0 / (4.0 * RB2_PI * epsilon_0 * 1.0);
constexpr double q_0 = 1.602176634e-19;
constexpr double m_0 = 9.1093837015e-31;
constexpr double m_u = 1.66053906660e-27;
constexpr double m_N2 = 28.0134 * m_u;
constexpr double m_N2p = m_N2 - m_0;
constexpr double length_scale = 1.0e-9;
constexpr double soft = length_scale * length_scale;  // 1e-18 m, added to r
}  // namespace rb2k

// ---- kernel parameter blocks (passed by value) -----------------------------------------
struct PlanarParams {
    double two_d;  // 2*d
    double E_z;
    int    nic;    // N_ic_max
    int    do_ic;
    int    far_ok; // d >= 1 um: the far image partners may skip the softening term (rb2_inv_r3_far)
};
struct TipParams {
    double a_foci, shift_z, pre_fac_E_tip, eta_1;
    double z_0, r_tip;  // sphere centre height (h_tip - r_tip) and radius
    double unit_scale_num, unit_scale_den;  // pre_fac_E_tip_unit_voltage, pre_fac_E_tip
    int    do_ic;
};
struct StepParams {
    int    geometry;
    double dt, dt2;
    double box_z;
    double d;
    int    planes_N;
    double planes_z[RB2_PLANES_MAX];
    PlanarParams pl;
    TipParams    tip;
};

struct DevCounters {  // cumulative since the last rb2_remove_marked
    int mark_part, mark_elec, mark_ion, mark_atom;
    int top_part, bot_part, top_elec, bot_elec, top_ion, bot_ion;
    int n_events, pad;
    int recom_part, recom_elec, recom_ion, pad2;  // remove_recom marks (collisions)
    int ion_part, ion_elec, ion_atom, pad3;       // remove_ion marks (src/mod_pair.F90:273-279, :327-333)
};

struct DevArrays {
    double4 *pq = nullptr;
    double *prev_pos = nullptr, *vel = nullptr, *acc = nullptr, *acc_prev = nullptr, *acc_prev2 = nullptr;
    double *mass = nullptr;
    int *species = nullptr, *step = nullptr, *emitter = nullptr, *section = nullptr, *life = nullptr, *id = nullptr;
};

#define RB2_P2P_MAX 16

struct Rb2Ctx {
    bool         init = false;
    rb2_config   cfg{};
    int          dev = 0;
    int          sm_count = 148;
    cudaStream_t stream = nullptr;
    int          cap = 0;
    int          n = 0;  // nrPart
    rb2_counts   counts{};
    int          part_begin = 0, part_end = -1;  // i-partition (end<0: all)

    DevArrays a, b;  // current / spare
    int *mask = nullptr;
    unsigned char *evcnt = nullptr;
    unsigned short *evbits = nullptr;
    int *prefix = nullptr, *blocksum = nullptr;  // scan scratch
    unsigned long long *life_hist = nullptr;     // [(RB2_MAX_LIFE_TIME+1)*4]

    DevCounters *d_counters = nullptr, *h_counters = nullptr;  // device / pinned host
    double *d_red = nullptr, *h_red = nullptr;                  // velocity-update reductions (16 doubles)
    double *d_redpart = nullptr; int redpart_blocks = 0;
    // ramo_current_emit(sec, emit) (options "ramo_sections" / "ramo_emitters"; 0 sections = off)
    int    ramo_n_sec = 0, ramo_n_emit = 1, ramo_blocks = 0;
    double *d_ramo_part = nullptr, *d_ramo_sec = nullptr, *h_ramo_sec = nullptr;
    int *d_total = nullptr, *h_total = nullptr;                 // scan totals

    double *partial = nullptr; size_t partial_bytes = 0;       // pair-kernel partial sums
    // pair-symmetric kernel (rb2_pair_sym.cu)
    int    pair_mode = 0;                 // 0 auto, 1 gather, 2 pair-symmetric
    int    sym_min_n = 3500;              // auto: use the symmetric kernel from this N on (tools/sym_crossover.py)
    int    pair_rank = 0, pair_world = 1; // ownership of (target, group) CTAs across processes
    size_t sym_budget_bytes = (size_t)2048 << 20;  // scratch for the (set, source tile) / (group, target) partial sums
    int    sym_tpl = 0;                   // targets per lane of the pair-symmetric kernel: 0 auto, 1, 2
    int    sym_far = 1;                   // option "sym_far": allow the cheaper far-partner inverse cube when d >= 1 um
    // ... which also needs every charged particle between the plates (0 <= z <= d): true for anything the time step
    // produces (a particle that leaves the gap is marked and loses its charge in the same kernel), checked on the host
    // for whatever the caller hands in (rb2_upload_particles, rb2_add_particles, rb2_accel_host)
    bool   far_state_ok = true;           // the resident particles
    bool   far_call_ok = true;            // the particle set of the rb2_accel_host call in flight
    int    sym_kmax = 12, sym_gmax = 24;  // caps of the work-unit shape (options "sym_kmax", "sym_gmax"; tools/sym_unit_sweep.py)
    double sym_waves = 64.0;              // work units are sized for about this many waves per band launch and rank (tools/run_r2_2gpu_waves.sh)
    double *sym_bufI = nullptr, *sym_bufJ = nullptr, *sym_raw = nullptr;
    size_t sym_bufI_bytes = 0, sym_bufJ_bytes = 0, sym_raw_bytes = 0;
    int    sym_n_pad = 0;
    double *sym_raw_cur = nullptr;        // partial sums of the current evaluation (sym_raw, or a slot of the exchange block)
    unsigned char *sym_owner = nullptr; size_t sym_owner_cap = 0;  // cost-balanced deal of the work units to the ranks
    unsigned long long sym_owner_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // ... valid for (tiles, superblocks, T, K, G, band width, world, rank)
    int2 *sym_units = nullptr; size_t sym_units_cap = 0;           // this rank's units (set, group), all bands, in dealing order
    std::vector<size_t> sym_unit_off; std::vector<int> sym_unit_cnt;  // ... per band
    long long sym_plans = 0; double sym_plan_ms = 0.0;               // how often the unit list was rebuilt, host time spent on it
    // peer-memory exchange (rb2_p2p.cu)
    void  *p2p_local = nullptr;           // this rank's exchange block (exported over CUDA IPC)
    void  *p2p_peer[RB2_P2P_MAX] = {};    // every rank's block as mapped here ([rank] == p2p_local)
    int    p2p_world = 0, p2p_npad_max = 0;
    unsigned long long p2p_epoch = 0;     // evaluations since attach; parity selects the partial-sum slot
    bool   p2p_ipc = false;               // peers mapped through CUDA IPC (one process per GPU) -- else plain peer access
    unsigned sym_attr_mask = 0;           // k_pair_sym instantiations whose shared-memory limit has been raised on this device
    volatile int *p2p_err = nullptr;      // mapped host word the finalise kernel reports a failed exchange in
    int   *p2p_err_dev = nullptr;         // ... its device address
    int    last_pair_kernel = 0;          // 1 gather, 2 symmetric
    rb2_event *d_events = nullptr; int ev_cap = 0;
    int    ev_min = 65536;                // initial size of the record buffer (option "event_buffer")
    int    mh_ctas_per_sm = 2;            // sampler: CTAs per SM of the cooperative kernel (option "mh_ctas_per_sm", 1..4)
    int    tip_field_small = 1;           // tip field: CTA-per-point kernel for small batches (option "tip_field_small")
    int    mh_small = 1;                  // sampler: single-barrier kernel for few chains (option "mh_small", 0 = off)
    int    mh_small_max = 512;            // ... up to this many chains (option "mh_small_max", <= 512)
    std::vector<rb2_event> host_events;

    // staging (grow-only): field points / fields, add/mark arguments, rb2_accel_host
    double *d_pts = nullptr, *d_fld = nullptr, *h_pts = nullptr, *h_fld = nullptr; int fld_cap = 0;
    double4 *d_extra = nullptr; int extra_cap = 0;
    double *d_stage_d = nullptr; size_t stage_d_cap = 0;  // doubles
    int *d_stage_i = nullptr; size_t stage_i_cap = 0;     // ints
    double *h_stage = nullptr; size_t h_stage_bytes = 0;  // pinned
    double *d_supq = nullptr, *h_supq = nullptr; size_t supq_cap = 0;  // planar supply quadrature: table, nodes, E_z, partial sums
    double4 *d_tip_img = nullptr; size_t tip_img_cap = 0;           // tip accelerations: sphere image of every particle
    double *d_sup_grid = nullptr, *h_sup = nullptr; int sup_M = 0;  // tip supply grid: nodes, normals, areas, fields, partial sums

    cudaEvent_t ev_a0 = nullptr, ev_a1 = nullptr, ev_s0 = nullptr, ev_s1 = nullptr, ev_c = nullptr;
    // CUDA graph of the fused step (rb2_step): captured the second time in a row that a step would queue exactly the same
    // launches (same particle count, arrays, scalars, scratch) and replayed while that stays so (option "step_graph")
    int   use_graph = 1;
    bool  capturing = false;
    unsigned long long graph_key = 0, prev_step_key = 0;
    cudaGraphExec_t graph_exec = nullptr;
    long long graph_launches = 0, graph_replays = 0;
    bool  graph_accel_timed = false;
    bool  accel_timed = false;
    int   last_grid_x = 0, last_grid_y = 0, last_block = 0, last_split = 0;
    long long launches = 0;
};

// One context per device.  The classic set-up has one (one GPU per process); rb2_set_devices adds a replica of the whole
// particle state on each further device of THIS process: every state-changing call is made on all of them, the pair
// work of the pair-symmetric kernel is split over them and exchanged through peer memory (rb2_p2p.cu), everything that
// only reads the state (field batches, samplers, downloads) uses the first one.  g_rb2 is the context the current call
// works on.
extern Rb2Ctx  g_rb2_all[RB2_P2P_MAX];
extern Rb2Ctx *g_rb2_cur;
extern int     g_rb2_ndev;
#define g_rb2 (*g_rb2_cur)
extern char   g_rb2_err[512];

int rb2_fail(int code, const char *fmt, ...);

#define RB2_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return rb2_fail(RB2_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)
#define RB2_REQUIRE_INIT()                                                     \
    do {                                                                       \
        if (!g_rb2.init) return rb2_fail(RB2_ERR_NOT_INIT, "rb2_init has not been called"); \
    } while (0)
#define RB2_LAUNCHED(n_) (g_rb2.launches += (n_))

StepParams rb2_make_step_params(const rb2_config &c);
// event record that also works while the step is being captured into a graph (a timing event needs an explicit node)
inline cudaError_t rb2_event_record(Rb2Ctx &c, cudaEvent_t e)
{
    return c.capturing ? cudaEventRecordWithFlags(e, c.stream, cudaEventRecordExternal) : cudaEventRecord(e, c.stream);
}
int rb2_ensure_stage(Rb2Ctx &ctx, size_t n_doubles, size_t n_ints);
inline bool rb2_far_allowed(const Rb2Ctx &c) { return c.sym_far && c.far_state_ok && c.far_call_ok; }
inline bool rb2_all_between_plates(const double *pos3, int n, double d)
{
    for (int k = 0; k < n; ++k) {
        const double z = pos3[3 * (size_t)k + 2];
        if (!(z >= 0.0 && z <= d)) return false;
    }
    return true;
}
int rb2_tip_supply_set_grid_impl(Rb2Ctx &ctx, int M, const double *pts, const double *nrm, const double *area);
int rb2_tip_supply_impl(Rb2Ctx &ctx, double *n_s_out, double *F_sum_out);
int rb2_planar_supply_level_impl(Rb2Ctx &ctx, const rb2_mh_config *cfg, const double *w_theta_host, int kind, int K,
                                 const double *shifts, int n_done, int n_new, double *sums_out, double *ez_sum_out);

// pair / field kernels (rb2_pair.cu)
int rb2_launch_accel(Rb2Ctx &ctx, const double4 *pq, const double *mass, int n, int i_begin, int i_end, double *acc_out);
int rb2_launch_field(Rb2Ctx &ctx, const double4 *pq, int n, const double4 *extra, int n_extra,
                     const double *d_pts, int M, double *d_fld);
// device-resident Metropolis-Hastings sampler (rb2_mh.cu)
int rb2_launch_mh_planar(Rb2Ctx &ctx, const rb2_mh_config *cfg, const double *w_theta_host, int M, unsigned long long seed,
                         double *df_out, double *F_out, double *pos_out, double *a_rate_io, double *mh_std_io);
int rb2_launch_surface_field(Rb2Ctx &ctx, const double *d_pts, int M, double *d_Ez);
int rb2_launch_mh_planar_serial(Rb2Ctx &ctx, const rb2_mh_config *cfg, const double *w_theta_host, int M, unsigned long long seed,
                                double *df_out, double *F_out, double *pos_out, int *emit_out, double *a_rate_io, double *mh_std_io);
int rb2_launch_mh_tip(Rb2Ctx &ctx, int M, int ndim, unsigned long long seed, double *eta_f_out, double *df_out, double *pos_out,
                      double *a_rate_io, double *mh_std_io);
// collisions (rb2_collisions.cu)
void rb2_collisions_release(Rb2Ctx &ctx);
int rb2_fetch_counters(Rb2Ctx &ctx);  // device counters -> ctx.counts (rb2_api.cu)
// nearest-electron sweep (rb2_nearest.cu)
int rb2_launch_nearest(Rb2Ctx &ctx, double *d_dist, int *d_id);
// peer-memory exchange (rb2_p2p.cu)
double *rb2_p2p_begin_evaluation(Rb2Ctx &ctx, int n_pad);
int rb2_launch_accel_sym_exchange_finalize(Rb2Ctx &ctx, const double4 *pq, const double *mass, int n, double *acc_out);
int rb2_p2p_release(Rb2Ctx &ctx);
int rb2_p2p_link_local(Rb2Ctx *all, int n);  // rb2_set_devices: exchange blocks + peer pointers inside one process
int rb2_p2p_check(Rb2Ctx &ctx);  // after a stream synchronisation: RB2_ERR_CUDA when the last exchange failed
// pair-symmetric kernel (rb2_pair_sym.cu)
int rb2_launch_accel_sym_partial(Rb2Ctx &ctx, const double4 *pq, int n);
int rb2_launch_accel_sym_finalize(Rb2Ctx &ctx, const double4 *pq, const double *mass, int n, double *acc_out);
// integrate kernels (rb2_integrate.cu)
int rb2_launch_pack(Rb2Ctx &ctx, const double *pos3, const double *q, int n, double4 *pq);
int rb2_launch_unpack(Rb2Ctx &ctx, const double4 *pq, int n, double *pos3, double *q);
int rb2_launch_update_position(Rb2Ctx &ctx);
int rb2_launch_events(Rb2Ctx &ctx);
int rb2_rebuild_events(Rb2Ctx &ctx, int n_events, bool after_velocity_update);
int rb2_launch_update_velocity(Rb2Ctx &ctx);
int rb2_launch_ramo_sections(Rb2Ctx &ctx);
int rb2_launch_compact(Rb2Ctx &ctx, int step);
int rb2_launch_add(Rb2Ctx &ctx, int k, int slot0, int id0, int step, const double *d_pos, const double *d_vel,
                   const int *d_species, const int *d_emit, const int *d_sec, const int *d_life);
int rb2_launch_mark(Rb2Ctx &ctx, int k, const int *d_index, const int *d_reason);
int rb2_launch_fill_defaults(Rb2Ctx &ctx, int n, bool species, bool step, bool emitter, bool section, bool life, bool id);
int rb2_launch_fill_mask(Rb2Ctx &ctx, int n);
int rb2_launch_fp64_peak(Rb2Ctx &ctx, double ms_target, double *tflops, float *ms);

// ---- device math ---------------------------------------------------------------------------
#ifdef __CUDACC__
// Added to dx^2 before the other squares are accumulated: keeps every squared distance
// strictly positive, so that exactly coincident points (s == 0) get a finite rsqrt seed and
// weight, which the zero offset then multiplies to exactly 0 -- the value the reference's
// softened 1/r^3 gives.  1e-60 m^2 is 36 orders of magnitude below any physical s here
// (it never changes a rounded result) and costs nothing: the DMUL becomes a DFMA.
// (An integer clamp of the seed's high word was measured 7% slower: ALU and MUFU
// instructions share the FP64 pipe's dispatch port on sm_100, see DESIGN.md.)
#define RB2_S_FLOOR 1.0e-60

__device__ __forceinline__ double rb2_rsqrt_seed(double s)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));  // MUFU.RSQ64H
    return y;
}

// w = 1 / (sqrt(s) + 1e-18)^3 for s > 0  (reference: r = sqrt(..) + length_scale**2;
// inv_r3 = 1/(r*r*r), src/mod_verlet.F90:1302-1303, src/acc_ic_planar_series.inc:22-23).
// 7 FP64-pipe instructions + 1 MUFU instead of sqrt + add + 2 mul + divide:
//   y0 ~ s^-1/2 (>= 18 good bits), e = 1 - s*y0^2,
//   s^-3/2 = y0^3 (1-e)^-3/2 = y0^3 (1 + 3/2 e + 15/8 e^2 + O(e^3)),   |e^3| < 1e-15
//   (r+eps)^-3 = r^-3 (1 - 3 eps/r + O((eps/r)^2)),  eps/r <= 1e-6 for r >= 1e-12 m
__device__ __forceinline__ double rb2_inv_r3_soft(double s)
{
    const double y0 = rb2_rsqrt_seed(s);
    const double t = y0 * y0;
    const double e = fma(-s, t, 1.0);
    const double c0 = fma(-3.0 * rb2k::soft, y0, 1.0);
    const double p = fma(e, fma(1.875, e, 1.5), c0);
    return (y0 * t) * p;
}

// The same without the softening term: for partners that are at least a gap width d away (the n = +-1 images other than
// the one behind the anode) 3 eps / r <= 3e-18 / d, i.e. <= 3e-12 for d >= 1 um -- below the 1e-11 bar even if such
// partners carried the whole force; one FP64 instruction less each.  Only used when PlanarParams.far_ok says so.
__device__ __forceinline__ double rb2_inv_r3_far(double s)
{
    const double y0 = rb2_rsqrt_seed(s);
    const double t = y0 * y0;
    const double e = fma(-s, t, 1.0);
    const double p = fma(e, fma(1.875, e, 1.5), 1.0);
    return (y0 * t) * p;
}

// Softened and plain inverse cube of the same s from one seed (the tip's Coulomb term and its sphere-image term share
// their distance): 8 FP64 instructions for both.
__device__ __forceinline__ void rb2_inv_r3_both(double s, double &w_soft, double &w_plain)
{
    const double y0 = rb2_rsqrt_seed(s);
    const double t = y0 * y0;
    const double e = fma(-s, t, 1.0);
    const double p0 = fma(e, fma(1.875, e, 1.5), 1.0);
    const double p1 = fma(-3.0 * rb2k::soft, y0, p0);
    const double y3 = y0 * t;
    w_plain = y3 * p0;
    w_soft = y3 * p1;
}

// The first-order softening above is good to 6 (eps/r)^2: 6e-14 at r = 1e-11 m, but 6e-10 at 1e-13 m.  Pairs whose
// LATERAL offset is below 1e-11 m -- the only ones in which any of the distances of the pair (direct, or to an image
// partner when both particles sit at an electrode) can be that short -- are flagged in the fast loops (one compare per
// pair) and re-evaluated with the reference's own operations: r = sqrt(s) + length_scale**2; 1/(r*r*r)
// (src/mod_verlet.F90:1302-1303), IEEE square root and divide.
#define RB2_CLOSE_DXY2 1.0e-22
#ifndef RB2_CLOSE_INT
#define RB2_CLOSE_INT 1  // measured at N = 1e6 with the slow path out of line (tools/build_variants.sh, two runs): no flag 2552 / 2555 ms,
#endif                   // integer compare of the high word 2538 / 2541 ms, FP64 compare 2569 / 2572 ms; N = 1e4: 0.315 / 0.315 / 0.319 ms
__device__ __forceinline__ bool rb2_is_close(double dxy2)
{
#ifdef RB2_CLOSE_OFF  // measurement only (tools/build_variants.sh): the fast path alone
    return false;
#elif RB2_CLOSE_INT
    // positive doubles order like their high words; an integer compare keeps the FP64 pipe free
    return __double2hiint(dxy2) < 0x3B5E392A;  // high word of 1.0e-22 (0x3B5E392010175EE6), rounded up
#else
    return dxy2 < RB2_CLOSE_DXY2;
#endif
}
__device__ __forceinline__ double rb2_inv_r3_exact(double s)
{
    const double r = __dadd_rn(__dsqrt_rn(s), rb2k::soft);
    return __ddiv_rn(1.0, __dmul_rn(__dmul_rn(r, r), r));
}
template <bool EXACT>
__device__ __forceinline__ double rb2_inv_r3_sel(double s)
{
    return EXACT ? rb2_inv_r3_exact(s) : rb2_inv_r3_soft(s);
}

// Vacuum field of the hyperboloid tip, src/acc_tip_field_E.inc:13-34 ==
// field_E_Hyperboloid, src/mod_hyperboloid_tip.f90:115-154.
__device__ __forceinline__ void rb2_tip_field_E(const TipParams &T, double x_1, double y_1, double z_1, double &fE_x,
                                                double &fE_y, double &fE_z)
{
    const double zp = z_1 + T.a_foci - T.shift_z, zm = z_1 - T.a_foci - T.shift_z;
    const double r_p = sqrt(x_1 * x_1 + y_1 * y_1 + zp * zp);
    const double r_m = sqrt(x_1 * x_1 + y_1 * y_1 + zm * zm);
    const double xi = (r_p + r_m) / (2.0 * T.a_foci);
    const double eta = (r_p - r_m) / (2.0 * T.a_foci);
    double phi;
    if ((fabs(x_1) < 1.0e-18) && (fabs(y_1) < 1.0e-18)) phi = 0.0;
    else phi = atan2(y_1, x_1);
    const double pre_fac_xyz = T.pre_fac_E_tip * 1.0 / (xi * xi - eta * eta);
    double fac_xy;
    if (fabs(xi - 1.0) < 1.0e-6) fac_xy = 0.0;
    else fac_xy = eta * sqrt((xi * xi - 1.0) / (1.0 - eta * eta));
    fE_x = -1.0 * pre_fac_xyz * fac_xy * cos(phi);
    fE_y = -1.0 * pre_fac_xyz * fac_xy * sin(phi);
    fE_z = pre_fac_xyz * xi;
}
#endif

/*
 * rumdeed_b200.h -- C ABI of the B200-native replacement for RUMDEED's
 * per-timestep hot path (all-pairs Coulomb + image charges, Beeman step,
 * batched surface-field evaluation).
 *
 * This is the drop-in boundary: every entry point replaces one call site of the
 * reference Fortran host (cited per function, paths relative to the RUMDEED
 * tree) and is what `fortran/mod_b200_bridge.F90` binds with ISO_C_BINDING
 * (see INTEGRATION.md).  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - Every function returns RB2_OK (0) or a negative RB2_ERR_* code and never
 *     throws or calls exit(); rb2_last_error_string() describes the last error.
 *   - (3,n) arrays use the Fortran host layout: column-major, i.e. xyzxyz...
 *   - Particle indices are 0-based on this side (Fortran slot - 1).
 *   - The caller owns all host arrays; the library owns all device memory and
 *     the authoritative particle state between rb2_upload_particles and
 *     rb2_download_particles.
 *   - Calls are made from ONE host thread (the Fortran master thread, outside
 *     any OpenMP region); the library is not re-entrant.
 *   - There is NO CPU fallback: every compute entry point fails with
 *     RB2_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef RUMDEED_B200_H
#define RUMDEED_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB2_OK              0
#define RB2_ERR_NOT_INIT   -1
#define RB2_ERR_CUDA       -2
#define RB2_ERR_ARG        -3
#define RB2_ERR_CAPACITY   -4
#define RB2_ERR_GEOMETRY   -5

/* src/mod_verlet.F90:62-65 (ACC_GEOM_*) */
#define RB2_GEOM_PLANAR 1
#define RB2_GEOM_TIP    2
/* src/mod_global.F90:101-118 */
#define RB2_SPECIES_ELEC 1
#define RB2_SPECIES_ION  2
#define RB2_SPECIES_ATOM 3
#define RB2_REMOVE_TOP   1
#define RB2_REMOVE_BOT   2
#define RB2_REMOVE_RECOM 3
#define RB2_REMOVE_ION   4

#define RB2_PLANES_MAX    10    /* planes_N_max, src/mod_global.F90:337 */
#define RB2_MAX_LIFE_TIME 1000  /* src/mod_global.F90:280 */

/* Scalars the hot path reads from mod_global / mod_hyperboloid_tip after Init_*.
 * (E_z, d, N_ic_max, image_charge: src/mod_verlet.F90:1238-1248.) */
typedef struct rb2_config {
    int    geometry;        /* RB2_GEOM_* == ACC_Geometry(), src/mod_verlet.F90:1168 */
    int    image_charge;    /* logical image_charge */
    int    N_ic_max;
    int    planes_N;
    double V_s;
    double d;               /* gap spacing */
    double E_z;             /* -V_d/d, src/mod_verlet.F90:2052 */
    double box_dim[3];
    double time_step;
    double planes_z[RB2_PLANES_MAX];
    /* hyperboloid tip, src/mod_hyperboloid_tip.f90:11-21 */
    double a_foci, eta_1, shift_z, pre_fac_E_tip, pre_fac_E_tip_unit_voltage, h_tip, r_tip, max_xi;
    int    capacity;        /* MAX_PARTICLES, src/mod_global.F90:98 */
    int    device;          /* CUDA ordinal, -1 = keep the current device */
} rb2_config;

/* Counters of mod_global (nrPart, nrElec, ... and the nr*_remove* family). */
typedef struct rb2_counts {
    int nrPart, nrElec, nrIon, nrAtom, nrID, nrPart_dropped;
    int nrPart_remove, nrElec_remove, nrIon_remove, nrAtom_remove;
    int nrPart_remove_top, nrPart_remove_bot;
    int nrElec_remove_top, nrElec_remove_bot;
    int nrIon_remove_top, nrIon_remove_bot;
    int nrPart_remove_ion, nrElec_remove_ion, nrAtom_remove_ion;  /* reason RB2_REMOVE_ION, src/mod_pair.F90:273-279, :327-333 */
} rb2_counts;

/* One absorbed-electron or plane-crossing record, in the order the serial
 * reference writes them (ascending particle index; absorb before planes).
 * kind 1: density_absorb_top.bin, 2: density_absorb_bot.bin (src/mod_pair.F90:243-257),
 * kind 3: planes-<plane+1>.bin (src/mod_verlet.F90:360). x,y in length_scale units. */
typedef struct rb2_event {
    int    kind, plane, index;
    double x, y, vx, vy, vz;
    int    emit, sec, id;
} rb2_event;

/* What Update_Position(step) leaves in mod_global for the writers. */
typedef struct rb2_step_result {
    double ramo_current[4];      /* per species, index = species id (src/mod_verlet.F90:488) */
    double avg_part_vel[3];      /* already divided by the counts (src/mod_verlet.F90:428-447) */
    double avg_elec_vel[3];
    double avg_ion_vel[3];
    int    n_events;             /* records waiting in rb2_get_events */
    rb2_counts counts;
    float  accel_ms;             /* device time of the pair kernel(s), CUDA events */
    float  step_ms;              /* device time of the whole step */
} rb2_step_result;

/* ---- life cycle: end of Init_* / Clean_up (src/main.F90:146, :266) ------------ */
int rb2_init(const rb2_config *cfg);
int rb2_finalize(void);
/* Re-read the scalar parameters (Set_Voltage changes E_z; tests switch image_charge). */
int rb2_update_config(const rb2_config *cfg);
const char *rb2_last_error_string(void);
/* 1 when a usable sm_100 device is present (no state needed). */
int rb2_device_available(void);

/* ---- particle state: replaces the `!$acc update device` sites
 *      (src/mod_verlet.F90:92-94, :1256-1263).  NULL int arrays default to
 *      species=elec, step=0, emitter=1, section=1, life=-1, id=slot. ---------- */
int rb2_upload_particles(int n, const double *pos, const double *prev_pos, const double *vel,
                         const double *acc, const double *acc_prev, const double *acc_prev2,
                         const double *charge, const double *mass,
                         const int *species, const int *step, const int *emitter,
                         const int *section, const int *life, const int *id, int nrID);
/* Any output pointer may be NULL.  Arrays must hold nrPart entries. */
int rb2_download_particles(double *pos, double *prev_pos, double *vel,
                           double *acc, double *acc_prev, double *acc_prev2,
                           double *charge, double *mass,
                           int *species, int *step, int *emitter, int *section, int *life, int *id,
                           int *mask);
int rb2_get_counts(rb2_counts *out);

/* Add_Particle (src/mod_pair.F90:29-159) for k particles, appended in call order;
 * ids are assigned from the library's nrID.  Particles beyond capacity are counted
 * in nrPart_dropped, like the reference. */
int rb2_add_particles(int k, const double *pos, const double *vel, const int *species,
                      int step, const int *emit, const int *sec, const int *life);
/* MAX_PARTICLES - nrPart: how many of the next rb2_add_particles calls' particles the store still accepts (the rest
 * is dropped and counted, src/mod_pair.F90:36-43).  Host-side state only: no device synchronisation. */
int rb2_capacity_left(int *out);
/* Mark_Particles_Remove (src/mod_pair.F90:169-339) for k host-chosen particles.  reason: RB2_REMOVE_TOP / _BOT for any
 * species, RB2_REMOVE_RECOM for electrons and ions, RB2_REMOVE_ION for electrons and atoms.  A reason outside 1..4 is the reference's
 * 'Error unknown remove case' and fails with RB2_ERR_ARG before anything is marked; a defined reason on a species the
 * reference has no case for (e.g. an ion with RB2_REMOVE_ION) marks the particle without a per-reason count, which is
 * what the reference does after printing its message. */
int rb2_mark_remove(int k, const int *index, const int *reason);
/* Remove_Particles (src/mod_pair.F90:352-562): stable compaction, counters reset. */
int rb2_remove_marked(int step, rb2_counts *out);
/* life_time(1:MAX_LIFE_TIME, 1:nrSpecies) as [lt][species] with lt, species 0-based +1
 * padding, i.e. out[(lt)*4 + species]; (RB2_MAX_LIFE_TIME+1)*4 entries. */
int rb2_get_life_time(long long *out);

/* ---- dynamics ------------------------------------------------------------------ */
/* Update_Position(step) (src/main.F90:190 -> src/mod_verlet.F90:123-162):
 * Beeman position update + boundary + planes, acceleration, velocity + Ramo. */
int rb2_step(int step, rb2_step_result *out);
/* The three phases separately (the reference's unit tests call them one by one). */
int rb2_update_position(int step);                 /* src/mod_verlet.F90:197-232 */
int rb2_accel_only(void);                          /* Calculate_Acceleration_Particles, :597 (overwrites) */
int rb2_update_velocity(rb2_step_result *out);     /* src/mod_verlet.F90:449-509 */
int rb2_get_events(int max_events, rb2_event *out, int *n_out);
/* ramo_current_emit(1:n_sec, 1:n_emit) of the last velocity update (src/mod_verlet.F90:489-492, written by
 * Write_Ramo_Current when write_ramo_sec is set, src/mod_pair.F90:822-826): the Ramo current of the particles that
 * carry section sec of emitter emit, Fortran layout out[(emit-1)*n_sec + sec-1].  Accumulated by rb2_step /
 * rb2_update_velocity once rb2_set_option("ramo_sections", S) (and "ramo_emitters", default 1) has switched it on;
 * a keyed sum in a fixed order (bit-identical from run to run).  Entries beyond the configured S are 0. */
int rb2_get_ramo_sections(int n_sec, int n_emit, double *out);
/* Stateless form of the reference's OpenACC call (upload positions, run the
 * kernel, copy the accelerations out: src/mod_verlet.F90:1254-1340) with HOST
 * buffers; used for the end-to-end measurement. */
int rb2_accel_host(int n, const double *pos, const double *charge, const double *mass, double *acc_out);

/* ---- field evaluation ------------------------------------------------------------- */
/* Calc_Field_at_Batch (src/mod_verlet.F90:1635); M = 1 is Calc_Field_at (:1466).
 * Synchronous on return. */
int rb2_field_batch(int M, const double *pos_in, double *field_out);
/* Same, plus the contribution of n_new particles that are not in the store yet
 * (exact by linearity): keeps the serial semantics of the default samplers
 * (src/mod_field_emission_v2.F90:1122, src/mod_photo_emission.f90:603-686). */
int rb2_field_batch_delta(int M, const double *pos_in, int n_new, const double *new_pos,
                          const double *new_charge, double *field_out);
/* Particles_To_Device / Release_Device_Particles (src/mod_verlet.F90:85-111): the
 * state is already device resident, kept for source compatibility. */
int rb2_field_window_open(void);
int rb2_field_window_close(void);

/* ---- multi-GPU plumbing --------------------------------------------------
 * The i-range [i_begin, i_end) of the acceleration evaluation this process owns
 * (global, 0-based).  Default: everything. */
int rb2_set_partition(int i_begin, int i_end);
/* Pair-symmetric evaluation split over processes: (target superblock, source group) work units are
 * dealt round-robin to `world` processes; each computes partial raw sums (rb2_accel_partial), the host
 * plumbing all-reduces the buffer "raw" (3 x padded-N doubles), then every process finalises. */
int rb2_set_pair_rank(int rank, int world);
int rb2_accel_partial(void);
int rb2_accel_finalize(void);
/* The exchange over NVLink peer memory instead of an all-reduce by the host plumbing: every process exports its
 * exchange block (partial sums for up to n_max particles + flags) as a CUDA IPC handle of RB2_P2P_HANDLE_BYTES
 * bytes, the host plumbing gathers the `world` handles (rank order, any transport) and every process attaches
 * them.  From then on rb2_accel_finalize -- and rb2_step / rb2_accel_only, which then accept a split -- waits for
 * the peers' partial sums, adds them in rank order straight from the peers' memory and finalises, in one kernel;
 * no host involvement, no other collective.  rb2_p2p_attach also sets the pair rank. */
#define RB2_P2P_HANDLE_BYTES 64
int rb2_p2p_export(int n_max, void *handle_out);
int rb2_p2p_attach(int world, int rank, const void *handles);
int rb2_p2p_detach(void);
/* ONE process driving several GPUs (the Fortran host is a single process, src/main.F90:5-29): call once after rb2_init,
 * before any particle exists.  devices[0] must be the device of rb2_init.  A replica of the particle store is kept on
 * every listed device (all state-changing calls go to all of them), the pair work of rb2_step / rb2_accel_only /
 * rb2_accel_host is split over them and the partial sums are exchanged through NVLink peer memory inside the finalise
 * kernel -- the scheme of rb2_p2p_attach without the handle exchange.  Everything that only reads the state (field
 * batches, samplers, downloads, counters) is served by the first device.  Planar geometry; not combinable with
 * rb2_p2p_attach or the collision step. */
int rb2_set_devices(int n_devices, const int *devices);
/* Tunables: "pair_mode" 0 auto / 1 gather / 2 pair-symmetric, "sym_min_n", "sym_budget_mb", "sym_waves", "sym_kmax", "sym_gmax",
 * "sym_far" (1 / 0: with d >= 1 um the acceleration kernels evaluate the three image partners that are at least d away
 * without the softening term, a change of <= 3e-12 of those terms; 1 by default; it switches itself off while a particle
 * handed in by the caller lies outside 0 <= z <= d),
 * "sym_tpl" (targets per lane of the pair-symmetric kernel: 0 auto, 1, 2), "step_graph" (1 / 0: replay rb2_step as a CUDA graph while
 * consecutive steps queue identical work), "ramo_sections" / "ramo_emitters" (size
 * of the per-section Ramo table, 0 sections = off), "event_buffer" (initial number of
 * absorb / plane-crossing records the device buffer holds; it grows on demand); emission samplers: "mh_small"
 * (1 / 0: single-barrier kernel for few chains), "mh_small_max" (its chain limit, <= 512), "mh_ctas_per_sm" (1..4,
 * many-chain kernel), "tip_field_small" (1 / 0: CTA-per-point tip field kernel for small batches). */
int rb2_set_option(const char *name, double value);
/* Host-only view of the planning of the pair-symmetric work (no device, no rb2_init needed): for n particles dealt to
 * `world` ranks, shape_out[6] = {targets per lane T, K superblocks x G source tiles per work unit, source tiles per band,
 * number of source tiles, number of target superblocks}, rank_cost_out[world] = tile pairs dealt to every rank,
 * units_out[3 * cap] = (band, set, group) of THIS rank's units in launch order (n_units_out of them),
 * table_hash_out = hash of the owner table every rank must agree on.  tpl 0 = auto; sm_count, waves, kmax, gmax,
 * budget_mb as the options of the same names (148, 64, 12, 24, 2048 by default).  Used by the multi-process CPU tests. */
int rb2_sym_plan_probe(int n, int tpl, int world, int rank, int sm_count, double waves, int kmax, int gmax, double budget_mb,
                       int *shape_out, long long *rank_cost_out, int *units_out, int cap, int *n_units_out,
                       unsigned long long *table_hash_out);
/* Device pointer + byte size of the (3,capacity) acceleration buffer so that the
 * host plumbing (torch.distributed / NCCL) can all-gather the slices in place. */
int rb2_device_buffer(const char *name, void **dev_ptr, size_t *bytes);
/* Block the host until the library's stream is idle. */
int rb2_synchronize(void);
/* CUDA stream handle (cudaStream_t) the library launches on. */
int rb2_stream(void **stream_out);

/* E_z only, at M points ON the cathode plane (pos_in[3k+2] must be 0), planar geometry.  Same value as
 * the z component of rb2_field_batch (Calc_Field_at_Batch, src/mod_verlet.F90:1635) at such points, but
 * computed from the mirror-antisymmetric form of the image series (E_x = E_y = 0 there when image charges
 * are on): less than half the arithmetic.  What the planar emission integrands and samplers need
 * (field(3) in src/mod_field_emission_v2.F90:668-745, :1122-1458). */
int rb2_field_surface_z(int M, const double *pos_in, double *Ez_out);

/* Sample_Elec_Position (src/mod_pair.F90:975-1037), the sweep only: for each of the nrPart particles the distance
 * to the nearest OTHER electron and its 0-based index (rows that are not electrons: 1000.0 and -1, the reference's
 * initial value).  Bit-exact with the serial scan (lowest index wins ties).  Either pointer may be NULL. */
int rb2_nearest_electron(double *dist_out, int *id_out);

/* ---- device-resident emission sampler ------------------------------------------------- */
/* Lock-step Metropolis-Hastings over the planar emitter: replaces the host loop of
 * Metropolis_Hastings_rectangle_J_batch (src/mod_field_emission_v2.F90:1284-1458; kind 1) and, with
 * kind 2, runs the chains of src/mod_field_thermo_emission.F90:198-364 in the same lock-step form.
 * All jump iterations (proposal, M x N surface field, accept / reject, shared step adaptation) are
 * enqueued on the device; the host is blocked once, for the result. */
typedef struct rb2_mh_config {
    int    kind;          /* 1: log electron supply (Elec_Supply_log :589); 2: ln J_GTF (thermal-field) */
    int    ndim;          /* jump iterations (25*8 for kind 1, 25 for kind 2) */
    int    ndim_first;    /* warm-up iterations using init_std and leaving MH_std alone */
    int    image_charge;  /* Schottky-Nordheim barrier functions t_y / v_y on (1) or 1.0 (0) */
    int    y_num, x_num;  /* work-function checkerboard, src/mod_work_function.F90:389-487 */
    double emit_pos[2], emit_dim[2];
    double T_temp;        /* kind 2 */
    double init_std;      /* warm-up step as a fraction of the emitter side (0.10) */
    double target_rate, std_gain, std_min, std_max; /* MH_std_update :603-612: 0.35, 0.025, 0.00005 / 0.005, 0.125 */
} rb2_mh_config;
/* M chains; w_theta is the [y_num][x_num] work-function table (host).  Outputs (host): log escape
 * probability df_out[M] (kind 1; -HUGE for a chain that found no favourable spot), surface field
 * F_out[M] (>= 0 marks a failed chain) and positions pos_out[3M].  a_rate_io / mh_std_io carry the
 * sampler's adaptive state across calls (a_rate / MH_std of the reference module). */
int rb2_mh_planar(const rb2_mh_config *cfg, const double *w_theta, int M, unsigned long long seed,
                  double *df_out, double *F_out, double *pos_out, double *a_rate_io, double *mh_std_io);

/* The reference's DEFAULT sampler, mh_batch = .false.: the chains of a time step one after the other, chain s on the field
 * of the store plus the electrons emitted by chains 0 .. s-1 in this call, the shared step adapted once per chain
 * (Metropolis_Hastings_rectangle_J inside the insert loop of Do_Field_Emission_Planar_rectangle,
 * src/mod_field_emission_v2.F90:1122-1265, :322-380; kind 2: src/mod_field_thermo_emission.F90:198-364).  Strictly
 * sequential by construction; the whole loop runs in ONE kernel (one CTA, ~2.5 us per field evaluation instead of one
 * ~40 us host round trip each).  emit_out[k] = 1: candidate k passed the emission test ln u <= D_f (kind 2: found a
 * start) and was counted in the field of the later chains -- the caller inserts exactly those, in order, at z = 1 nm
 * (rb2_add_particles).  At most 1024 emitted electrons per call. */
int rb2_mh_planar_serial(const rb2_mh_config *cfg, const double *w_theta, int M, unsigned long long seed,
                         double *df_out, double *F_out, double *pos_out, int *emit_out, double *a_rate_io, double *mh_std_io);

/* Lock-step chains on the hyperboloid tip: Metro_algo_tip_v3 (src/mod_emission_tip.f90:1241-1390) for the M candidates
 * of a time step together -- (xi, phi) proposals with reflection in xi and wrap in phi, target ln S + 1/2 ln(xi^2 - eta_1^2)
 * on the NORMAL field component, shared adaptive step (one update per jump after the warm-up quarter).  Every jump is
 * queued on the device (proposals, the tip field kernel of rb2_field_batch, accept / reject); the host is blocked only
 * for the start-spot rounds and once for the result.  Outputs (host): normal surface field eta_f_out[M] (1.0 marks a chain
 * that found no favourable spot), escape probability df_out[M] (Escape_Prob_Tip :1734-1760; 0 for failed chains) and
 * positions pos_out[3M] on the tip surface.  a_rate_io / mh_std_io: a_rate / MH_std of the reference module (:50). */
int rb2_mh_tip(int M, int ndim, unsigned long long seed, double *eta_f_out, double *df_out, double *pos_out,
               double *a_rate_io, double *mh_std_io);

/* One level of the planar supply quadrature on the device (Do_Surface_Integration_FE / _Simple,
 * src/mod_field_emission_v2.F90:635-745, src/mod_field_thermo_emission.F90:394-466; the host's stand-in for Cuba is a
 * rank-1 lattice with K <= 8 random shifts, refined level by level): the nodes k = n_done+1 .. n_done+n_new of every
 * shifted lattice -- u = frac(k a1 + shift(1,r)), v = frac(k a2 + shift(2,r)) over the emitter rectangle of cfg -- are
 * generated on the device, the cathode-plane field kernel of rb2_field_surface_z runs on them, and the integrand
 * (kind 1: Elec_Supply_V2 :579-587; kind 2: J_GTF dt / q_0) at the node's work function is summed per shift in a fixed
 * tree.  sums_out[r] = sum of the integrand over the new nodes of shift r (the caller multiplies by the emitter area),
 * *ez_sum_out = sum of E_z over all new nodes.  Needs image_charge = 1 or 0 alike (E_z from the surface kernel). */
int rb2_planar_supply_level(const rb2_mh_config *cfg, const double *w_theta, int kind, int K, const double *shifts, int n_done,
                            int n_new, double *sums_out, double *ez_sum_out);

/* The tip's supply sum on the device: Do_Field_Emission_Tip_OLDCODE (src/mod_emission_tip.f90:431-481) adds
 * Elec_Supply(A_k, F_k) (:1710-1718) over a 100 x 100 (xi, phi) midpoint grid of the tip surface in every time step.
 * rb2_tip_supply_set_grid: the M nodes pts(3,M), unit surface normals normals(3,M) (surface_normal,
 * src/mod_hyperboloid_tip.f90:25-34) and patch areas area(M) (Tip_Area) -- geometry only, handed over once.
 * rb2_tip_supply: field of the resident particles at the nodes (the kernel of rb2_field_batch), F_k = normal . field,
 * n_s = sum over F_k < 0 of A_k a_FN F_k^2 dt / (q_0 w_theta t_y(l)^2) with w_theta = 4.7 eV (:45), and the plain sum of
 * the F_k (the caller divides by M for F_avg).  256 nodes per CTA are added in a fixed tree, the CTA sums in CTA order
 * on the host: deterministic, but not the reference's serial order (agreement ~1e-15 relative). */
int rb2_tip_supply_set_grid(int M, const double *pts, const double *normals, const double *area);
int rb2_tip_supply(double *n_s_out, double *F_sum_out);

/* ---- electron / N2 collisions (collision_mode 1 and 2) -------------------------
 * Do_Electron_Atom_Collisions (src/mod_collisions.F90:30-76, called from src/main.F90:202 through
 * Do_Collisions, src/mod_verlet.F90:164-170), one-time-step variants:
 *   continuous ionisation  Do_Continuous_Ionization_ots   src/mod_collisions.F90:558-705
 *   discrete recombination Do_Discrete_Recombination_ots  src/mod_collisions.F90:86-245
 *                          (O(nrIon x nrElec) sweep with the quartic of src/mod_polynomialroots.F90)
 * The per-electron collision data of Update_Collision_Data (:2103-2134) is recomputed from the velocities inside
 * the kernels; rb2_collision_data returns it for inspection.  Discrete ionisation (modes 3, 4: N2 atoms as
 * particles) is not on the device path: rb2_collisions_init refuses those modes. */
typedef struct rb2_collision_config {
    int    collision_mode;    /* 1 continuous ionisation, 2 = 1 + discrete recombination */
    int    ion_life_time;     /* time steps, src/mod_global.F90:210 */
    double n_d;               /* N2 number density P/(k_b T), src/main.F90:382-383 */
    double cyl_radius;        /* emitters_dim(1,1): collisions only inside this radius (:594) */
    int    n_tot, n_ion;      /* rows of N2-tot-cross.txt / N2-ion-cross.txt (Read_Cross_Section, :1909-1983) */
    const double *tot_energy, *tot_data, *ion_energy, *ion_data;  /* eV, 1e-20 m^2 */
} rb2_collision_config;

/* One recombination: the arguments of Write_Recombination_Data (src/mod_pair.F90:919-926) plus what the two
 * Mark_Particles_Remove(.., remove_recom) calls write to density_absorb_recom.bin (:265-272, :308-315).
 * Slots are 0-based; ion_life = step - particles_step(ion). */
typedef struct rb2_recomb_record {
    int    step, elec_slot, ion_slot, elec_emit, ion_life;
    int    elec_sec, elec_id, ion_emit, ion_sec, ion_id;
    double ion_pos[3], elec_pos[3];
    double elec_speed, dist, recom_rad, t;
} rb2_recomb_record;

/* One ionisation: the arguments of Write_Ionization_Data (src/mod_pair.F90:930-935) plus the state of the three
 * particles involved.  in_slot is 0-based; new_id / ion_id are the particle ids given to the ejected electron and
 * the ion (-1 when the store was full and the particle was dropped). */
typedef struct rb2_ionization_record {
    int    step, in_slot, new_id, ion_id, elec_emit, pad;
    double pos[3];
    double in_speed, out_speed, new_speed;
    double new_vel[3], ejec_pos[3], ejec_vel[3], ion_pos[3];
    double E1, collE, ejecE;
} rb2_ionization_record;

typedef struct rb2_collision_result {
    int nrCollisions, nrIonizations, nrRecombinations;   /* the three columns of collisions.dt (:74-75) */
    int nrIonsExpired;                                    /* ions removed for their age (:121-124) */
    int nrPart_remove_recom, nrElec_remove_recom, nrIon_remove_recom; /* since the last Remove_Particles */
    int n_candidates;         /* (ion, electron) pairs that reached the quartic test (diagnostic) */
    rb2_counts counts;
    float ms;                 /* device time of the call, CUDA events */
} rb2_collision_result;

int rb2_collisions_init(const rb2_collision_config *cfg);
/* Update_Collision_Data_All_ots (:2155-2169).  out (may be NULL) receives 5 doubles per particle slot:
 * cur_energy, ion_cross_sec, ion_cross_rad, recom_cross_rad, tot_cross_sec (zeros for non-electrons). */
int rb2_collision_data(double *out);
int rb2_continuous_ionization(int step, unsigned long long seed, rb2_collision_result *out);
int rb2_discrete_recombination(int step, rb2_collision_result *out);
/* Do_Electron_Atom_Collisions(step): ionisation, then (mode 2) recombination; nrCollisions includes the
 * recombinations like the reference's collisions.dt line. */
int rb2_do_collisions(int step, unsigned long long seed, rb2_collision_result *out);
/* Records of the last call, in the order the serial reference writes them (ascending ion / electron slot). */
int rb2_get_recombination_records(int max_records, rb2_recomb_record *out, int *n_out);
int rb2_get_ionization_records(int max_records, rb2_ionization_record *out, int *n_out);

/* Test hook: the device transcription of QuarticRoots (src/mod_polynomialroots.F90:336-510, used by the recombination
 * test) on n polynomials; coeffs = {quartic, cubic, quadratic, linear, constant} each (quartic != 0, else code 0 and NaN
 * roots).  codes_out[n]: the routine's return code (31, 42, 44, 23; 0 when the constant term is zero, where the reference
 * leaves its code unset); roots_out[8n]: re, im of z(1..4). */
int rb2_probe_quartic_roots(int n, const double *coeffs, int *codes_out, double *roots_out);

/* ---- measurement helpers --------------------------------------------------------------- */
/* Independent-DFMA-chain micro-benchmark: measured FP64 peak of this GPU in TFLOP/s
 * (FMA = 2 flops) over about `ms_target` milliseconds. */
int rb2_fp64_peak(double ms_target, double *tflops_out, float *ms_out);
/* Launch statistics since init / last reset: kernels launched by this library. */
int rb2_launch_count(long long *out, int reset);
/* Named counters: "graph_replays" (rb2_step calls served by replaying the captured CUDA graph), "graph_launches"
 * (kernels + copies inside that graph), "launches" (same as rb2_launch_count). */
int rb2_get_stat(const char *name, double *out);
/* Device time of the last acceleration evaluation (ms), and its launch geometry. */
int rb2_last_accel_info(float *ms, int *grid_x, int *grid_y, int *block, int *j_split);

#ifdef __cplusplus
}
#endif
#endif /* RUMDEED_B200_H */

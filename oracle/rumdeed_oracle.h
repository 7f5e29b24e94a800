/*
 * rumdeed_oracle.h -- CPU restatement of RUMDEED's per-timestep hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check
 * in __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product path (rumdeed_b200/, librumdeed_b200.so) never links,
 * imports or calls anything in this directory.
 *
 * Every function cites the reference Fortran it restates (paths relative to the
 * RUMDEED source tree).  The reference itself cannot be built in this image (no
 * Fortran compiler), so the restatement is pinned against the golden vectors in
 * the reference's own test module src/mod_tests.F90 (see tests/test_oracle_*.py).
 *
 * Array layout follows the Fortran host: (3,N) column-major, i.e. xyzxyz...
 */
#ifndef RUMDEED_ORACLE_H
#define RUMDEED_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* species / removal flags: src/mod_global.F90:101-118 */
enum { ORC_SPECIES_UNKNOWN = 0, ORC_SPECIES_ELEC = 1, ORC_SPECIES_ION = 2, ORC_SPECIES_ATOM = 3 };
enum { ORC_REMOVE_UNKNOWN = 0, ORC_REMOVE_TOP = 1, ORC_REMOVE_BOT = 2, ORC_REMOVE_RECOM = 3, ORC_REMOVE_ION = 4 };
enum { ORC_GEOM_OTHER = 0, ORC_GEOM_PLANAR = 1, ORC_GEOM_TIP = 2 };  /* src/mod_verlet.F90:62-65 */

#define ORC_PLANES_MAX 10        /* src/mod_global.F90:337 */
#define ORC_MAX_LIFE_TIME 1000   /* src/mod_global.F90:280 */
#define ORC_MAX_SECTIONS  (96 * 96) /* src/mod_global.F90:100 */
#define ORC_MAX_EMITTERS  1       /* src/mod_global.F90:99 */

/* Physical constants, src/mod_global.F90:26-75,333 */
typedef struct {
    double pi, h, k_b, c, mu_0, epsilon_0, m_u, h_bar, m_0, q_0;
    double m_N2, m_N2p, length_scale, time_scale, div_fac_c;
    double a_FN, b_FN, l_const; /* src/mod_field_emission_v2.F90:36-47 */
} orc_constants;

/* Run parameters the hot path reads (namelist + Init_* derived values). */
typedef struct {
    int    geometry;      /* ORC_GEOM_* */
    int    image_charge;  /* logical */
    int    N_ic_max;
    int    planes_N;
    double V_s, d, E_z;
    double box_dim[3];
    double time_step;
    double planes_z[ORC_PLANES_MAX];
    /* hyperboloid tip, src/mod_hyperboloid_tip.f90:11-21, src/mod_emission_tip.f90:105-125 */
    double d_tip, R_base, h_tip;
    double a_foci, eta_1, theta_tip, r_tip, max_xi, shift_z;
    double pre_fac_E_tip, pre_fac_E_tip_unit_voltage;
} orc_params;

void orc_get_constants(orc_constants *c);
void orc_params_planar(orc_params *p, double V_s, double d, const double box_dim[3], double time_step,
                       int image_charge, int N_ic_max);
void orc_params_tip(orc_params *p, double V_s, double d_tip, double R_base, double h_tip,
                    const double box_dim[3], double time_step, int image_charge);

/* --- geometry math ------------------------------------------------------- */
void orc_force_image_charges_v2(const orc_params *p, const double pos_1[3], const double pos_2[3], double out[3]);
void orc_sphere_ic_field(const orc_params *p, const double pos_1[3], const double pos_2[3], double out[3]);
void orc_field_E_planar(const orc_params *p, const double pos[3], double out[3]);
void orc_field_E_hyperboloid(const orc_params *p, const double pos[3], double out[3]);
void orc_field_E(const orc_params *p, const double pos[3], double out[3]);
void orc_image_charge_effect(const orc_params *p, const double pos_1[3], const double pos_2[3], double out[3]);
void orc_E_zunit(const orc_params *p, const double pos[3], double out[3]);
double orc_xi_coor(const orc_params *p, double x, double y, double z);
double orc_eta_coor(const orc_params *p, double x, double y, double z);
double orc_phi_coor(double x, double y);
void orc_xyz_corr(const orc_params *p, double xi, double eta, double phi, double out[3]);
void orc_surface_normal(const orc_params *p, const double pos[3], double out[3]);
double orc_field_normal(const orc_params *p, const double pos[3], const double field[3]);
double orc_tip_area(const orc_params *p, double xi_1, double xi_2, double phi_1, double phi_2);

/* --- acceleration (all ADD into acc unless stated) ------------------------ */
/* src/mod_verlet.F90:625-751  generic pair loop through the geometry functions */
void orc_accel_generic(const orc_params *p, int n, const double *pos, const double *q, const double *m,
                       const int *species, double *acc);
/* src/mod_verlet.F90:763-884  planar specialised pair loop (i<j scatter, OpenMP) */
void orc_accel_planar(const orc_params *p, int n, const double *pos, const double *q, const double *m,
                      const int *species, double *acc);
/* Same loop restricted to rows i0<=i<i1 (0-based) -- bounded sample for the CPU baseline.
 * Returns the number of unordered pairs evaluated. */
long long orc_accel_planar_rows(const orc_params *p, int n, const double *pos, const double *q, const double *m,
                                const int *species, double *acc, int i0, int i1, int i_stride);
/* src/mod_verlet.F90:1217-1429  gather formulation (OpenACC), OVERWRITES acc */
void orc_accel_gather(const orc_params *p, int n, const double *pos, const double *q, const double *m,
                      double *acc);
/* Same formulas in long double with compensated sums: the "truth" both orders are compared to.
 * Rows i0<=i<i1 only (0-based); acc_out has 3*(i1-i0) entries. */
void orc_accel_gather_ld(const orc_params *p, int n, const double *pos, const double *q, const double *m,
                         int i0, int i1, double *acc_out);
/* same for a list of rows (OpenMP over the list) */
void orc_accel_gather_ld_rows(const orc_params *p, int n, const double *pos, const double *q, const double *m,
                              int nrows, const int *rows, double *acc_out);

/* --- field ------------------------------------------------------------------ */
/* src/mod_verlet.F90:1466-1529 */
void orc_calc_field_at(const orc_params *p, int n, const double *pos, const double *q, const int *species,
                       const double pt[3], double out[3]);
/* src/mod_verlet.F90:1635-1911 (batch == point by point) */
void orc_calc_field_at_batch(const orc_params *p, int n, const double *pos, const double *q, const int *species,
                             int M, const double *pts, double *out);
void orc_calc_field_at_ld(const orc_params *p, int n, const double *pos, const double *q, const int *species,
                          const double pt[3], double out[3]);

/* --- particle store (src/mod_global.F90:128-172, src/mod_pair.F90) --------- */
typedef struct {
    int    kind;   /* 1 absorb top, 2 absorb bot, 3 plane crossing */
    int    plane;  /* plane index (0-based) for kind 3 */
    int    index;  /* particle slot (0-based) at the time of the event */
    double x, y;   /* in units of length_scale, as written to the .bin files */
    double vx, vy, vz;
    int    emit, sec, id;
} orc_event;

typedef struct {
    int capacity;
    double *pos, *prev_pos, *vel, *acc, *acc_prev, *acc_prev2; /* (3,cap) */
    double *charge, *mass;
    int *species, *step, *emitter, *section, *life, *id, *mask;
    int nrPart, nrElec, nrIon, nrAtom, nrID, nrPart_dropped;
    int nrPart_remove, nrElec_remove, nrIon_remove, nrAtom_remove;
    int nrPart_remove_top, nrPart_remove_bot, nrElec_remove_top, nrElec_remove_bot;
    int nrIon_remove_top, nrIon_remove_bot;
    int charge_rev;
    long long life_time[ORC_MAX_LIFE_TIME + 1][4]; /* [lt][species] */
    double ramo_current[4];                          /* per species (1-based like Fortran) */
    double avg_part_vel[3], avg_elec_vel[3], avg_ion_vel[3];
    orc_event *events; int n_events, cap_events;
    /* ramo_current_emit(1:MAX_SECTIONS, 1:MAX_EMITTERS), column-major like Fortran: [(emit-1)*MAX_SECTIONS + sec-1]
     * (src/mod_global.F90:272, zeroed per step src/mod_verlet.F90:155, accumulated :489-492) */
    double *ramo_current_emit;
} orc_store;

orc_store *orc_store_new(int capacity);
void orc_store_free(orc_store *s);
void orc_store_clear_events(orc_store *s);
/* src/mod_pair.F90:29-159 ; returns slot (0-based) or -1 when dropped */
int  orc_add_particle(orc_store *s, const orc_params *p, const double pos[3], const double vel[3],
                      int species, int step, int emit, int life, int sec);
/* src/mod_pair.F90:169-339 */
void orc_mark_particle_remove(orc_store *s, int i, int reason);
/* src/mod_pair.F90:352-562 */
void orc_remove_particles(orc_store *s, int step);
/* src/mod_verlet.F90:197-232 + :325-367 + src/mod_emission_tip.f90:1627-1647 */
void orc_update_position(orc_store *s, const orc_params *p);
/* src/mod_verlet.F90:597-620 on the store (dispatch: planar specialised / generic) */
void orc_update_acceleration(orc_store *s, const orc_params *p);
/* src/mod_verlet.F90:449-509 + :428-447 */
void orc_update_velocity(orc_store *s, const orc_params *p);
/* src/mod_verlet.F90:123-162 : position, acceleration, velocity */
void orc_step(orc_store *s, const orc_params *p);

/* --- Fowler-Nordheim helpers, src/mod_field_emission_v2.F90:515-625 -------- */
double orc_fn_v_y(const orc_params *p, double F, double w_theta);
double orc_fn_t_y(const orc_params *p, double F, double w_theta);
double orc_fn_escape_prob_log(const orc_params *p, double F, double w_theta);
double orc_fn_elec_supply_log(const orc_params *p, double F, double w_theta);
double orc_fn_elec_supply_v2(const orc_params *p, double F, double w_theta);
/* src/mod_emission_tip.f90:1657-1764 (work function 4.7 eV hard coded there) */
double orc_tip_v_y(const orc_params *p, double F, double w_theta);
double orc_tip_t_y(const orc_params *p, double F, double w_theta);
double orc_tip_escape_prob(const orc_params *p, double F, double w_theta);
double orc_tip_elec_supply(const orc_params *p, double A, double F, double w_theta);

/* Sample_Elec_Position, src/mod_pair.F90:975-1037: for every electron the distance to the nearest OTHER electron
 * (species test only: electrons already marked for removal still count) and that electron's 0-based index; rows
 * that are not electrons keep the initial distance 1000.0 and get index -1.  Strict `<` over ascending j: the
 * lowest index wins a tie.  dist = sqrt(dx*dx + dy*dy + dz*dz), summed left to right without contraction. */
void orc_nearest_elec(int n, const double *pos, const int *species, double *dist_out, int *id_out);
int orc_max_threads(void);
void orc_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif

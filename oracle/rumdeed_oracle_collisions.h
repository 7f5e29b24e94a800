/*
 * rumdeed_oracle_collisions.h -- CPU restatement of RUMDEED's electron / N2 collision step
 * (continuous ionisation + discrete recombination, collision_mode 1 and 2).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (see rumdeed_oracle.h).
 *
 * Restates src/mod_collisions.F90 (one-time-step "_ots" variants) and the quartic solver of
 * src/mod_polynomialroots.F90 function by function.  Pinned against the reference's own
 * Test_Collision_Math (src/mod_tests.F90:1742-1786: normal / folded normal values, Kramers cross
 * section at 10 and 100 eV); the polynomial solver has no vector in the reference and is checked
 * against numpy.roots.  The reference's RANDOM_NUMBER stream is compiler specific, so everything that
 * draws random numbers is comparable only statistically.
 */
#ifndef RUMDEED_ORACLE_COLLISIONS_H
#define RUMDEED_ORACLE_COLLISIONS_H

#include "rumdeed_oracle_emission.h" /* orc_rng */

#ifdef __cplusplus
extern "C" {
#endif

/* N2 cross-section tables (Read_Cross_Section, src/mod_collisions.F90:1909-1983): energy [eV], data [1e-20 m^2] */
typedef struct {
    int n_tot, n_ion;
    const double *tot_energy, *tot_data, *ion_energy, *ion_data;
} orc_cross_tables;

/* src/mod_global.F90:50-68: R_inf, Ryd, N_n, N_bind, Z_eff */
typedef struct { double R_inf, Ryd, N_n, N_bind, Z_eff, Z_eff2; } orc_coll_constants;
void orc_coll_get_constants(orc_coll_constants *c);

/* src/mod_collisions.F90:1872-1907 */
double orc_normal_dist(double mu, double sigma, double x);
double orc_folded_normal_dist(double mu, double sigma, double x);
double orc_folded_normal_max(double mu, double sigma);
/* src/mod_collisions.F90:1443-1449 */
double orc_kramers_cross_section(double energy);
/* BinarySearch, src/mod_global.F90:649-693 (0-based i1, i2; returns the nearer index) */
int orc_binary_search(const double *list, int n, double value, int *i1, int *i2);
/* src/mod_collisions.F90:1985-2044 */
double orc_find_cross_tot_data(const orc_cross_tables *T, double energy);
double orc_find_cross_ion_data(const orc_cross_tables *T, double energy);
/* Update_Collision_Data, src/mod_collisions.F90:2103-2134 for one velocity; out = {cur_energy, ion_cross_sec,
 * ion_cross_rad, recom_cross_rad, tot_cross_sec} */
void orc_update_collision_data(const orc_cross_tables *T, const double vel[3], double out[5]);

/* SolvePolynomial, src/mod_polynomialroots.F90:528-563 (+ QuarticRoots :336-510, CubicRoots :178-333,
 * QuadraticRoots :125-175).  roots = {re1, im1, re2, im2, re3, im3, re4, im4}.  Like the reference, root3 is
 * only assigned when code > 23 and root4 never is (the test is `outputCode > 99`): unassigned roots are NaN here.
 * The module variable outputCode survives between calls in the reference (CubicRoots does not always set it);
 * orc_poly_reset_code() sets it to 0. */
void orc_solve_polynomial(double quartic, double cubic, double quadratic, double linear, double constant,
                          int *code, double roots[8]);
void orc_poly_reset_code(void);

/* One (ion, electron) test of Do_Discrete_Recombination_ots, src/mod_collisions.F90:128-196: returns 1 when
 * recombination happens within the time step; t_out = time of entry, dist_out = |elec(t) - ion|. */
int orc_recombination_pair(const double ion_pos[3], const double elec_pos[3], const double elec_vel[3],
                           const double elec_acc[3], double recom_rad, double time_step, double *t_out,
                           double *dist_out);

typedef struct {
    int    step;
    double ion_pos[3];
    double elec_speed, dist, recom_rad;
    int    elec_slot, ion_slot, elec_emit, ion_life; /* 0-based slots; ion_life = step - particles_step(ion) */
    double t;
} orc_recomb_event;

/* Do_Discrete_Recombination_ots (serial order), src/mod_collisions.F90:86-245.  mask is updated (0 = marked) and
 * reason_out[i] (may be NULL) receives ORC_REMOVE_TOP for ions whose life time is over and ORC_REMOVE_RECOM for
 * recombined particles.  Returns nrRecombinations; *n_expired = ions removed for their age. */
int orc_discrete_recombination_ots(int n, const double *pos, const double *vel, const double *acc,
                                   const int *species, int *mask, const int *life, const int *step_born,
                                   const int *emitter, const double *recom_rad, int step, double time_step,
                                   orc_recomb_event *events, int max_events, int *reason_out, int *n_expired);

/* Get_Injected_Vec / Get_Ejected_Vec, src/mod_collisions.F90:1452-1577 */
void orc_get_injected_vec(orc_rng *r, double T, const double par_vel[3], double out[3]);
void orc_get_ejected_vec(orc_rng *r, double W, double T, const double par_vel[3], double out[3]);

typedef struct {
    int    step, in_slot;               /* colliding electron (0-based slot) */
    double pos[3];                      /* its position */
    double in_speed, out_speed, new_speed;
    double new_vel[3];                  /* colliding electron after the collision */
    double ejec_pos[3], ejec_vel[3];    /* ejected electron */
    double ion_pos[3];                  /* created ion (at rest) */
    double E1, collE, ejecE;
    int    elec_emit;                   /* emitter of the colliding electron before it is set to ion_emitter */
} orc_ionization_event;

/* The electron loop of Do_Continuous_Ionization_ots, src/mod_collisions.F90:558-705, in serial order.  vel and
 * emitter of colliding electrons are updated in place; the particles to add (electron then ion per event, in
 * event order) are described by the events.  Returns nrIonizations; *nrCollisions as in the reference. */
int orc_continuous_ionization_ots(orc_rng *r, const orc_cross_tables *T, int n, const double *pos,
                                  const double *prev_pos, double *vel, const int *species, const int *mask,
                                  int *emitter, double n_d, double cyl_radius, int step,
                                  orc_ionization_event *events, int max_events, int *nrCollisions);

#ifdef __cplusplus
}
#endif
#endif

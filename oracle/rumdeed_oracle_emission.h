/*
 * rumdeed_oracle_emission.h -- emission samplers of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 * See rumdeed_oracle_emission.c; reference citations per function are in the .c file.
 */
#ifndef RUMDEED_ORACLE_EMISSION_H
#define RUMDEED_ORACLE_EMISSION_H

#include <stdint.h>

#include "rumdeed_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_SUPPLY_FE = 1, ORC_SUPPLY_GTF = 2 };

typedef struct { uint64_t s[4]; } orc_rng;

/* One rectangular emitter + checkerboard work function + the sampler state the reference keeps
 * in module variables (a_rate, MH_std, residual). */
typedef struct {
    const orc_params *p;
    orc_store *store;
    double emit_pos[3], emit_dim[3];
    int    y_num, x_num;
    const double *w_theta_arr; /* [y_num][x_num], rows as in the `work` file */
    double T_temp;
    double a_rate, MH_std;     /* src/mod_field_emission_v2.F90:66-67 (1.0, 0.0125) */
    double MH_std_tip;         /* src/mod_emission_tip.f90:50 (1.0, clamped on first use) */
    double residual;
} orc_emission;

void   orc_rng_seed(orc_rng *r, uint64_t seed);
double orc_rng_uniform(orc_rng *r);
void   orc_box_muller(orc_rng *r, const double mean[2], const double std[2], double out[2]);
int    orc_rand_poisson(orc_rng *r, double lambda);
void   orc_get_mb_velocity(orc_rng *r, double T_temp, double out[3]);

double orc_w_theta_xy(const orc_emission *E, const double pos[3], int *sec);
double orc_kevin_jgtf_v2(double F, double T, double w_theta);

double orc_supply_integrand(const orc_emission *E, int kind, const double xx[2], double field_out[3]);
double orc_supply_grid(const orc_emission *E, int kind, int n, double F_avg[3]);

int  orc_mh_rectangle_J(orc_emission *E, orc_rng *r, double *df_out, double *F_out, double pos_out[3]);
void orc_mh_rectangle_J_batch(orc_emission *E, orc_rng *r, int M, double *df_out, double *F_out, double *pos_out);
int  orc_do_field_emission_planar(orc_emission *E, orc_rng *r, int step, double N_sup, int mh_batch, double *df_avg_out);

int  orc_mh_rectangle_J_thermo(orc_emission *E, orc_rng *r, double pos_out[3]);
int  orc_do_field_thermo_emission_planar(orc_emission *E, orc_rng *r, int step, double N_sup);

double orc_get_laser_energy(orc_rng *r, double laser_energy, double laser_variation);
int  orc_do_photo_emission_rectangle(orc_emission *E, orc_rng *r, int step, double p_eV, int photon_mode, int max_elec_emit);

double orc_tip_supply_grid(const orc_emission *E, int nr_xi, int nr_phi, double *F_avg_out);
int  orc_metro_algo_tip_v3(orc_emission *E, orc_rng *r, int ndim, double *xi_out, double *phi_out, double *eta_f_out,
                           double *df_cur, double par_pos[3]);
int  orc_do_field_emission_tip(orc_emission *E, orc_rng *r, int step, double n_s);

#ifdef __cplusplus
}
#endif
#endif

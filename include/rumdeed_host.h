/*
 * rumdeed_host.h -- C entry points of librumdeed_host.so, the C++ mirror of the RUMDEED Fortran
 * host around the hot path (input namelist, emission plugins init / do-emission / clean-up,
 * main loop, output writers; see rumdeed_b200/host/rh_host.hpp for the reference citations).
 * Everything below runs on top of the device library through include/rumdeed_b200.h.
 */
#ifndef RUMDEED_HOST_H
#define RUMDEED_HOST_H

#ifdef __cplusplus
extern "C" {
#endif

/* In-memory equivalent of the `input` namelist + `work` + `laser` files (SI units). */
typedef struct rh_setup {
    int    emission_mode;      /* 1 photo, 3 tip, 9 thermal-field, 10 field emission (src/main.F90:106-144) */
    double V_s;
    double box_dim[3];         /* m */
    double time_step;          /* s */
    int    image_charge, N_ic_max;
    double emitters_pos[3], emitters_dim[3]; /* m; for the tip: d_tip, R_base, h_tip */
    int    emitters_type, emitters_delay;
    double T_temp;
    int    mh_batch;           /* 0 serial chains (the reference's default; one device kernel per step), -1 serial chains as a host
                                  loop of single-point field calls, 1 lock-step chains (host loop), 2 lock-step chains on the GPU */
    int    planes_N;
    double planes_z[10];       /* m */
    double cuba_epsabs, cuba_epsrel;
    int    cuba_mineval, cuba_maxeval;
    int    work_y_num, work_x_num;
    const double *work_w_theta; /* [y_num][x_num] rows as in the `work` file */
    int    laser_gauss_mode, laser_mode, photon_mode;
    double laser_energy, laser_variation, gauss_center, gauss_width, gauss_amplitude;
    int    max_particles;
    unsigned long long seed;
    int    ramo_sections;      /* > 0: keep ramo_current_emit(1:ramo_sections, 1) per step (write_ramo_sec, src/mod_global.F90:352) */
} rh_setup;

typedef struct rh_state {
    int    step;
    int    nrPart, nrElec, nrIon, nrID;
    int    nrElecEmit;          /* emitted in the last step */
    long long nrEmitted_total, nrAbsorbed_top, nrAbsorbed_bot;
    double N_sup, df_avg, a_rate, MH_std, MH_std_tip;
    double F_avg[3];
    int    neval, fail;
    double integral_error;
    double ramo_current[4];
    double ramo_total;          /* sum over species, A */
    double ramo_integral;       /* sum of I*dt, C */
    double avg_elec_vel[3];
    float  accel_ms, step_ms;
    double t_dev_step, t_dev_accel;               /* device time (CUDA events) of rb2_step and of its pair kernels, summed */
    double t_emission, t_md_step, t_remove, t_io; /* wall-clock seconds spent so far in: ptr_Do_Emission, rb2_step,
                                                     rb2_remove_marked, the text/binary writers */
    long long nrIonizations_total, nrRecombinations_total; /* collisions (COLLISION_MODE 1, 2) */
    double t_collisions, t_dev_collisions;        /* wall clock / device time of Do_Collisions, summed */
    double t_em_quad, t_em_mh, t_em_add;          /* planar field emission: supply quadrature, sampler, accept + insert */
    long long n_candidates_total;                 /* emission candidates (sum of N_round) */
} rh_state;

void *rh_create_from_dir(const char *dir, int write_files, unsigned long long seed, int max_particles);
void *rh_create(const rh_setup *setup);
int   rh_init(void *sim);
int   rh_step(void *sim, int step);
int   rh_run(void *sim, int first_step, int n_steps);
int   rh_get_state(void *sim, rh_state *out);
int   rh_steps_in_input(void *sim);
/* ramo_current_emit(1:n_sec, 1) of the last step (needs WRITE_RAMO_SEC in the deck or rh_setup.ramo_sections > 0) */
int   rh_get_ramo_sections(void *sim, int n_sec, double *out);
/* Host-side switches: "photo_serial" (1: the photo-emission loop runs attempt by attempt, the reference's literal
 * sequence; 0: speculative device batches, same decisions), "ramo_sections" (before rh_init). */
int   rh_set_option(void *sim, const char *name, double value);
void  rh_destroy(void *sim);
const char *rh_last_error(void *sim);

/* samplers / quadrature, exposed for the parity tests (call after rh_init) */
int rh_cuba_integrate(void *sim, int kind, double *integral, double *error, int *neval, int *fail);
int rh_mh_rectangle_J(void *sim, double *df_out, double *F_out, double *pos_out);
int rh_mh_rectangle_J_batch(void *sim, int M, double *df_out, double *F_out, double *pos_out);
int rh_mh_rectangle_J_thermo(void *sim, double *pos_out);
/* mh_batch = 2 only: all M thermal-field chains in lock-step on the device (rb2_mh_planar kind 2) */
int rh_mh_rectangle_J_thermo_batch(void *sim, int M, double *pos_out, int *ok_out);
int rh_metro_algo_tip_v3(void *sim, int ndim, double *xi, double *phi, double *eta_f, double *df_cur, double *par_pos);
int rh_metro_algo_tip_v3_batch(void *sim, int M, int ndim, double *eta_f, double *df_cur, double *par_pos);
int rh_tip_supply_grid(void *sim, int nr_xi, int nr_phi, double *n_s, double *F_avg);
int rh_do_emission(void *sim, int step, int *n_emitted);
double rh_w_theta_xy(void *sim, const double *pos, int *sec);
double rh_kevin_jgtf_v2(double F, double T, double w_theta);

#ifdef __cplusplus
}
#endif
#endif

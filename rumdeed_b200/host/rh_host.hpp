// rh_host.hpp -- C++ host side above the C ABI: a mirror of the parts of the RUMDEED Fortran
// host that surround the hot path (the reference's toolchain, Fortran, is absent from the
// build image; INTEGRATION.md shows the Fortran binding).  Same names, argument meaning and
// error behaviour as the reference:
//   input namelist + work / laser files   src/main.F90:298-411, src/mod_work_function.F90:53-160
//   emission plugin interface             src/mod_global.F90:445-506 (init / do-emission / clean-up)
//   planar FE (mode 10), tip FE (mode 3), thermal-field (mode 9), photo (mode 1)
//   main loop + output writers            src/main.F90:175-266, src/mod_pair.F90:774-867
// All field evaluations, the particle store and the time step go through include/rumdeed_b200.h.
#pragma once

#include <stdint.h>
#include <stdio.h>

#include <functional>
#include <string>
#include <vector>

#include "rumdeed_b200.h"

namespace rh {

// ---- constants, src/mod_global.F90:26-75 ----------------------------------------------------
constexpr double pi = 3.141592653589793238462643383279502884197169399375105820974944592307816406286;
constexpr double h_planck = 6.62607015e-34;
constexpr double k_b = 1.380649e-23;
constexpr double c_light = 299792458.0;
constexpr double mu_0 = 1.25663706212e-6;
constexpr double epsilon_0 = 1.0 / (mu_0 * (c_light * c_light));
constexpr double h_bar = h_planck / (2.0 * pi);
constexpr double m_0 = 9.1093837015e-31;
constexpr double q_0 = 1.602176634e-19;
constexpr double q_02 = q_0 * q_0;
constexpr double length_scale = 1.0e-9;
constexpr double time_scale = 1.0e-12;
constexpr double cur_scale = 1.0;
constexpr double P_ntp = 101325.0;
constexpr int MAX_EMITTERS = 1;
constexpr int MAX_SECTIONS = 96 * 96;
constexpr int MAX_PARTICLES = 5000000;

// emission modes, src/mod_global.F90 (EMISSION_*)
enum { EMISSION_PHOTO = 1, EMISSION_TIP = 3, EMISSION_FIELD_THERMO = 9, EMISSION_FIELD_V2 = 10 };
enum { EMIT_CIRCLE = 1, EMIT_RECTANGLE = 2 };
enum { species_elec = 1, species_ion = 2, species_atom = 3 };

// ---- random numbers (the reference uses the compiler's RANDOM_NUMBER; unpinned) -------------
struct Rng {
    uint64_t s[4];
    void seed(uint64_t v);
    uint64_t next();
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    void box_muller(const double mean[2], const double std[2], double out[2]);  // src/mod_global.F90:578-595
    int poisson(double lambda);                                                 // src/mod_global.F90:600-643
};

// ---- the /input/ namelist + derived globals (src/mod_global.F90, src/main.F90:298-411) -------
struct Globals {
    double V_s = 0.0, V_d = 0.0, d = 0.0, E_z = 0.0;
    double box_dim[3] = {0, 0, 0};
    double time_step = 0.0, time_step2 = 0.0;
    int steps = 0, nrEmit = 1, emission_mode = 0;
    bool image_charge = true;
    int N_ic_max = 0;
    int collision_mode = 0, collision_delay = 0, ion_life_time = 100000000;  // src/mod_global.F90:207-210
    double T_temp = 293.15, P_abs = 1.0;
    double emitters_pos[3] = {0, 0, 0}, emitters_dim[3] = {0, 0, 0};
    int emitters_type = 0, emitters_delay = 0;
    int planes_N = 10;
    double planes_z[RB2_PLANES_MAX] = {5.0, 10.0, 25.0, 50.0, 75.0, 100.0, 125.0, 250.0, 500.0, 750.0};
    bool mh_batch = false;
    bool mh_device = false;  // lock-step chains run by rb2_mh_planar (implies mh_batch)
    bool mh_host = false;    // mh_batch = .false.: the serial chains as a HOST loop of M = 1 field calls (default: one device kernel)
    bool write_ramo_sec = false;                             // src/mod_global.F90:352: ramo_current.bin per section
    bool write_position_file = false;                        // src/mod_global.F90:353
    bool sample_elec_file = false; int sample_elec_rate = 500;  // src/mod_global.F90:360-361
    int cuba_method = 2;
    double cuba_epsabs = 0.5, cuba_epsrel = 1.0e-3;
    int cuba_mineval = 1000, cuba_maxeval = 5000000;
    int max_particles = MAX_PARTICLES;  // capacity of the device store (MAX_PARTICLES in the reference)
    uint64_t seed = 0;                  // 0: from /dev/urandom like the reference (src/main.F90:747-793)
};

// ---- work function, src/mod_work_function.F90 --------------------------------------------------
struct WorkFunction {
    int type = 1;  // WORK_CHECKBOARD
    int y_num = 1, x_num = 1;
    std::vector<double> w_theta_arr;  // [y_num][x_num]
    int read(const std::string &path, std::string &err);
    double w_theta_xy(const Globals &g, const double pos[3], int *sec) const;
};

// ---- laser file, src/mod_photo_emission.f90:56-147 ----------------------------------------------
struct Laser {
    int gauss_mode = 2, laser_mode = 1, photon_mode = 1;
    double laser_energy = 4.7, laser_variation = 0.0;
    double gauss_center = 0.0, gauss_width = 1.0, gauss_amplitude = 0.0;
    int read(const std::string &path, std::string &err);
};

struct Sim;

// procedure pointers bound by Init_* (src/mod_global.F90:496-506)
struct Pointers {
    std::function<int(Sim &, int)> ptr_Do_Emission;
    std::function<int(Sim &)> ptr_Clean_Up;
    const char *name = "";
};

struct QuadResult {
    double integral = 0.0, error = 0.0;
    int neval = 0, fail = 0;
    double F_avg[3] = {0, 0, 0};
};

struct StepLog {  // what the writers print per step
    int nrElecEmit = 0;
    double N_sup = 0.0, df_avg = 0.0;
    double F_avg[3] = {0, 0, 0};
    int neval = 0, fail = 0;
    double integral_error = 0.0;
};

struct Sim {
    Globals g;
    WorkFunction work;
    Laser laser;
    Pointers ptr;
    Rng rng;
    Rng rng_photo;                    // candidate spots of the photo-emission loop (own stream, see rh_emission.cpp)
    std::vector<double> photo_cand;   // ... drawn but not consumed yet
    size_t photo_head = 0;
    bool photo_serial = false;        // run Do_Photo_Emission_Rectangle attempt by attempt (reference sequence; tests)
    rb2_config cfg{};
    rb2_counts counts{};
    rb2_step_result last{};
    StepLog slog;
    std::string dir, out_dir, err;
    bool write_files = false;
    int cur_step = 0;
    double cur_time = 0.0;
    // sampler state (module variables of the reference)
    double a_rate = 1.0, MH_std = 0.0125;         // src/mod_field_emission_v2.F90:66-67
    double MH_std_tip = 1.0, a_rate_tip = 0.5;    // src/mod_emission_tip.f90:50
    double residual = 0.0;
    long long nrEmitted_total = 0, nrAbsorbed_top = 0, nrAbsorbed_bot = 0;
    double ramo_integral = 0.0;  // sum over steps of I*dt (Shockley-Ramo charge)
    double t_emission = 0.0, t_md_step = 0.0, t_remove = 0.0, t_io = 0.0, t_dev_step = 0.0, t_dev_accel = 0.0;  // wall-clock seconds per phase of the main loop
    // tip geometry scalars (src/mod_hyperboloid_tip.f90)
    double d_tip = 0, R_base = 0, h_tip = 0, a_foci = 0, eta_1 = 0, theta_tip = 0, r_tip = 0, max_xi = 0, shift_z = 0;
    double pre_fac_E_tip = 0, pre_fac_E_tip_unit_voltage = 0;
    // output units
    FILE *ud_ramo = nullptr, *ud_emit = nullptr, *ud_absorb = nullptr, *ud_absorb_top = nullptr, *ud_absorb_bot = nullptr;
    FILE *ud_field = nullptr, *ud_integrand = nullptr, *ud_volt = nullptr, *ud_density_emit = nullptr;
    FILE *ud_pos = nullptr;  // out/position.bin (Write_Position)
    FILE *ud_ramo_sec = nullptr;  // out/ramo_current.bin (Write_Ramo_Current with write_ramo_sec, src/mod_pair.F90:822-826)
    FILE *ud_density_emit_elec = nullptr, *ud_density_emit_ion = nullptr, *ud_density_emit_atom = nullptr;  // src/mod_pair.F90:92, :103, :114
    std::vector<double> ramo_current_emit;  // (MAX_SECTIONS, MAX_EMITTERS) of the last step when write_ramo_sec / sections are on
    int ramo_sections = 0;                  // size of the device table (work.y_num * work.x_num), 0 = off
    FILE *ud_coll = nullptr, *ud_ionization_data = nullptr, *ud_recombination_data = nullptr, *ud_density_absorb_recom = nullptr,
         *ud_absorb_recom = nullptr;  // collision outputs (src/main.F90:603-641, :719)
    long long nrIonizations_total = 0, nrRecombinations_total = 0;
    int recom_counts[3] = {0, 0, 0};  // nrPart/nrElec/nrIon_remove_recom since the last Remove_Particles
    double t_collisions = 0.0, t_dev_collisions = 0.0;
    double t_em_quad = 0.0, t_em_mh = 0.0, t_em_add = 0.0;  // planar field emission: supply quadrature, sampler, accept + insert
    long long n_candidates_total = 0;                       // sum of N_round
    FILE *ud_density_absorb_top = nullptr, *ud_density_absorb_bot = nullptr, *planes_ud[RB2_PLANES_MAX] = {nullptr};
    std::vector<double> scratch_pts, scratch_fld, scratch_ez;
    // tip supply grid (Tip_Supply_Grid): geometry-only quantities, computed once
    std::vector<double> tip_grid_pts, tip_grid_nrm, tip_grid_area;
    int tip_grid_key[2] = {0, 0};
    bool tip_grid_on_device = false;  // rb2_tip_supply_set_grid done for the cached grid
    double tip_grid_geom[4] = {0, 0, 0, 0};

    int fail(const std::string &m) { err = m; return -1; }
    int check(int rc, const char *where);

    // mod_verlet / mod_pair through the C ABI
    int Calc_Field_at(const double pos[3], double field[3]);
    int Calc_Field_at_Batch(int M, const double *pos_in, double *field_out);
    int Calc_Field_at_Surface(int M, const double *pos_in, double *field_out);
    int Add_Particle(const double par_pos[3], const double par_vel[3], int species, int step, int emit, int life, int sec);
    int Add_Particles(int k, const double *pos, const double *vel, int species, int step, int emit, int life, const int *sec);
    // the density_emit*.bin records + host counters of one ACCEPTED particle (src/mod_pair.F90:85-123)
    void record_added(const double pos[3], int species, int emit, int sec);
    // geometry helpers (src/mod_hyperboloid_tip.f90:25-112, 156-163)
    void xyz_corr(double xi, double eta, double phi, double out[3]) const;
    void surface_normal(const double pos[3], double out[3]) const;
    double Field_normal(const double pos[3], const double field[3]) const;
    double Tip_Area(double xi_1, double xi_2, double phi_1, double phi_2) const;
};

// ---- input ------------------------------------------------------------------------------------------
int Read_Input_Variables(const std::string &path, Globals &g, std::string &err);

// ---- life cycle (src/main.F90:100-151, :233-266) ----------------------------------------------------
int Init(Sim &s);       // allocate the device store, bind the emission plugin, open the output files
int Clean_up(Sim &s);
int Step(Sim &s, int step);  // one iteration of the main loop, src/main.F90:175-219

// ---- emission plugins ---------------------------------------------------------------------------------
int Init_Field_Emission_v2(Sim &s);
int Init_Emission_Tip(Sim &s);
int Init_Field_Thermo_Emission(Sim &s);
int Init_Photo_Emission(Sim &s);

// ---- collisions (src/mod_collisions.F90; rh_collisions.cpp) --------------------------------------------
int Init_Collisions(Sim &s);          // Read_Cross_Section + rb2_collisions_init + output files
int Do_Collisions(Sim &s, int step);  // src/mod_verlet.F90:164-170

// FN helpers (src/mod_field_emission_v2.F90:515-625)
double v_y(const Sim &s, double F, double w_theta);
double t_y(const Sim &s, double F, double w_theta);
double Escape_Prob_log(const Sim &s, double F, double w_theta);
double Elec_Supply_log(const Sim &s, double F, double w_theta);
double Elec_Supply_V2(const Sim &s, double F, double w_theta);
double Get_Kevin_Jgtf_v2(double F, double T, double w_theta);  // src/mod_kevin_rjgtf_v2.f90:56

// samplers exposed for the tests
int Metropolis_Hastings_rectangle_J(Sim &s, int emit, double *df_out, double *F_out, double pos_out[3]);
int Metropolis_Hastings_rectangle_J_batch(Sim &s, int M, int emit, double *df_out, double *F_out, double *pos_out);
int Metropolis_Hastings_rectangle_J_thermo(Sim &s, int emit, double pos_out[3]);
int Metropolis_Hastings_rectangle_J_thermo_batch(Sim &s, int M, double *pos_out, int *ok_out);
int Metro_algo_tip_v3(Sim &s, int ndim, double *xi, double *phi, double *eta_f, double *df_cur, double par_pos[3]);
int Metro_algo_tip_v3_batch(Sim &s, int M, int ndim, double *eta_f_out, double *df_out, double *pos_out);
int Tip_Supply_Grid(Sim &s, int nr_xi, int nr_phi, double *n_s, double *F_avg);

// Cuba_Integrate stand-in (src/mod_cuba_integration.F90:95-169): randomised lattice rule honouring
// cuba_epsabs / cuba_epsrel / cuba_mineval / cuba_maxeval; integrand evaluated in batches on the device.
enum { SUPPLY_FE = 1, SUPPLY_GTF = 2 };
int Cuba_Integrate(Sim &s, int kind, int emit, QuadResult *out);

}  // namespace rh

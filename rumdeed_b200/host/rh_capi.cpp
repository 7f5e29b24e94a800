// rh_capi.cpp -- C entry points of librumdeed_host.so (include/rumdeed_host.h) and the `rumdeed_b200_run`
// executable (program RUMDEED, src/main.F90:31-296).
#include <signal.h>
#include <string.h>

#include <new>

#include "rh_host.hpp"
#include "rumdeed_host.h"

using namespace rh;

extern "C" {

void *rh_create_from_dir(const char *dir, int write_files, unsigned long long seed, int max_particles)
{
    Sim *s = new (std::nothrow) Sim();
    if (!s) return nullptr;
    s->dir = dir ? dir : ".";
    s->out_dir = s->dir + "/out";
    s->write_files = write_files != 0;
    if (Read_Input_Variables(s->dir + "/input", s->g, s->err)) return s;  // error is reported by rh_init
    if (seed) s->g.seed = seed;
    if (max_particles > 0) s->g.max_particles = max_particles;
    {  // the work file is read here already (Init_* only reads it when it is still missing); the tip has none
        std::string ignore;
        if (s->work.read(s->dir + "/work", ignore)) s->work.w_theta_arr.clear();
    }
    return s;
}

void *rh_create(const rh_setup *u)
{
    Sim *s = new (std::nothrow) Sim();
    if (!s || !u) return s;
    Globals &g = s->g;
    g.emission_mode = u->emission_mode;
    g.V_s = g.V_d = u->V_s;
    for (int k = 0; k < 3; ++k) { g.box_dim[k] = u->box_dim[k]; g.emitters_pos[k] = u->emitters_pos[k]; g.emitters_dim[k] = u->emitters_dim[k]; }
    g.d = g.box_dim[2];
    g.E_z = -1.0 * g.V_d / g.d;
    g.time_step = u->time_step;
    g.time_step2 = u->time_step * u->time_step;
    g.steps = 1;
    g.image_charge = u->image_charge != 0;
    g.N_ic_max = u->N_ic_max;
    g.emitters_type = u->emitters_type;
    g.emitters_delay = u->emitters_delay;
    g.T_temp = u->T_temp;
    g.mh_batch = u->mh_batch > 0;
    g.mh_device = u->mh_batch >= 2;
    g.mh_host = u->mh_batch < 0;
    g.planes_N = u->planes_N;
    for (int k = 0; k < 10; ++k) g.planes_z[k] = u->planes_z[k];
    if (u->cuba_epsabs > 0) g.cuba_epsabs = u->cuba_epsabs;
    if (u->cuba_epsrel > 0) g.cuba_epsrel = u->cuba_epsrel;
    if (u->cuba_mineval > 0) g.cuba_mineval = u->cuba_mineval;
    if (u->cuba_maxeval > 0) g.cuba_maxeval = u->cuba_maxeval;
    if (u->max_particles > 0) g.max_particles = u->max_particles;
    g.seed = u->seed;
    s->ramo_sections = u->ramo_sections > 0 ? u->ramo_sections : 0;
    if (u->work_w_theta && u->work_y_num > 0 && u->work_x_num > 0) {
        s->work.type = 1;
        s->work.y_num = u->work_y_num;
        s->work.x_num = u->work_x_num;
        s->work.w_theta_arr.assign(u->work_w_theta, u->work_w_theta + (size_t)u->work_y_num * u->work_x_num);
    } else {
        s->work.y_num = s->work.x_num = 1;
        s->work.w_theta_arr.assign(1, 2.0);
    }
    if (u->laser_gauss_mode) {
        s->laser.gauss_mode = u->laser_gauss_mode; s->laser.laser_mode = u->laser_mode; s->laser.photon_mode = u->photon_mode;
        s->laser.laser_energy = u->laser_energy; s->laser.laser_variation = u->laser_variation;
        s->laser.gauss_center = u->gauss_center; s->laser.gauss_width = u->gauss_width; s->laser.gauss_amplitude = u->gauss_amplitude;
    }
    return s;
}

int rh_init(void *p)
{
    Sim *s = (Sim *)p;
    if (!s) return -1;
    if (!s->err.empty()) return -1;
    return Init(*s);
}

int rh_step(void *p, int step) { Sim *s = (Sim *)p; return s ? Step(*s, step) : -1; }

int rh_run(void *p, int first_step, int n_steps)
{
    Sim *s = (Sim *)p;
    if (!s) return -1;
    for (int i = first_step; i < first_step + n_steps; ++i)
        if (Step(*s, i)) return -1;
    return 0;
}

int rh_get_state(void *p, rh_state *o)
{
    Sim *s = (Sim *)p;
    if (!s || !o) return -1;
    memset(o, 0, sizeof(*o));
    o->step = s->cur_step;
    o->nrPart = s->counts.nrPart; o->nrElec = s->counts.nrElec; o->nrIon = s->counts.nrIon; o->nrID = s->counts.nrID;
    o->nrElecEmit = s->slog.nrElecEmit;
    o->nrEmitted_total = s->nrEmitted_total; o->nrAbsorbed_top = s->nrAbsorbed_top; o->nrAbsorbed_bot = s->nrAbsorbed_bot;
    o->N_sup = s->slog.N_sup; o->df_avg = s->slog.df_avg; o->a_rate = s->a_rate; o->MH_std = s->MH_std; o->MH_std_tip = s->MH_std_tip;
    for (int k = 0; k < 3; ++k) { o->F_avg[k] = s->slog.F_avg[k]; o->avg_elec_vel[k] = s->last.avg_elec_vel[k]; }
    o->neval = s->slog.neval; o->fail = s->slog.fail; o->integral_error = s->slog.integral_error;
    for (int k = 0; k < 4; ++k) o->ramo_current[k] = s->last.ramo_current[k];
    o->ramo_total = s->last.ramo_current[1] + s->last.ramo_current[2] + s->last.ramo_current[3];
    o->ramo_integral = s->ramo_integral;
    o->accel_ms = s->last.accel_ms; o->step_ms = s->last.step_ms;
    o->t_dev_step = s->t_dev_step; o->t_dev_accel = s->t_dev_accel;
    o->t_emission = s->t_emission; o->t_md_step = s->t_md_step; o->t_remove = s->t_remove; o->t_io = s->t_io;
    o->nrIonizations_total = s->nrIonizations_total; o->nrRecombinations_total = s->nrRecombinations_total;
    o->t_collisions = s->t_collisions; o->t_dev_collisions = s->t_dev_collisions;
    o->t_em_quad = s->t_em_quad; o->t_em_mh = s->t_em_mh; o->t_em_add = s->t_em_add; o->n_candidates_total = s->n_candidates_total;
    return 0;
}

int rh_steps_in_input(void *p) { Sim *s = (Sim *)p; return s ? s->g.steps : 0; }

int rh_get_ramo_sections(void *p, int n_sec, double *out)
{
    Sim *s = (Sim *)p;
    if (!s || !out || n_sec < 1) return -1;
    if (s->ramo_sections < 1) { s->err = "per-section Ramo current is off (WRITE_RAMO_SEC / ramo_sections)"; return -1; }
    for (int k = 0; k < n_sec; ++k) out[k] = k < (int)s->ramo_current_emit.size() ? s->ramo_current_emit[(size_t)k] : 0.0;
    return 0;
}

int rh_set_option(void *p, const char *name, double value)
{
    Sim *s = (Sim *)p;
    if (!s || !name) return -1;
    if (!strcmp(name, "photo_serial")) s->photo_serial = value != 0.0;
    else if (!strcmp(name, "ramo_sections")) s->ramo_sections = value > 0 ? (int)value : 0;
    else { s->err = std::string("unknown host option ") + name; return -1; }
    return 0;
}

void rh_destroy(void *p)
{
    Sim *s = (Sim *)p;
    if (!s) return;
    Clean_up(*s);
    delete s;
}

const char *rh_last_error(void *p) { Sim *s = (Sim *)p; return s ? s->err.c_str() : "null simulation"; }

int rh_cuba_integrate(void *p, int kind, double *integral, double *error, int *neval, int *fail)
{
    Sim *s = (Sim *)p;
    QuadResult q;
    if (!s || Cuba_Integrate(*s, kind, 1, &q)) return -1;
    if (integral) *integral = q.integral;
    if (error) *error = q.error;
    if (neval) *neval = q.neval;
    if (fail) *fail = q.fail;
    return 0;
}
int rh_mh_rectangle_J(void *p, double *df, double *F, double *pos) { return Metropolis_Hastings_rectangle_J(*(Sim *)p, 1, df, F, pos); }
int rh_mh_rectangle_J_batch(void *p, int M, double *df, double *F, double *pos) { return Metropolis_Hastings_rectangle_J_batch(*(Sim *)p, M, 1, df, F, pos); }
int rh_mh_rectangle_J_thermo(void *p, double *pos) { return Metropolis_Hastings_rectangle_J_thermo(*(Sim *)p, 1, pos); }
int rh_mh_rectangle_J_thermo_batch(void *p, int M, double *pos, int *ok)
{
    Sim &s = *(Sim *)p;
    if (!s.g.mh_device) { s.err = "rh_mh_rectangle_J_thermo_batch needs mh_batch = 2 (device-resident chains)"; return -2; }
    return Metropolis_Hastings_rectangle_J_thermo_batch(s, M, pos, ok);
}
int rh_metro_algo_tip_v3(void *p, int ndim, double *xi, double *phi, double *eta_f, double *df_cur, double *par_pos)
{
    return Metro_algo_tip_v3(*(Sim *)p, ndim, xi, phi, eta_f, df_cur, par_pos);
}
int rh_metro_algo_tip_v3_batch(void *p, int M, int ndim, double *eta_f, double *df_cur, double *par_pos)
{
    return Metro_algo_tip_v3_batch(*(Sim *)p, M, ndim, eta_f, df_cur, par_pos);
}
int rh_tip_supply_grid(void *p, int nr_xi, int nr_phi, double *n_s, double *F_avg) { return Tip_Supply_Grid(*(Sim *)p, nr_xi, nr_phi, n_s, F_avg); }
int rh_do_emission(void *p, int step, int *n_emitted)
{
    Sim *s = (Sim *)p;
    if (!s || s->ptr.ptr_Do_Emission(*s, step)) return -1;
    s->nrEmitted_total += s->slog.nrElecEmit;
    if (n_emitted) *n_emitted = s->slog.nrElecEmit;
    return 0;
}
double rh_w_theta_xy(void *p, const double *pos, int *sec)
{
    Sim *s = (Sim *)p;
    if (!s || s->work.w_theta_arr.empty()) return 0.0;
    return s->work.w_theta_xy(s->g, pos, sec);
}
double rh_kevin_jgtf_v2(double F, double T, double w) { return Get_Kevin_Jgtf_v2(F, T, w); }

}  // extern "C"

#ifdef RH_MAIN
// program RUMDEED, src/main.F90:31-296: read the input, run the time loop, clean up.  SIGINT /
// SIGUSR1 end the loop cleanly (src/main.F90:214-217).
static volatile sig_atomic_t g_stop = 0;
static void on_signal(int) { g_stop = 1; }

int main(int argc, char **argv)
{
    const char *dir = argc > 1 ? argv[1] : ".";
    unsigned long long seed = argc > 2 ? strtoull(argv[2], nullptr, 10) : 0;
    int steps_override = argc > 3 ? atoi(argv[3]) : 0;
    int max_particles = argc > 4 ? atoi(argv[4]) : 0;
    printf("RUMDEED (B200 hot path): Reading input values\n");
    void *sim = rh_create_from_dir(dir, 1, seed, max_particles);
    printf("RUMDEED: Initialzing\n");
    if (rh_init(sim)) { fprintf(stderr, "%s\n", rh_last_error(sim)); return 1; }
    signal(SIGINT, on_signal);
    signal(SIGUSR1, on_signal);
    const int steps = steps_override > 0 ? steps_override : rh_steps_in_input(sim);
    printf("RUMDEED: Starting main loop\n");
    for (int i = 1; i <= steps && !g_stop; ++i) {
        if (rh_step(sim, i)) { fprintf(stderr, "%s\n", rh_last_error(sim)); rh_destroy(sim); return 1; }
        if (steps >= 10 && i % (steps / 10) == 0) {
            rh_state st;
            rh_get_state(sim, &st);
            printf("RUMDEED: %3d%% step %d nrElec %d I = %.4E A\n", (int)(100.0 * i / steps), i, st.nrElec, st.ramo_total);
            fflush(stdout);
        }
    }
    rh_state st;
    rh_get_state(sim, &st);
    printf("RUMDEED: Main loop finished: emitted %lld absorbed top %lld bot %lld, %d electrons in the gap\n", st.nrEmitted_total,
           st.nrAbsorbed_top, st.nrAbsorbed_bot, st.nrElec);
    printf("RUMDEED: wall clock per phase [s]: emission %.3f  MD step %.3f  removal %.3f  writers %.3f  (%d steps); device time of the MD steps %.3f, of their pair kernels %.3f\n",
           st.t_emission, st.t_md_step, st.t_remove, st.t_io, st.step, st.t_dev_step, st.t_dev_accel);
    if (st.n_candidates_total)
        printf("RUMDEED: emission split [s]: supply quadrature %.3f  sampler %.3f  accept + insert %.3f; %.1f candidates per step\n",
               st.t_em_quad, st.t_em_mh, st.t_em_add, (double)st.n_candidates_total / (st.step > 0 ? st.step : 1));
    if (st.nrIonizations_total || st.nrRecombinations_total || st.t_collisions > 0.0)
        printf("RUMDEED: collisions: %lld ionisations, %lld recombinations, %d ions in the gap; wall clock %.3f s, device %.3f s\n",
               st.nrIonizations_total, st.nrRecombinations_total, st.nrIon, st.t_collisions, st.t_dev_collisions);
    {
        double plans = 0.0, plan_ms = 0.0, replays = 0.0;
        if (!rb2_get_stat("sym_plans", &plans) && !rb2_get_stat("sym_plan_ms", &plan_ms) && !rb2_get_stat("graph_replays", &replays))
            printf("RUMDEED: pair-symmetric work-unit lists built %.0f times (%.3f s on the host), %.0f step-graph replays\n", plans,
                   plan_ms * 1e-3, replays);
    }
    rh_destroy(sim);
    printf("RUMDEED: Program finished\n");
    return 0;
}
#endif

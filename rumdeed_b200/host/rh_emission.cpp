// rh_emission.cpp -- the emission plugins (init / do-emission / clean-up) of the host mirror.
// Every surface-field evaluation goes to the device through rb2_field_batch; emitted electrons are
// appended to the device store with rb2_add_particles straight away, which gives the serial
// samplers exactly the reference's semantics (each chain sees the electrons emitted before it).
#include <math.h>
#include <string.h>

#include <algorithm>

#include <chrono>

#include "rh_host.hpp"

namespace rh {

// ---- Fowler-Nordheim helpers, src/mod_field_emission_v2.F90:36-47, :515-625 ----------------------
static const double a_FN = q_02 / (16.0 * (pi * pi) * h_bar);
static const double b_FN = -4.0 / (3.0 * h_bar) * sqrt(2.0 * m_0 * q_0);
static const double l_const = q_0 / (4.0 * pi * epsilon_0);

double v_y(const Sim &s, double F, double w_theta)
{
    if (s.g.image_charge) {
        double l = l_const * (-1.0 * F) / (w_theta * w_theta);
        if (l > 1.0) l = 1.0;
        return 1.0 - l + 1.0 / 6.0 * l * log(l);
    }
    return 1.0;
}
double t_y(const Sim &s, double F, double w_theta)
{
    if (s.g.image_charge) {
        double l = l_const * (-1.0 * F) / (w_theta * w_theta);
        if (l > 1.0) l = 1.0;
        return 1.0 + l * (1.0 / 9.0 - 1.0 / 18.0 * log(l));
    }
    return 1.0;
}
double Escape_Prob_log(const Sim &s, double F, double w_theta)
{
    const double sw = sqrt(w_theta);
    return b_FN * (sw * sw * sw) * v_y(s, F, w_theta) / (-1.0 * F);
}
double Elec_Supply_log(const Sim &s, double F, double w_theta)
{
    return 2.0 * log(-1.0 * F) - 2.0 * log(t_y(s, F, w_theta)) - log(w_theta);
}
double Elec_Supply_V2(const Sim &s, double F, double w_theta)
{
    const double t = t_y(s, F, w_theta);
    return (s.g.time_step / q_0) * a_FN / ((t * t) * w_theta) * (F * F);
}

// ---- Jensen's GTF current density, src/mod_kevin_rjgtf_v2.f90:32-175 -------------------------------
static double Nns(double n, double s)
{
    if (n == 1.0) return (s + 1.0) * exp(-s);
    const double x = n * n, y = 1.0 / x, z = (n - 1.0) * s;
    double sng;
    if (fabs(z) > 1.0e-5) sng = (x + 1.0) * (x * exp(-s) - exp(-n * s)) / (x - 1.0);
    else sng = (0.5 * (x + 1.0) * exp(-s) / (n + 1.0)) * ((1.0 - n) * s * s + 2.0 * (1.0 + n) + 2.0 * s);
    const double sn = -x * (0.10593434 * x + 0.35506593);
    const double sd = -y * (0.10593434 * y + 0.35506593);
    const double sfn = x * exp(-s);
    return std::max(sng + sn * exp(-n * s) + x * sd * exp(-s), sfn);
}
double Get_Kevin_Jgtf_v2(double F, double T, double w_theta)
{
    const double kpi = 3.14159265358979324, kb = 1.0 / 11604.50635, hbar = 0.6582119571, c = 299.7924580;
    const double mo = 5.685630103, afs = 1.0 / 137.035999084, Qo = afs * hbar * c / 4.0;
    const double cm = 1.0e7, Amp = 6.241509074e3;
    const double Arld = (mo * (kb * kb) / (2.0 * (kpi * kpi) * (hbar * hbar * hbar))) * (cm * cm) / Amp;
    const double chem = 7.0;
    const double Fo = fabs(F) * 1.0e-9, To = T, Phi = w_theta;
    if (Fo < 1.0e-9) return 0.0;
    const double yo = sqrt(4.0 * Qo * Fo) / Phi;
    const double phix = Phi - sqrt(4.0 * Qo * Fo);
    const double ty = 1.0 + (yo * yo) * (1.0 - log(yo)) / 9.0;
    const double vy = 1.0 - (yo * yo) * (3.0 - log(yo)) / 3.0;
    const double Tmin = (hbar * Fo / (4.0 * kb * ty)) * sqrt(2.0 / (mo * Phi));
    const double Tmax = hbar * Fo / (kb * kpi * sqrt(mo * Phi * yo));
    const double betaT = 1.0 / (kb * To);
    const double betau = (2.0 / (hbar * Fo)) * sqrt(2.0 * mo * Phi) * ty;
    const double betap = (kpi / (hbar * Fo)) * sqrt(mo * Phi * yo);
    const double theto = (4.0 * sqrt(2.0 * mo * (Phi * Phi * Phi)) / (3.0 * hbar * Fo)) * vy;
    double nft, sft;
    if (To < Tmin) { nft = betaT / betau; sft = theto; }
    else if (To > Tmax) { nft = betaT / betap; sft = betap * phix; }
    else {
        const double Ap = 3.0 * (betap + betau) - 6.0 * theto / phix;
        const double Bp = -2.0 * (betap + 2.0 * betau) + 6.0 * theto / phix;
        const double Cp = betau - betaT;
        const double po = (-Bp - sqrt(Bp * Bp - 4.0 * Ap * Cp)) / (2.0 * Ap);
        const double Em = chem + po * phix;
        const double theta = ((1.0 - po) * (1.0 - po)) * (2.0 * po + 1.0) * theto - phix * po * (1.0 - po) * ((1.0 - po) * betau - po * betap);
        nft = 1.0;
        sft = theta + betaT * (Em - chem);
    }
    return (Arld * Nns(nft, sft) * (To * To)) * 1.0e4;
}

// ---- Sim helpers ------------------------------------------------------------------------------------------
int Sim::check(int rc, const char *where)
{
    if (rc != RB2_OK) {
        err = std::string("librumdeed_b200 error in ") + where + ": " + rb2_last_error_string();
        return -1;
    }
    return 0;
}
int Sim::Calc_Field_at(const double pos[3], double field[3]) { return check(rb2_field_batch(1, pos, field), "rb2_field_batch"); }
int Sim::Calc_Field_at_Batch(int M, const double *pos_in, double *field_out)
{
    if (M < 1) return 0;
    return check(rb2_field_batch(M, pos_in, field_out), "rb2_field_batch");
}
// Calc_Field_at_Batch for points ON the planar cathode (z = 0), which is all the planar emission modules ask
// for: with image charges on, E_x = E_y = 0 there and E_z comes from the cheaper rb2_field_surface_z.
int Sim::Calc_Field_at_Surface(int M, const double *pos_in, double *field_out)
{
    if (M < 1) return 0;
    if (!g.image_charge) return check(rb2_field_batch(M, pos_in, field_out), "rb2_field_batch");
    scratch_ez.resize(M);
    if (check(rb2_field_surface_z(M, pos_in, scratch_ez.data()), "rb2_field_surface_z")) return -1;
    for (int k = 0; k < M; ++k) { field_out[3 * k] = 0.0; field_out[3 * k + 1] = 0.0; field_out[3 * k + 2] = scratch_ez[k]; }
    return 0;
}
// The part of Add_Particle that the host keeps: the density_emit*.bin records (src/mod_pair.F90:92, :103, :114, :123)
// and the mirror of nrID / nrPart / nrElec / nrIon.  Only called for particles the store accepted: a particle dropped
// at MAX_PARTICLES writes nothing and does not advance nrID in the reference (:36-43).
void Sim::record_added(const double pos[3], int species, int emit, int sec)
{
    const double p3[3] = {pos[0] / length_scale, pos[1] / length_scale, pos[2] / length_scale};
    FILE *fs = species == species_elec ? ud_density_emit_elec : species == species_ion ? ud_density_emit_ion : ud_density_emit_atom;
    if (fs) {  // pos/length_scale, emit, nrID
        const int tail[2] = {emit, counts.nrID};
        fwrite(p3, sizeof(double), 3, fs);
        fwrite(tail, sizeof(int), 2, fs);
    }
    if (ud_density_emit) {  // pos/length_scale, emit, sec, nrID, species
        const int tail[4] = {emit, sec, counts.nrID, species};
        fwrite(p3, sizeof(double), 3, ud_density_emit);
        fwrite(tail, sizeof(int), 4, ud_density_emit);
    }
    counts.nrID += 1;
    counts.nrPart += 1;
    if (species == species_elec) counts.nrElec += 1; else if (species == species_ion) counts.nrIon += 1;
}
int Sim::Add_Particle(const double par_pos[3], const double par_vel[3], int species, int step, int emit, int life, int sec)
{
    return Add_Particles(1, par_pos, par_vel, species, step, emit, life, &sec);
}
// k calls of Add_Particle in one trip to the device (same order, same ids).  With k > 1 only valid where no field
// evaluation sits between the individual calls, i.e. behind the lock-step samplers (mh_batch), whose fields are all taken
// before the first insertion -- one host/device round trip (~40 us) per step instead of one per emitted electron.
int Sim::Add_Particles(int k, const double *pos, const double *vel, int species, int step, int emit, int life, const int *sec)
{
    if (k < 1) return 0;
    int room = 0;
    if (check(rb2_capacity_left(&room), "rb2_capacity_left")) return -1;
    std::vector<int> sp((size_t)k, species), em((size_t)k, emit), lf((size_t)k, life);
    int rc = check(rb2_add_particles(k, pos, vel, sp.data(), step, em.data(), sec, lf.data()), "rb2_add_particles");
    if (rc) return rc;
    const int accepted = k < room ? k : room;  // the rest was dropped and counted in nrPart_dropped
    for (int i = 0; i < accepted; ++i) {
        const int s_i = sec[i] > MAX_SECTIONS ? MAX_SECTIONS : sec[i];
        record_added(&pos[(size_t)3 * i], species, emit, s_i);
    }
    return 0;
}
void Sim::xyz_corr(double xi, double eta, double phi, double out[3]) const
{
    const double xy = a_foci * sqrt((xi * xi - 1.0) * (1.0 - eta * eta));
    out[0] = xy * cos(phi);
    out[1] = xy * sin(phi);
    out[2] = a_foci * xi * eta + shift_z;
}
void Sim::surface_normal(const double pos[3], double out[3]) const
{
    const double eta_fac = eta_1 / sqrt(1 - eta_1 * eta_1);
    const double div_fac = -1.0 / sqrt(pos[0] * pos[0] + pos[1] * pos[1] + (a_foci * a_foci) * (1 - eta_1 * eta_1));
    const double nx = eta_fac * pos[0] * div_fac, ny = eta_fac * pos[1] * div_fac, nz = 1.0;
    const double nrm = sqrt(nx * nx + ny * ny + nz * nz);
    out[0] = nx / nrm; out[1] = ny / nrm; out[2] = nz / nrm;
}
double Sim::Field_normal(const double pos[3], const double field[3]) const
{
    double u[3];
    surface_normal(pos, u);
    return u[0] * field[0] + u[1] * field[1] + u[2] * field[2];
}
double Sim::Tip_Area(double xi_1, double xi_2, double phi_1, double phi_2) const
{
    const double e2 = eta_1 * eta_1;
    const double fac_1 = xi_1 * sqrt(xi_1 * xi_1 - e2) - e2 * log(xi_1 + sqrt(xi_1 * xi_1 - e2));
    const double fac_2 = xi_2 * sqrt(xi_2 * xi_2 - e2) - e2 * log(xi_2 + sqrt(xi_2 * xi_2 - e2));
    return 0.5 * (a_foci * a_foci) * sqrt(1.0 - e2) * (phi_2 - phi_1) * (fac_2 - fac_1);
}

static void fill_mh_config(const Sim &s, int kind, rb2_mh_config &c);

// ---- surface integration: stand-in for Cuba_Integrate ----------------------------------------------------
// Cuba (Divonne, src/mod_cuba_integration.F90:95-169) is not available.  The replacement is a
// randomised rank-1 lattice (R2 sequence, 8 independent Cranley-Patterson shifts): the estimate is
// the mean over shifts, the error their standard error; the point count doubles until
// error <= max(epsabs, epsrel*|I|) (the reference's stopping rule), bounded by mineval / maxeval.
// The integrand is integrand_cuba_fe_v (src/mod_field_emission_v2.F90:668-745) or
// integrand_cuba_simple (src/mod_field_thermo_emission.F90:394-446), one device batch per level.
int Cuba_Integrate(Sim &s, int kind, int emit, QuadResult *out)
{
    (void)emit;
    const Globals &g = s.g;
    const int K = 8;
    const double a1 = 0.7548776662466927600495088963585286919, a2 = 0.5698402909980532659113999581195686488;
    double shift[K][2], sum[K];
    for (int r = 0; r < K; ++r) { shift[r][0] = s.rng.uniform(); shift[r][1] = s.rng.uniform(); sum[r] = 0.0; }
    const double A = g.emitters_dim[0] * g.emitters_dim[1];
    int n_done = 0, n_next = std::max(16, (g.cuba_mineval + K - 1) / K);
    QuadResult q;
    double fsum[3] = {0, 0, 0};
    for (;;) {
        const int n_new = n_next - n_done, M = n_new * K;
        if (g.mh_device) {
            // the level on the device (rb2_planar_supply_level): nodes, cathode-plane field, integrand and the per-shift
            // sums there; the convergence test below is unchanged
            rb2_mh_config c{};
            fill_mh_config(s, kind == SUPPLY_FE ? 1 : 2, c);
            double part[K], ez = 0.0;
            if (s.check(rb2_planar_supply_level(&c, s.work.w_theta_arr.data(), kind == SUPPLY_FE ? 1 : 2, K, &shift[0][0], n_done, n_new, part, &ez),
                        "rb2_planar_supply_level")) return -1;
            for (int r = 0; r < K; ++r) sum[r] += A * part[r];
            fsum[2] += ez;
        } else {
        s.scratch_pts.resize((size_t)3 * M);
        s.scratch_fld.resize((size_t)3 * M);
        for (int r = 0; r < K; ++r)
            for (int k = 0; k < n_new; ++k) {
                const double kk = (double)(n_done + k + 1);
                double u = kk * a1 + shift[r][0], v = kk * a2 + shift[r][1];
                u -= floor(u); v -= floor(v);
                double *p = &s.scratch_pts[(size_t)3 * (r * n_new + k)];
                p[0] = g.emitters_pos[0] + u * g.emitters_dim[0];
                p[1] = g.emitters_pos[1] + v * g.emitters_dim[1];
                p[2] = 0.0;
            }
        if (s.Calc_Field_at_Surface(M, s.scratch_pts.data(), s.scratch_fld.data())) return -1;
        for (int r = 0; r < K; ++r)
            for (int k = 0; k < n_new; ++k) {
                const size_t o = (size_t)3 * (r * n_new + k);
                const double *f = &s.scratch_fld[o], *p = &s.scratch_pts[o];
                fsum[0] += f[0]; fsum[1] += f[1]; fsum[2] += f[2];
                double ff = 0.0;
                if (f[2] < 0.0) {
                    const double w = s.work.w_theta_xy(g, p, nullptr);
                    ff = (kind == SUPPLY_FE) ? Elec_Supply_V2(s, f[2], w) : Get_Kevin_Jgtf_v2(f[2], g.T_temp, w) * (g.time_step / q_0);
                }
                sum[r] += A * ff;
            }
        }
        n_done = n_next;
        q.neval = n_done * K;
        double mean = 0.0, var = 0.0;
        for (int r = 0; r < K; ++r) mean += sum[r] / n_done;
        mean /= K;
        for (int r = 0; r < K; ++r) { const double dlt = sum[r] / n_done - mean; var += dlt * dlt; }
        q.integral = mean;
        q.error = sqrt(var / (K - 1) / K);
        const double tol = std::max(g.cuba_epsabs, g.cuba_epsrel * fabs(mean));
        if (q.error <= tol) { q.fail = 0; break; }
        if (q.neval * 2 > g.cuba_maxeval) { q.fail = 1; break; }
        n_next = n_done * 2;
    }
    for (int k = 0; k < 3; ++k) q.F_avg[k] = fsum[k] / q.neval;
    *out = q;
    return 0;
}

// ---- planar field emission (mode 10), src/mod_field_emission_v2.F90 -----------------------------------------
// check_limits_metro_rec, :1466-1516
static void check_limits_metro_rec(const Globals &g, double pos[3])
{
    const double x_max = g.emitters_pos[0] + g.emitters_dim[0], x_min = g.emitters_pos[0];
    const double y_max = g.emitters_pos[1] + g.emitters_dim[1], y_min = g.emitters_pos[1];
    if (pos[0] > x_max) pos[0] = x_max - (pos[0] - x_max);
    else if (pos[0] < x_min) pos[0] = (x_min - pos[0]) + x_min;
    if (pos[1] > y_max) pos[1] = y_max - (pos[1] - y_max);
    else if (pos[1] < y_min) pos[1] = (y_min - pos[1]) + y_min;
}
// MH_std_update, :603-612 with the constants :70-76
static void MH_std_update(Sim &s, double rate)
{
    s.MH_std = s.MH_std * exp(0.025 * (rate - 0.35));
    if (s.MH_std > 0.1250) s.MH_std = 0.1250;
    else if (s.MH_std < 0.00005) s.MH_std = 0.00005;
}
static const double HUGE_NEG = -1.7976931348623157e308;

// Serial chain, :1122-1265
int Metropolis_Hastings_rectangle_J(Sim &s, int emit, double *df_out, double *F_out, double pos_out[3])
{
    (void)emit;
    const Globals &g = s.g;
    const int ndim = 25 * 8, ndim_first = (int)lround(ndim * 0.25);
    int jump_a = 0, jump_r = 0, count = 0;
    double std[2] = {g.emitters_dim[0] * 0.10, g.emitters_dim[1] * 0.10};
    double cur_pos[3], new_pos[3], field[3];
    for (;;) {
        cur_pos[0] = s.rng.uniform() * g.emitters_dim[0] + g.emitters_pos[0];
        cur_pos[1] = s.rng.uniform() * g.emitters_dim[1] + g.emitters_pos[1];
        cur_pos[2] = 0.0;
        if (s.Calc_Field_at_Surface(1, cur_pos, field)) return -2;
        if (field[2] < 0.0) break;
        if (++count > 10000) {
            *F_out = 1.0; *df_out = HUGE_NEG;
            pos_out[0] = g.emitters_pos[0]; pos_out[1] = g.emitters_pos[1]; pos_out[2] = 0.0;
            fprintf(stderr, " Failed to find spot for emission\n");
            return -1;
        }
    }
    *F_out = field[2];
    double sup_cur = Elec_Supply_log(s, field[2], s.work.w_theta_xy(g, cur_pos, nullptr));
    for (int i = 1; i <= ndim; ++i) {
        if (i > ndim_first) { std[0] = g.emitters_dim[0] * s.MH_std; std[1] = g.emitters_dim[1] * s.MH_std; }
        s.rng.box_muller(cur_pos, std, new_pos);
        new_pos[2] = 0.0;
        check_limits_metro_rec(g, new_pos);
        if (s.Calc_Field_at_Surface(1, new_pos, field)) return -2;
        if (field[2] >= 0.0) { if (i > ndim_first) jump_r++; continue; }
        const double sup_new = Elec_Supply_log(s, field[2], s.work.w_theta_xy(g, new_pos, nullptr));
        const double alpha = sup_new - sup_cur;
        bool accept = sup_new >= sup_cur;
        if (!accept) accept = log(s.rng.uniform()) <= alpha;
        if (accept) {
            memcpy(cur_pos, new_pos, sizeof(cur_pos)); sup_cur = sup_new; *F_out = field[2];
            if (i > ndim_first) jump_a++;
        } else if (i > ndim_first) jump_r++;
    }
    if (jump_a + jump_r > 0) {
        s.a_rate = (double)jump_a / (double)(jump_r + jump_a);
        MH_std_update(s, s.a_rate);
    }
    memcpy(pos_out, cur_pos, sizeof(cur_pos));
    *df_out = Escape_Prob_log(s, *F_out, s.work.w_theta_xy(g, cur_pos, nullptr));
    return 0;
}

static void fill_mh_config(const Sim &s, int kind, rb2_mh_config &c);

// The DEFAULT sampler (mh_batch = .false.): the serial chains of a time step, each seeing the electrons emitted by the
// ones before it, as ONE device kernel (rb2_mh_planar_serial) instead of one host/device round trip per jump.
// emit_out[k] = 1: candidate k was emitted (and already counted in the field of the later chains).
static int MH_planar_serial_device(Sim &s, int kind, int M, double *df_out, double *F_out, double *pos_out, int *emit_out)
{
    rb2_mh_config c{};
    fill_mh_config(s, kind, c);
    if (s.check(rb2_mh_planar_serial(&c, s.work.w_theta_arr.data(), M, s.rng.next(), df_out, F_out, pos_out, emit_out, &s.a_rate, &s.MH_std),
                "rb2_mh_planar_serial")) return -2;
    return 0;
}

// mh_device: the same lock-step chains with every jump iteration enqueued on the GPU (rb2_mh_planar);
// kind 1 = field emission (:1284-1458), kind 2 = thermal-field chains (src/mod_field_thermo_emission.F90:198-364)
static int MH_planar_device(Sim &s, int kind, int M, double *df_out, double *F_out, double *pos_out)
{
    rb2_mh_config c{};
    fill_mh_config(s, kind, c);
    if (s.check(rb2_mh_planar(&c, s.work.w_theta_arr.data(), M, s.rng.next(), df_out, F_out, pos_out, &s.a_rate, &s.MH_std), "rb2_mh_planar"))
        return -2;
    return 0;
}

static void fill_mh_config(const Sim &s, int kind, rb2_mh_config &c)
{
    const Globals &g = s.g;
    c.kind = kind;
    c.ndim = (kind == 2) ? 25 : 25 * 8;
    c.ndim_first = (kind == 2) ? 0 : (int)lround(c.ndim * 0.25);
    c.image_charge = g.image_charge ? 1 : 0;
    c.y_num = s.work.y_num; c.x_num = s.work.x_num;
    for (int k = 0; k < 2; ++k) { c.emit_pos[k] = g.emitters_pos[k]; c.emit_dim[k] = g.emitters_dim[k]; }
    c.T_temp = g.T_temp;
    c.init_std = 0.10; c.target_rate = 0.35; c.std_gain = 0.025;
    c.std_min = (kind == 2) ? 0.005 : 0.00005; c.std_max = 0.1250;
}

// Lock-step batch, :1284-1458: one device batch per jump iteration
int Metropolis_Hastings_rectangle_J_batch(Sim &s, int M, int emit, double *df_out, double *F_out, double *pos_out)
{
    (void)emit;
    const Globals &g = s.g;
    if (g.mh_device) return MH_planar_device(s, 1, M, df_out, F_out, pos_out);
    const int ndim = 25 * 8, ndim_first = (int)lround(ndim * 0.25);
    std::vector<int> act(M), ok(M, 0);
    std::vector<double> cur((size_t)3 * M, 0.0), w_pos((size_t)3 * M), w_field((size_t)3 * M), sup_cur(M);
    double std[2] = {g.emitters_dim[0] * 0.10, g.emitters_dim[1] * 0.10};
    int n_act = M, count = 0;
    for (int k = 0; k < M; ++k) act[k] = k;
    while (n_act > 0) {
        for (int k = 0; k < n_act; ++k) {
            w_pos[3 * k] = s.rng.uniform() * g.emitters_dim[0] + g.emitters_pos[0];
            w_pos[3 * k + 1] = s.rng.uniform() * g.emitters_dim[1] + g.emitters_pos[1];
            w_pos[3 * k + 2] = 0.0;
        }
        if (s.Calc_Field_at_Surface(n_act, w_pos.data(), w_field.data())) return -2;
        const int old = n_act;
        n_act = 0;
        for (int k = 0; k < old; ++k) {
            const int mc = act[k];
            if (w_field[3 * k + 2] < 0.0) {
                memcpy(&cur[3 * mc], &w_pos[3 * k], 3 * sizeof(double));
                F_out[mc] = w_field[3 * k + 2];
                sup_cur[mc] = Elec_Supply_log(s, w_field[3 * k + 2], s.work.w_theta_xy(g, &w_pos[3 * k], nullptr));
                ok[mc] = 1;
            } else act[n_act++] = mc;
        }
        count++;
        if (count > 10000 && n_act > 0) {
            fprintf(stderr, " Failed to find spot for emission\n");
            for (int k = 0; k < n_act; ++k) {
                const int mc = act[k];
                F_out[mc] = 1.0; sup_cur[mc] = HUGE_NEG;
                cur[3 * mc] = g.emitters_pos[0]; cur[3 * mc + 1] = g.emitters_pos[1]; cur[3 * mc + 2] = 0.0;
            }
            break;
        }
    }
    for (int i = 1; i <= ndim; ++i) {
        int it_a = 0, it_r = 0;
        if (i > ndim_first) { std[0] = g.emitters_dim[0] * s.MH_std; std[1] = g.emitters_dim[1] * s.MH_std; }
        n_act = 0;
        for (int mc = 0; mc < M; ++mc) {
            if (!ok[mc]) continue;
            act[n_act] = mc;
            s.rng.box_muller(&cur[3 * mc], std, &w_pos[3 * n_act]);
            w_pos[3 * n_act + 2] = 0.0;
            check_limits_metro_rec(g, &w_pos[3 * n_act]);
            n_act++;
        }
        if (n_act == 0) break;
        if (s.Calc_Field_at_Surface(n_act, w_pos.data(), w_field.data())) return -2;
        for (int k = 0; k < n_act; ++k) {
            const int mc = act[k];
            if (w_field[3 * k + 2] >= 0.0) { it_r++; continue; }
            const double sup_new = Elec_Supply_log(s, w_field[3 * k + 2], s.work.w_theta_xy(g, &w_pos[3 * k], nullptr));
            const double alpha = sup_new - sup_cur[mc];
            bool accept = sup_new >= sup_cur[mc];
            if (!accept) accept = log(s.rng.uniform()) <= alpha;
            if (accept) {
                memcpy(&cur[3 * mc], &w_pos[3 * k], 3 * sizeof(double)); sup_cur[mc] = sup_new; F_out[mc] = w_field[3 * k + 2]; it_a++;
            } else it_r++;
        }
        if (i > ndim_first && it_a + it_r > 0) {
            s.a_rate = (double)it_a / (double)(it_a + it_r);
            MH_std_update(s, s.a_rate);
        }
    }
    memcpy(pos_out, cur.data(), (size_t)3 * M * sizeof(double));
    for (int mc = 0; mc < M; ++mc)
        df_out[mc] = ok[mc] ? Escape_Prob_log(s, F_out[mc], s.work.w_theta_xy(g, &cur[3 * mc], nullptr)) : HUGE_NEG;
    return 0;
}

// Do_Field_Emission_Planar_rectangle, :261-395
static int Do_Field_Emission_Planar_rectangle(Sim &s, int step, int emit)
{
    const Globals &g = s.g;
    using clk = std::chrono::steady_clock;
    auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const auto t0 = clk::now();
    if (s.check(rb2_field_window_open(), "rb2_field_window_open")) return -1;  // Particles_To_Device
    QuadResult q;
    if (Cuba_Integrate(s, SUPPLY_FE, emit, &q)) return -1;  // Do_Surface_Integration_FE, :635-657
    const auto t1 = clk::now();
    s.t_em_quad += secs(t0, t1);
    const double N_sup = q.integral;
    const int N_round = (int)lround(N_sup + s.residual);
    s.residual = N_sup - N_round;
    s.slog.N_sup = N_sup; s.slog.neval = q.neval; s.slog.fail = q.fail; s.slog.integral_error = q.error;
    memcpy(s.slog.F_avg, q.F_avg, sizeof(q.F_avg));
    if (s.ud_integrand)
        fprintf(s.ud_integrand, "%3d  %8d  %8d  %4d  %12.4E  %12.4E  %12.4E\n", emit, 1, q.neval, q.fail, q.integral, q.error, 0.0);
    std::vector<double> mh_df, mh_F, mh_pos;
    std::vector<int> mh_emit;
    // mh_batch = .false. (the reference's default): the serial chains run on the device in one kernel unless MH_HOST asks
    // for the host loop (one M = 1 field call per jump)
    const bool serial_dev = !g.mh_batch && !g.mh_host && N_round > 0;
    if ((g.mh_batch || serial_dev) && N_round > 0) {
        mh_df.resize(N_round); mh_F.resize(N_round); mh_pos.resize((size_t)3 * N_round);
        if (serial_dev) {
            mh_emit.resize(N_round);
            if (MH_planar_serial_device(s, 1, N_round, mh_df.data(), mh_F.data(), mh_pos.data(), mh_emit.data()) == -2) return -1;
        } else if (Metropolis_Hastings_rectangle_J_batch(s, N_round, emit, mh_df.data(), mh_F.data(), mh_pos.data()) == -2) return -1;
    }
    const auto t2 = clk::now();
    s.t_em_mh += secs(t1, t2);
    s.n_candidates_total += N_round;
    if (s.check(rb2_field_window_close(), "rb2_field_window_close")) return -1;  // Release_Device_Particles
    int nrElecEmit = 0;
    double df_avg = 0.0;
    std::vector<double> add_pos;  // lock-step path: the insertions of this step, made in one call behind the loop
    std::vector<int> add_sec;
    for (int k = 0; k < N_round; ++k) {
        double D_f, F, par_pos[3];
        if (g.mh_batch || serial_dev) { D_f = mh_df[k]; F = mh_F[k]; memcpy(par_pos, &mh_pos[(size_t)3 * k], sizeof(par_pos)); }
        else if (Metropolis_Hastings_rectangle_J(s, emit, &D_f, &F, par_pos) == -2) return -1;
        if (F >= 0.0) D_f = HUGE_NEG;
        df_avg += exp(D_f);
        // the device's serial loop has made the emission test itself (the later chains had to see the result)
        const bool emitted = serial_dev ? (mh_emit[k] != 0) : (log(s.rng.uniform()) <= D_f);
        if (emitted) {
            par_pos[2] = 1.0 * length_scale;
            const double par_vel[3] = {0.0, 0.0, 0.0};
            int sec = 1;
            (void)s.work.w_theta_xy(g, par_pos, &sec);
            if (g.mh_batch || serial_dev) { add_pos.insert(add_pos.end(), par_pos, par_pos + 3); add_sec.push_back(sec); }
            else if (s.Add_Particle(par_pos, par_vel, species_elec, step, emit, -1, sec)) return -1;
            nrElecEmit++;
        }
    }
    if (!add_sec.empty()) {
        const std::vector<double> zero(add_pos.size(), 0.0);
        if (s.Add_Particles((int)add_sec.size(), add_pos.data(), zero.data(), species_elec, step, emit, -1, add_sec.data())) return -1;
    }
    s.t_em_add += secs(t2, clk::now());
    s.slog.df_avg = (N_sup != 0.0) ? df_avg / N_sup : 0.0;
    s.slog.nrElecEmit += nrElecEmit;
    if (s.ud_field)
        fprintf(s.ud_field, "%8d  %16.8E  %16.8E  %16.8E  %16.8E  %16.8E  %16.8E  %16.8E\n", step, q.F_avg[0], q.F_avg[1], q.F_avg[2],
                N_sup, s.slog.df_avg, s.a_rate, s.MH_std);
    return 0;
}

// Do_Field_Emission, :170-205
static int Do_Field_Emission(Sim &s, int step)
{
    s.slog = StepLog{};
    if (s.g.emitters_delay < step) {
        if (s.g.emitters_type == EMIT_CIRCLE || s.g.emitters_type == EMIT_RECTANGLE) {
            if (Do_Field_Emission_Planar_rectangle(s, step, 1)) return -1;
        } else {
            fprintf(stderr, "RUMDEED: WARNING unknown emitter type!!\n");
        }
    }
    return 0;
}

int Init_Field_Emission_v2(Sim &s)
{
    s.residual = 0.0;
    s.a_rate = 1.0;
    s.MH_std = 0.0125;
    if (s.work.w_theta_arr.empty() && s.work.read(s.dir + "/work", s.err)) return -1;  // Read_work_function
    s.ptr.name = "Field emission V2";
    s.ptr.ptr_Do_Emission = Do_Field_Emission;
    s.ptr.ptr_Clean_Up = [](Sim &) { return 0; };
    return 0;
}

// ---- thermal-field emission (mode 9), src/mod_field_thermo_emission.F90 ----------------------------------------
// check_limits_metro_rec, :369-389 (NOT the same rule as the FE module)
static void check_limits_metro_rec_tfe(const Globals &g, double pos[3])
{
    double sx = (pos[0] - g.emitters_pos[0]) / g.emitters_dim[0];
    double sy = (pos[1] - g.emitters_pos[1]) / g.emitters_dim[1];
    if (sx > 1.0 || sx < 0.0) sx = 1.0 - (sx - floor(sx));
    if (sy > 1.0 || sy < 0.0) sy = 1.0 - (sy - floor(sy));
    pos[0] = sx * g.emitters_dim[0] + g.emitters_pos[0];
    pos[1] = sy * g.emitters_dim[1] + g.emitters_pos[1];
}
static const double TINY = 2.2250738585072014e-308;

// Metropolis_Hastings_rectangle_J, :198-364
int Metropolis_Hastings_rectangle_J_thermo(Sim &s, int emit, double pos_out[3])
{
    (void)emit;
    const Globals &g = s.g;
    const int ndim = 25;
    int jump_a = 0, jump_r = 0, count = 0;
    const double std[2] = {g.emitters_dim[0] * s.MH_std, g.emitters_dim[1] * s.MH_std};
    double cur_pos[3], new_pos[3], field[3], cur_w;
    for (;;) {
        cur_pos[0] = s.rng.uniform() * g.emitters_dim[0] + g.emitters_pos[0];
        cur_pos[1] = s.rng.uniform() * g.emitters_dim[1] + g.emitters_pos[1];
        cur_pos[2] = 0.0;
        if (s.Calc_Field_at_Surface(1, cur_pos, field)) return -2;
        cur_w = s.work.w_theta_xy(g, cur_pos, nullptr);
        if (field[2] < 0.0) break;
        if (++count > 10000) {
            pos_out[0] = g.emitters_pos[0]; pos_out[1] = g.emitters_pos[1]; pos_out[2] = 0.0;
            fprintf(stderr, " Failed to find spot for emission\n");
            return -1;
        }
    }
    double df_cur = log(std::max(Get_Kevin_Jgtf_v2(field[2], g.T_temp, cur_w), TINY));
    for (int i = 1; i <= ndim; ++i) {
        s.rng.box_muller(cur_pos, std, new_pos);
        new_pos[2] = 0.0;
        check_limits_metro_rec_tfe(g, new_pos);
        if (s.Calc_Field_at_Surface(1, new_pos, field)) return -2;
        const double new_w = s.work.w_theta_xy(g, new_pos, nullptr);
        if (field[2] > 0.0) { jump_r++; continue; }
        const double df_new = log(std::max(Get_Kevin_Jgtf_v2(field[2], g.T_temp, new_w), TINY));
        const double alpha = df_new - df_cur;
        bool accept = df_new >= df_cur;
        if (!accept) accept = log(s.rng.uniform()) <= alpha;
        if (accept) { memcpy(cur_pos, new_pos, sizeof(cur_pos)); df_cur = df_new; cur_w = new_w; jump_a++; }
        else jump_r++;
    }
    if (jump_a + jump_r > 0) {
        s.a_rate = (double)jump_a / (double)(jump_r + jump_a);
        s.MH_std = s.MH_std * exp(0.025 * (s.a_rate - 0.35));
        if (s.MH_std > 0.1250) s.MH_std = 0.1250; else if (s.MH_std < 0.005) s.MH_std = 0.005;
    }
    memcpy(pos_out, cur_pos, sizeof(cur_pos));
    return 0;
}

// All chains of one time step in lock-step on the device (mh_device); ok_out[k] = 0 for a chain that
// found no favourable spot (the serial routine's "failed" return).
int Metropolis_Hastings_rectangle_J_thermo_batch(Sim &s, int M, double *pos_out, int *ok_out)
{
    if (M < 1) return 0;
    std::vector<double> df(M), F(M);
    const int rc = MH_planar_device(s, 2, M, df.data(), F.data(), pos_out);
    if (rc) return rc;
    for (int k = 0; k < M; ++k) ok_out[k] = F[k] <= 0.0;
    return 0;
}

// Get_MB_Velocity, src/mod_velocity.f90:53-68
static void Get_MB_Velocity(Sim &s, double out[3])
{
    const double mean[2] = {0.0, 0.0};
    const double sd = sqrt(k_b * s.g.T_temp / m_0);
    const double std[2] = {sd, sd};
    double a[2], b[2];
    s.rng.box_muller(mean, std, a);
    s.rng.box_muller(mean, std, b);
    out[0] = a[0]; out[1] = b[0]; out[2] = fabs(b[1]);
}

// Do_Field_Thermo_Emission_Planar_simple, :136-192
static int Do_Field_Thermo_Emission(Sim &s, int step)
{
    const Globals &g = s.g;
    s.slog = StepLog{};
    if (!(g.emitters_delay < step)) return 0;
    QuadResult q;
    if (Cuba_Integrate(s, SUPPLY_GTF, 1, &q)) return -1;  // Do_Surface_Integration_Simple, :448-466
    const double N_sup = q.integral;
    s.slog.N_sup = N_sup; s.slog.neval = q.neval; s.slog.fail = q.fail; s.slog.integral_error = q.error;
    memcpy(s.slog.F_avg, q.F_avg, sizeof(q.F_avg));
    const int N_round = s.rng.poisson(N_sup);
    int nrElecEmit = 0;
    std::vector<double> b_pos;
    std::vector<int> b_ok;
    const bool serial_dev = !g.mh_device && !g.mh_host && N_round > 0;  // the default serial chains, one device kernel
    if (g.mh_device && N_round > 0) {
        b_pos.resize((size_t)3 * N_round); b_ok.resize(N_round);
        if (Metropolis_Hastings_rectangle_J_thermo_batch(s, N_round, b_pos.data(), b_ok.data())) return -1;
    } else if (serial_dev) {
        std::vector<double> df(N_round), F(N_round);
        b_pos.resize((size_t)3 * N_round); b_ok.resize(N_round);
        if (MH_planar_serial_device(s, 2, N_round, df.data(), F.data(), b_pos.data(), b_ok.data())) return -1;
    }
    std::vector<double> add_pos, add_vel;
    std::vector<int> add_sec;
    for (int i = 0; i < N_round; ++i) {
        double par_pos[3], par_vel[3];
        if (g.mh_device || serial_dev) {
            if (!b_ok[i]) continue;
            memcpy(par_pos, &b_pos[(size_t)3 * i], sizeof(par_pos));
        } else {
            const int rc = Metropolis_Hastings_rectangle_J_thermo(s, 1, par_pos);
            if (rc == -2) return -1;
            if (rc < 0) continue;
        }
        par_pos[2] = 1.0 * length_scale;
        Get_MB_Velocity(s, par_vel);
        int sec = 1;
        (void)s.work.w_theta_xy(g, par_pos, &sec);
        if (serial_dev) {  // the device loop has already counted these electrons in the later chains' fields: one insert call
            add_pos.insert(add_pos.end(), par_pos, par_pos + 3); add_vel.insert(add_vel.end(), par_vel, par_vel + 3); add_sec.push_back(sec);
        } else if (s.Add_Particle(par_pos, par_vel, species_elec, step, 1, -1, sec)) return -1;
        nrElecEmit++;
    }
    if (!add_sec.empty() && s.Add_Particles((int)add_sec.size(), add_pos.data(), add_vel.data(), species_elec, step, 1, -1, add_sec.data())) return -1;
    s.slog.nrElecEmit = nrElecEmit;
    if (s.ud_field)
        fprintf(s.ud_field, "%8d  %16.8E  %16.8E  %16.8E  %16.8E  %16.8E  %16.8E  %16.8E\n", step, q.F_avg[0], q.F_avg[1], q.F_avg[2],
                N_sup, 0.0, s.a_rate, s.MH_std);
    return 0;
}

int Init_Field_Thermo_Emission(Sim &s)
{
    s.a_rate = 1.0;
    s.MH_std = 0.0125;
    if (s.work.w_theta_arr.empty() && s.work.read(s.dir + "/work", s.err)) return -1;
    s.ptr.name = "General Field+Thermionic emission";
    s.ptr.ptr_Do_Emission = Do_Field_Thermo_Emission;
    s.ptr.ptr_Clean_Up = [](Sim &) { return 0; };
    return 0;
}

// ---- photo emission (mode 1), src/mod_photo_emission.f90 ---------------------------------------------------------
// The candidate spots of Do_Photo_Emission_Rectangle come from their own generator (rng_photo, seeded from the run's
// seed) through a queue that survives the time step: the (x, y) of the k-th attempt of a run is the k-th pair of that
// stream no matter how many candidates a batch drew ahead, so the serial loop and the speculative one below make the
// same attempts in the same order.
static void photo_candidate(Sim &s, size_t k, double xy[2])
{
    const Globals &g = s.g;
    while ((s.photo_cand.size() - s.photo_head) / 2 <= k) {
        const double u = s.rng_photo.uniform(), v = s.rng_photo.uniform();
        s.photo_cand.push_back(g.emitters_pos[0] + g.emitters_dim[0] * u);
        s.photo_cand.push_back(g.emitters_pos[1] + g.emitters_dim[1] * v);
    }
    xy[0] = s.photo_cand[s.photo_head + 2 * k];
    xy[1] = s.photo_cand[s.photo_head + 2 * k + 1];
}
static void photo_consume(Sim &s, size_t n)
{
    s.photo_head += 2 * n;
    if (s.photo_head > 4096) { s.photo_cand.erase(s.photo_cand.begin(), s.photo_cand.begin() + (long)s.photo_head); s.photo_head = 0; }
}
static double photo_velocity_z(const Sim &s, double p_eV, const double par_pos[3])
{
    if (s.laser.photon_mode != 2) return 0.0;
    return sqrt((2.0 * ((p_eV - s.work.w_theta_xy(s.g, par_pos, nullptr)) * q_0)) / m_0);
}

// Do_Photo_Emission_Rectangle, :603-686, literally: one attempt at a time, one Calc_Field_at per probe, the accepted
// electron inserted at once.  Two host/device round trips per attempt plus one per electron: the reference sequence
// that the speculative loop below must reproduce (option photo_serial; tests).
static int Do_Photo_Emission_Rectangle_serial(Sim &s, int step, int emit, double p_eV, int maxElecEmit)
{
    const Globals &g = s.g;
    const int MAX_EMISSION_TRY = 100;
    int nrTry = 0, nrElecEmit = 0;
    while (nrTry <= MAX_EMISSION_TRY) {
        if (nrElecEmit >= maxElecEmit && maxElecEmit != -1) break;
        if (s.counts.nrElec >= g.max_particles - 1) { fprintf(stderr, "WARNING: Reached maximum number of electrons!!!\n"); break; }
        double xy[2], field[3];
        photo_candidate(s, 0, xy);
        photo_consume(s, 1);
        double par_pos[3] = {xy[0], xy[1], 0.0};
        nrTry++;
        if (s.work.w_theta_xy(g, par_pos, nullptr) <= p_eV) {
            if (s.Calc_Field_at(par_pos, field)) return -1;
            if (field[2] < 0.0) {
                par_pos[2] = 1.0 * length_scale;
                if (s.Calc_Field_at(par_pos, field)) return -1;
                if (field[2] < 0.0) {
                    const double par_vel[3] = {0.0, 0.0, photo_velocity_z(s, p_eV, par_pos)};
                    if (s.Add_Particle(par_pos, par_vel, species_elec, step, emit, -1, 1)) return -1;
                    nrElecEmit++;
                    nrTry = 0;
                }
            }
        }
    }
    s.slog.nrElecEmit += nrElecEmit;
    return 0;
}

// The same loop with device batches.  The accept-and-insert loop is inherently serial (each accepted electron changes
// the field the next attempt sees), but the candidate spots do not depend on the state: up to B upcoming attempts are
// evaluated speculatively as (z = 0, z = 1 nm) probe pairs in ONE rb2_field_batch_delta call against the store plus the
// electrons accepted so far in this step (exact by linearity); the first success is accepted and the rest of the batch,
// which saw a stale field, is evaluated again.  All accepted electrons are inserted behind the loop in one
// rb2_add_particles call (same order, same ids).  One round trip per accepted electron instead of three.
static int Do_Photo_Emission_Rectangle(Sim &s, int step, int emit, double p_eV, int maxElecEmit)
{
    if (s.photo_serial) return Do_Photo_Emission_Rectangle_serial(s, step, emit, p_eV, maxElecEmit);
    const Globals &g = s.g;
    const int MAX_EMISSION_TRY = 100, B = 32;
    int nrTry = 0, nrElecEmit = 0;
    std::vector<double> pts((size_t)6 * B), fld((size_t)6 * B);
    std::vector<double> new_pos, new_vel, new_q;  // accepted in this step, not in the store yet
    std::vector<int> idx(B);
    const int PENDING_MAX = 256;
    auto flush = [&]() -> int {
        const int k = (int)new_q.size();
        if (k < 1) return 0;
        const std::vector<int> sec((size_t)k, 1);
        if (s.Add_Particles(k, new_pos.data(), new_vel.data(), species_elec, step, emit, -1, sec.data())) return -1;
        new_pos.clear(); new_vel.clear(); new_q.clear();
        return 0;
    };
    while (nrTry <= MAX_EMISSION_TRY) {
        if (nrElecEmit >= maxElecEmit && maxElecEmit != -1) break;
        if (s.counts.nrElec + (int)new_q.size() >= g.max_particles - 1) { fprintf(stderr, "WARNING: Reached maximum number of electrons!!!\n"); break; }
        // a batch never looks further than the attempts left before the loop would stop
        const int want = std::min(B, MAX_EMISSION_TRY + 1 - nrTry);
        int m = 0;
        for (int k = 0; k < want; ++k) {
            double xy[2];
            photo_candidate(s, (size_t)k, xy);
            const double pos0[3] = {xy[0], xy[1], 0.0};
            idx[k] = -1;
            if (s.work.w_theta_xy(g, pos0, nullptr) <= p_eV) {
                idx[k] = m;
                double *p = &pts[(size_t)6 * m];
                p[0] = pos0[0]; p[1] = pos0[1]; p[2] = 0.0;
                p[3] = pos0[0]; p[4] = pos0[1]; p[5] = 1.0 * length_scale;
                m++;
            }
        }
        if (m > 0 && s.check(rb2_field_batch_delta(2 * m, pts.data(), (int)new_q.size(), new_pos.data(), new_q.data(), fld.data()),
                             "rb2_field_batch_delta")) return -1;
        int consumed = want;
        for (int k = 0; k < want; ++k) {
            nrTry++;
            if (idx[k] < 0) continue;
            const double *f = &fld[(size_t)6 * idx[k]];
            if (f[2] < 0.0 && f[5] < 0.0) {
                double xy[2];
                photo_candidate(s, (size_t)k, xy);
                const double par_pos[3] = {xy[0], xy[1], 1.0 * length_scale};
                const double par_vel[3] = {0.0, 0.0, photo_velocity_z(s, p_eV, par_pos)};
                new_pos.insert(new_pos.end(), par_pos, par_pos + 3);
                new_vel.insert(new_vel.end(), par_vel, par_vel + 3);
                new_q.push_back(-1.0 * q_0);
                nrElecEmit++;
                nrTry = 0;
                consumed = k + 1;  // the rest of the batch saw a stale field: evaluate it again
                break;
            }
        }
        photo_consume(s, (size_t)consumed);
        // keep the pending list short (it travels to the device with every batch): flushing it into the store in
        // order is the reference's immediate Add_Particle, only later
        if ((int)new_q.size() >= PENDING_MAX && flush()) return -1;
    }
    if (flush()) return -1;
    s.slog.nrElecEmit += nrElecEmit;
    return 0;
}

static int Do_Photo_Emission(Sim &s, int step)
{
    s.slog = StepLog{};
    int maxElecEmit = -1;
    if (s.laser.gauss_mode == 1) {  // Gauss_Emission, :820-841
        const double b = 1.0 / (2.0 * pi * s.laser.gauss_width * s.laser.gauss_width);
        const double b1 = -1.0 * b * (step - s.laser.gauss_center) * (step - s.laser.gauss_center);
        const double ge = (b1 < -500) ? 0.0 : s.laser.gauss_amplitude * exp(b1);
        maxElecEmit = s.rng.poisson(ge);
    }
    if (!(s.g.emitters_delay < step)) return 0;
    if (s.g.emitters_type != EMIT_RECTANGLE) return s.fail("RUMDEED: photo emission: only the rectangle emitter is on the device path");
    double p_eV = s.laser.laser_energy;  // Get_Fixed_Laser_Energy
    if (s.laser.laser_mode == 2) {       // Get_Laser_Energy, :850-866
        const double mean[2] = {s.laser.laser_energy, s.laser.laser_energy};
        const double std[2] = {s.laser.laser_variation, s.laser.laser_variation};
        double a[2], b[2];
        s.rng.box_muller(mean, std, a);
        s.rng.box_muller(mean, std, b);
        p_eV = fabs(b[1]);
    }
    return Do_Photo_Emission_Rectangle(s, step, 1, p_eV, maxElecEmit);
}

int Init_Photo_Emission(Sim &s)
{
    if (s.work.w_theta_arr.empty() && s.work.read(s.dir + "/work", s.err)) return -1;
    if (s.laser.read(s.dir + "/laser", s.err)) {
        if (s.dir.empty()) s.err.clear(); else return -1;  // in-memory set-ups keep the defaults
    }
    s.ptr.name = "Photo emission";
    s.ptr.ptr_Do_Emission = Do_Photo_Emission;
    s.ptr.ptr_Clean_Up = [](Sim &) { return 0; };
    return 0;
}

// ---- hyperboloid tip (mode 3, emitter type 1), src/mod_emission_tip.f90 -------------------------------------------
static const double w_theta_tip = 4.7;  // :45

static double tip_v_y(const Sim &s, double F) { return v_y(s, F, w_theta_tip); }
static double tip_t_y(const Sim &s, double F) { return t_y(s, F, w_theta_tip); }
// Elec_Supply, :1710-1718
static double Elec_Supply(const Sim &s, double A, double F)
{
    const double t = tip_t_y(s, F);
    return A * a_FN * (F * F) * s.g.time_step / (q_0 * w_theta_tip * (t * t));
}
// Escape_Prob_Tip, :1734-1760
static double Escape_Prob_Tip(const Sim &s, double F)
{
    const double sw = sqrt(w_theta_tip);
    return exp(b_FN * (sw * sw * sw) * tip_v_y(s, F) / fabs(F));
}
// Tip_fe_target_log, :1213-1219
static double Tip_fe_target_log(const Sim &s, double eta_f, double xi)
{
    const double t = tip_t_y(s, eta_f);
    const double sup = (s.g.time_step / q_0) * a_FN / ((t * t) * w_theta_tip) * (eta_f * eta_f);
    return log(std::max(sup, TINY)) + 0.5 * log(xi * xi - s.eta_1 * s.eta_1);
}

// The 100 x 100 (xi, phi) midpoint rule of Do_Field_Emission_Tip_OLDCODE, :431-481, as ONE device batch
int Tip_Supply_Grid(Sim &s, int nr_xi, int nr_phi, double *n_s_out, double *F_avg_out)
{
    const double len_phi = 2.0 * pi / nr_phi, len_xi = (s.max_xi - 1.0) / nr_xi;
    const int M = nr_xi * nr_phi;
    // The grid lives on the tip surface: mid points, surface normals and patch areas depend on the geometry only, so
    // they are computed once (1e4 points x ~10 sqrt / log / sin / cos calls were 3-4 ms of every time step).
    if (s.tip_grid_key[0] != nr_xi || s.tip_grid_key[1] != nr_phi || s.tip_grid_geom[0] != s.max_xi || s.tip_grid_geom[1] != s.eta_1 ||
        s.tip_grid_geom[2] != s.a_foci || s.tip_grid_geom[3] != s.shift_z) {
        s.tip_grid_pts.resize((size_t)3 * M); s.tip_grid_nrm.resize((size_t)3 * M); s.tip_grid_area.resize((size_t)M);
        for (int i = 1; i <= nr_xi; ++i)
            for (int j = 1; j <= nr_phi; ++j) {
                const size_t k = (size_t)(i - 1) * nr_phi + (j - 1);
                s.xyz_corr(1.0 + (i - 0.5) * len_xi, s.eta_1, (j - 0.5) * len_phi, &s.tip_grid_pts[3 * k]);
                s.surface_normal(&s.tip_grid_pts[3 * k], &s.tip_grid_nrm[3 * k]);
                s.tip_grid_area[k] = s.Tip_Area(1.0 + (i - 1.0) * len_xi, 1.0 + (i + 0.0) * len_xi, (j - 1.0) * len_phi, (j + 0.0) * len_phi);
            }
        s.tip_grid_key[0] = nr_xi; s.tip_grid_key[1] = nr_phi;
        s.tip_grid_geom[0] = s.max_xi; s.tip_grid_geom[1] = s.eta_1; s.tip_grid_geom[2] = s.a_foci; s.tip_grid_geom[3] = s.shift_z;
        s.tip_grid_on_device = false;
    }
    if (s.g.mh_device) {
        // the grid stays on the device: field, normal projection, supply function and the sum there, two numbers per
        // 256 nodes back (rb2_tip_supply); the host loop below is the MH_DEVICE = .false. path
        if (!s.tip_grid_on_device) {
            if (s.check(rb2_tip_supply_set_grid(M, s.tip_grid_pts.data(), s.tip_grid_nrm.data(), s.tip_grid_area.data()), "rb2_tip_supply_set_grid")) return -1;
            s.tip_grid_on_device = true;
        }
        double n_s = 0.0, F_sum = 0.0;
        if (s.check(rb2_tip_supply(&n_s, &F_sum), "rb2_tip_supply")) return -1;
        *n_s_out = n_s;
        if (F_avg_out) *F_avg_out = F_sum / ((double)nr_phi * nr_xi);
        return 0;
    }
    s.scratch_fld.resize((size_t)3 * M);
    if (s.Calc_Field_at_Batch(M, s.tip_grid_pts.data(), s.scratch_fld.data())) return -1;
    double n_s = 0.0, F_avg = 0.0;
    for (int k = 0; k < M; ++k) {  // same order as the reference's double loop (i outer, j inner)
        const double *u = &s.tip_grid_nrm[(size_t)3 * k], *f = &s.scratch_fld[(size_t)3 * k];
        const double F = u[0] * f[0] + u[1] * f[1] + u[2] * f[2];  // Field_normal
        F_avg += F;
        if (F < 0.0) n_s += Elec_Supply(s, s.tip_grid_area[k], F);
    }
    *n_s_out = n_s;
    if (F_avg_out) *F_avg_out = F_avg / ((double)nr_phi * nr_xi);
    return 0;
}

// Metro_algo_tip_v3, :1241-1390
int Metro_algo_tip_v3(Sim &s, int ndim, double *xi_out, double *phi_out, double *eta_f_out, double *df_cur, double par_pos[3])
{
    const int ndim_first = (int)lround(ndim * 0.25);
    int acc = 0, rej = 0, count = 0;
    if (s.MH_std_tip > 0.125) s.MH_std_tip = 0.125; else if (s.MH_std_tip < 0.0005) s.MH_std_tip = 0.0005;
    double std[2] = {(s.max_xi - 1.0) * 0.10, 2.0 * pi * 0.10};
    double cur_pos[3], new_pos[3], field[3], xi, phi, eta_f;
    for (;;) {
        const double u = s.rng.uniform(), v = s.rng.uniform();
        xi = 1.0 + (s.max_xi - 1.0) * u;
        phi = 2.0 * pi * v;
        s.xyz_corr(xi, s.eta_1, phi, cur_pos);
        if (s.Calc_Field_at(cur_pos, field)) return -2;
        eta_f = s.Field_normal(cur_pos, field);
        if (eta_f < 0.0) break;
        if (++count > 10000) {
            *xi_out = 1.0; *phi_out = 0.0; s.xyz_corr(1.0, s.eta_1, 0.0, par_pos); *eta_f_out = 1.0; *df_cur = 0.0;
            fprintf(stderr, " Failed to find spot for emission on the tip\n");
            return -1;
        }
    }
    double sup_cur = Tip_fe_target_log(s, eta_f, xi);
    const double zero[2] = {0.0, 0.0};
    for (int i = 1; i <= ndim; ++i) {
        if (i > ndim_first) { std[0] = (s.max_xi - 1.0) * s.MH_std_tip; std[1] = 2.0 * pi * s.MH_std_tip; }
        double step2[2];
        s.rng.box_muller(zero, std, step2);
        double new_xi = xi + step2[0];
        double new_phi = fmod(phi + step2[1], 2.0 * pi);
        if (new_phi < 0.0) new_phi += 2.0 * pi;  // Fortran modulo()
        if (new_xi > s.max_xi) new_xi = 2.0 * s.max_xi - new_xi;
        if (new_xi < 1.0) new_xi = 2.0 - new_xi;
        if (new_xi < 1.0 || new_xi > s.max_xi) { if (i > ndim_first) rej++; continue; }
        s.xyz_corr(new_xi, s.eta_1, new_phi, new_pos);
        if (s.Calc_Field_at(new_pos, field)) return -2;
        const double new_eta_f = s.Field_normal(new_pos, field);
        if (new_eta_f >= 0.0) { if (i > ndim_first) rej++; continue; }
        const double sup_new = Tip_fe_target_log(s, new_eta_f, new_xi);
        const double alpha = sup_new - sup_cur;
        bool accept = sup_new >= sup_cur;
        if (!accept) accept = log(s.rng.uniform()) <= alpha;
        if (accept) {
            memcpy(cur_pos, new_pos, sizeof(cur_pos)); xi = new_xi; phi = new_phi; eta_f = new_eta_f; sup_cur = sup_new;
            if (i > ndim_first) acc++;
        } else if (i > ndim_first) rej++;
    }
    if (acc + rej > 0) {
        s.a_rate_tip = (double)acc / (double)(acc + rej);
        s.MH_std_tip = s.MH_std_tip * exp(0.025 * (s.a_rate_tip - 0.35));
        if (s.MH_std_tip > 0.125) s.MH_std_tip = 0.125; else if (s.MH_std_tip < 0.0005) s.MH_std_tip = 0.0005;
    }
    memcpy(par_pos, cur_pos, sizeof(cur_pos));
    *xi_out = xi; *phi_out = phi; *eta_f_out = eta_f;
    *df_cur = Escape_Prob_Tip(s, eta_f);
    return 0;
}

// Lock-step variant of Metro_algo_tip_v3 for mh_batch = .true.: all chains of a time step advance
// together and every jump iteration is ONE device batch -- the scheme of
// Metropolis_Hastings_rectangle_J_batch (src/mod_field_emission_v2.F90:1284-1458) applied to the
// tip's (xi, phi) chains.  Same target, proposal, reflection and wrap rules as the serial chain; the
// shared step MH_std gets one update per jump iteration from the acceptance rate across the batch.
// The reference has no batched tip sampler (its serial one costs n_r x 81 single-point field sums per
// step); like the planar pair the two agree statistically, not run for run.
int Metro_algo_tip_v3_batch(Sim &s, int M, int ndim, double *eta_f_out, double *df_out, double *pos_out)
{
    if (s.g.mh_device) {  // the same lock-step chains with every jump queued on the GPU (rb2_mh_tip)
        if (M < 1) return 0;
        if (s.check(rb2_mh_tip(M, ndim, s.rng.next(), eta_f_out, df_out, pos_out, &s.a_rate_tip, &s.MH_std_tip), "rb2_mh_tip")) return -2;
        return 0;
    }
    const int ndim_first = (int)lround(ndim * 0.25);
    if (s.MH_std_tip > 0.125) s.MH_std_tip = 0.125; else if (s.MH_std_tip < 0.0005) s.MH_std_tip = 0.0005;
    std::vector<int> act(M), ok(M, 0);
    std::vector<double> xi(M, 1.0), phi(M, 0.0), eta_f(M, 1.0), sup_cur(M, 0.0), cur((size_t)3 * M, 0.0);
    std::vector<double> w_xi(M), w_phi(M), w_pos((size_t)3 * M), w_field((size_t)3 * M);
    int n_act = M, count = 0;
    for (int k = 0; k < M; ++k) act[k] = k;
    while (n_act > 0) {
        for (int k = 0; k < n_act; ++k) {
            const double u = s.rng.uniform(), v = s.rng.uniform();
            w_xi[k] = 1.0 + (s.max_xi - 1.0) * u;
            w_phi[k] = 2.0 * pi * v;
            s.xyz_corr(w_xi[k], s.eta_1, w_phi[k], &w_pos[(size_t)3 * k]);
        }
        if (s.Calc_Field_at_Batch(n_act, w_pos.data(), w_field.data())) return -2;
        const int old = n_act;
        n_act = 0;
        for (int k = 0; k < old; ++k) {
            const int mc = act[k];
            const double ef = s.Field_normal(&w_pos[(size_t)3 * k], &w_field[(size_t)3 * k]);
            if (ef < 0.0) {
                xi[mc] = w_xi[k]; phi[mc] = w_phi[k]; eta_f[mc] = ef; ok[mc] = 1;
                memcpy(&cur[(size_t)3 * mc], &w_pos[(size_t)3 * k], 3 * sizeof(double));
                sup_cur[mc] = Tip_fe_target_log(s, ef, w_xi[k]);
            } else act[n_act++] = mc;
        }
        if (++count > 10000 && n_act > 0) {
            fprintf(stderr, " Failed to find spot for emission on the tip\n");
            for (int k = 0; k < n_act; ++k) { const int mc = act[k]; s.xyz_corr(1.0, s.eta_1, 0.0, &cur[(size_t)3 * mc]); eta_f[mc] = 1.0; }
            break;
        }
    }
    double std[2] = {(s.max_xi - 1.0) * 0.10, 2.0 * pi * 0.10};
    const double zero[2] = {0.0, 0.0};
    for (int i = 1; i <= ndim; ++i) {
        int it_a = 0, it_r = 0;
        if (i > ndim_first) { std[0] = (s.max_xi - 1.0) * s.MH_std_tip; std[1] = 2.0 * pi * s.MH_std_tip; }
        n_act = 0;
        for (int mc = 0; mc < M; ++mc) {
            if (!ok[mc]) continue;
            double step2[2];
            s.rng.box_muller(zero, std, step2);
            double new_xi = xi[mc] + step2[0];
            double new_phi = fmod(phi[mc] + step2[1], 2.0 * pi);
            if (new_phi < 0.0) new_phi += 2.0 * pi;
            if (new_xi > s.max_xi) new_xi = 2.0 * s.max_xi - new_xi;
            if (new_xi < 1.0) new_xi = 2.0 - new_xi;
            if (new_xi < 1.0 || new_xi > s.max_xi) { it_r++; continue; }
            act[n_act] = mc; w_xi[n_act] = new_xi; w_phi[n_act] = new_phi;
            s.xyz_corr(new_xi, s.eta_1, new_phi, &w_pos[(size_t)3 * n_act]);
            n_act++;
        }
        if (n_act > 0 && s.Calc_Field_at_Batch(n_act, w_pos.data(), w_field.data())) return -2;
        for (int k = 0; k < n_act; ++k) {
            const int mc = act[k];
            const double ef = s.Field_normal(&w_pos[(size_t)3 * k], &w_field[(size_t)3 * k]);
            if (ef >= 0.0) { it_r++; continue; }
            const double sup_new = Tip_fe_target_log(s, ef, w_xi[k]);
            bool accept = sup_new >= sup_cur[mc];
            if (!accept) accept = log(s.rng.uniform()) <= sup_new - sup_cur[mc];
            if (accept) {
                xi[mc] = w_xi[k]; phi[mc] = w_phi[k]; eta_f[mc] = ef; sup_cur[mc] = sup_new;
                memcpy(&cur[(size_t)3 * mc], &w_pos[(size_t)3 * k], 3 * sizeof(double));
                it_a++;
            } else it_r++;
        }
        if (i > ndim_first && it_a + it_r > 0) {
            s.a_rate_tip = (double)it_a / (double)(it_a + it_r);
            s.MH_std_tip = s.MH_std_tip * exp(0.025 * (s.a_rate_tip - 0.35));
            if (s.MH_std_tip > 0.125) s.MH_std_tip = 0.125; else if (s.MH_std_tip < 0.0005) s.MH_std_tip = 0.0005;
        }
    }
    memcpy(pos_out, cur.data(), (size_t)3 * M * sizeof(double));
    for (int mc = 0; mc < M; ++mc) {
        eta_f_out[mc] = eta_f[mc];
        df_out[mc] = ok[mc] ? Escape_Prob_Tip(s, eta_f[mc]) : 0.0;
    }
    return 0;
}

// Do_Field_Emission_Tip_OLDCODE, :417-534
static int Do_Emission_Tip(Sim &s, int step)
{
    s.slog = StepLog{};
    if (s.g.emitters_type != 1) return s.fail("RUMDEED: tip emitter type != 1 (field emission) is not on the device path");
    double n_s = 0.0, F_avg = 0.0;
    using clk = std::chrono::steady_clock;
    auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const auto t0 = clk::now();
    if (Tip_Supply_Grid(s, 100, 100, &n_s, &F_avg)) return -1;
    const auto t1 = clk::now();
    s.t_em_quad += secs(t0, t1);
    s.slog.N_sup = n_s;
    s.slog.F_avg[2] = F_avg;
    const int n_r = (int)lround(n_s);
    if (n_r < 0) return s.fail("n_r < 0");
    std::vector<double> rnd(n_r);
    for (int k = 0; k < n_r; ++k) rnd[k] = s.rng.uniform();
    int nrElecEmit = 0;
    std::vector<double> b_F, b_D, b_pos;
    if (s.g.mh_batch && n_r > 0) {
        b_F.resize(n_r); b_D.resize(n_r); b_pos.resize((size_t)3 * n_r);
        if (Metro_algo_tip_v3_batch(s, n_r, 80, b_F.data(), b_D.data(), b_pos.data()) == -2) return -1;
    }
    const auto t2 = clk::now();
    s.t_em_mh += secs(t1, t2);
    s.n_candidates_total += n_r;
    for (int k = 0; k < n_r; ++k) {
        double xi, phi, F, D_f, par_pos[3];
        if (s.g.mh_batch) { F = b_F[k]; D_f = b_D[k]; memcpy(par_pos, &b_pos[(size_t)3 * k], sizeof(par_pos)); }
        else if (Metro_algo_tip_v3(s, 80, &xi, &phi, &F, &D_f, par_pos) == -2) return -1;
        if (F < 0.0 && rnd[k] <= D_f) {
            double nrm[3];
            s.surface_normal(par_pos, nrm);
            for (int c = 0; c < 3; ++c) par_pos[c] += nrm[c] * length_scale;
            const double par_vel[3] = {0.0, 0.0, 0.0};
            if (s.Add_Particle(par_pos, par_vel, species_elec, step, 1, -1, 1)) return -1;
            nrElecEmit++;
        }
    }
    s.slog.nrElecEmit = nrElecEmit;
    s.t_em_add += secs(t2, clk::now());
    return 0;
}

// Init_Emission_Tip, :81-127
int Init_Emission_Tip(Sim &s)
{
    Globals &g = s.g;
    const double eta_2 = 0.0;
    s.d_tip = g.emitters_dim[0];
    s.R_base = g.emitters_dim[1];
    s.h_tip = g.emitters_dim[2];
    g.d = s.d_tip + s.h_tip;
    s.max_xi = s.h_tip / s.d_tip + 1.0;
    s.a_foci = sqrt(s.d_tip * s.d_tip * s.R_base * s.R_base / (s.h_tip * s.h_tip + 2 * s.d_tip * s.h_tip) + s.d_tip * s.d_tip);
    s.eta_1 = -1.0 * s.d_tip / s.a_foci;
    s.theta_tip = acos(s.d_tip / s.a_foci);
    s.r_tip = s.a_foci * sin(s.theta_tip) * tan(s.theta_tip);
    s.shift_z = fabs(s.a_foci * s.eta_1 * s.max_xi);
    const double lg = log((1.0 + s.eta_1) / (1.0 - s.eta_1) * (1.0 - eta_2) / (1.0 + eta_2));
    s.pre_fac_E_tip_unit_voltage = 2.0 * 1.0 / (s.a_foci * lg);
    s.pre_fac_E_tip = 2.0 * g.V_s / (s.a_foci * lg);
    s.MH_std_tip = 1.0;
    s.a_rate_tip = 0.5;
    s.ptr.name = "emission from a tip";
    s.ptr.ptr_Do_Emission = Do_Emission_Tip;
    s.ptr.ptr_Clean_Up = [](Sim &) { return 0; };
    return 0;
}

}  // namespace rh

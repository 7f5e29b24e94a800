/*
 * rumdeed_oracle_emission.c -- CPU restatement of the emission samplers that drive the
 * surface-field evaluation.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference draws from the compiler's RANDOM_NUMBER (unpinned); this file
 * uses xoshiro256++ instead, so parity with it is statistical by construction.  The supply
 * integral is done by Cuba in the reference (absent); the oracle's stand-in is a
 * fixed midpoint grid (orc_supply_grid), which the product's adaptive quadrature must match
 * within the reference's own tolerance contract (epsabs 0.5 / epsrel 1e-3).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "rumdeed_oracle.h"
#include "rumdeed_oracle_emission.h"

static const double E_PI = 3.141592653589793238462643383279502884197169399375105820974944592307816406286;
static const double E_LEN = 1.0e-9;
static const double E_Q0 = 1.602176634e-19;
static const double E_M0 = 9.1093837015e-31;
static const double E_KB = 1.380649e-23;

/* ---- RNG ----------------------------------------------------------------------------- */
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
void orc_rng_seed(orc_rng *r, uint64_t seed)
{
    int i;
    for (i = 0; i < 4; ++i) { /* splitmix64 */
        uint64_t z = (seed += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        r->s[i] = z ^ (z >> 31);
    }
}
static inline uint64_t rng_next(orc_rng *r)
{
    uint64_t *s = r->s;
    const uint64_t result = rotl(s[0] + s[3], 23) + s[0];
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
double orc_rng_uniform(orc_rng *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }

/* box_muller (Marsaglia polar form), src/mod_global.F90:578-595 */
void orc_box_muller(orc_rng *r, const double mean[2], const double std[2], double out[2])
{
    double x0, x1, w, f;
    do {
        x0 = 2.0 * orc_rng_uniform(r) - 1.0;
        x1 = 2.0 * orc_rng_uniform(r) - 1.0;
        w = x0 * x0 + x1 * x1;
    } while (!((w < 1.0) && (w > 0.0)));
    f = sqrt((-2.0 * log(w)) / w);
    out[0] = x0 * f * std[0] + mean[0];
    out[1] = x1 * f * std[1] + mean[1];
}

/* Rand_Poisson, src/mod_global.F90:600-643 */
int orc_rand_poisson(orc_rng *r, double lambda)
{
    const double Poisson_Step = 500.0;
    double lambda_left = lambda, p = 1.0;
    int k = 0;
    do {
        k = k + 1;
        p = p * orc_rng_uniform(r);
        while ((p < 1.0) && (lambda_left > 0.0)) {
            if (lambda_left > Poisson_Step) { p = p * exp(Poisson_Step); lambda_left = lambda_left - Poisson_Step; }
            else { p = p * exp(lambda_left); lambda_left = 0.0; }
        }
    } while (p >= 1.0);
    return k - 1;
}

/* Get_MB_Velocity, src/mod_velocity.f90:53-68 */
void orc_get_mb_velocity(orc_rng *r, double T_temp, double out[3])
{
    double mean[2] = {0.0, 0.0}, std[2], a[2], b[2];
    std[0] = std[1] = sqrt(E_KB * T_temp / E_M0);
    orc_box_muller(r, mean, std, a);
    orc_box_muller(r, mean, std, b);
    out[0] = a[0]; out[1] = b[0]; out[2] = fabs(b[1]);
}

/* ---- work function: w_theta_checkerboard, src/mod_work_function.F90:389-487 ------------ */
double orc_w_theta_xy(const orc_emission *E, const double pos[3], int *sec)
{
    const double x_len = 1.0 / E->x_num, y_len = 1.0 / E->y_num;
    double x = (pos[0] - E->emit_pos[0]) / E->emit_dim[0];
    double y = (pos[1] - E->emit_pos[1]) / E->emit_dim[1];
    int x_i = (int)floor(x / x_len) + 1, y_i = (int)floor(y / y_len) + 1;
    if (x_i > E->x_num) x_i = E->x_num; else if (x_i < 1) x_i = 1;
    if (y_i > E->y_num) y_i = E->y_num; else if (y_i < 1) y_i = 1;
    if (sec) *sec = E->x_num * (y_i - 1) + x_i;
    y_i = E->y_num - y_i + 1; /* the file rows run top to bottom */
    return E->w_theta_arr[(y_i - 1) * E->x_num + (x_i - 1)];
}

/* ---- Jensen's general thermal-field current density, src/mod_kevin_rjgtf_v2.f90:32-175 ---- */
static double Nns(double n, double s)
{
    double x, y, z, sn, sd, sng, sfn, v;
    if (n == 1.0) return (s + 1.0) * exp(-s);
    x = n * n; y = 1.0 / x; z = (n - 1.0) * s;
    if (fabs(z) > 1.0e-5) sng = (x + 1.0) * (x * exp(-s) - exp(-n * s)) / (x - 1.0);
    else sng = (0.5 * (x + 1.0) * exp(-s) / (n + 1.0)) * ((1.0 - n) * s * s + 2.0 * (1.0 + n) + 2.0 * s);
    sn = -x * (0.10593434 * x + 0.35506593);
    sd = -y * (0.10593434 * y + 0.35506593);
    sfn = x * exp(-s);
    v = sng + sn * exp(-n * s) + x * sd * exp(-s);
    return v > sfn ? v : sfn;
}
double orc_kevin_jgtf_v2(double F, double T, double w_theta)
{
    const double pi = 3.14159265358979324, kb = 1.0 / 11604.50635, hbar = 0.6582119571, c = 299.7924580;
    const double mo = 5.685630103, afs = 1.0 / 137.035999084, Qo = afs * hbar * c / 4.0;
    const double cm = 1.0e7, Amp = 6.241509074e3;
    const double Arld = (mo * (kb * kb) / (2.0 * (pi * pi) * (hbar * hbar * hbar))) * (cm * cm) / Amp;
    const double chem = 7.0;
    double Fo = fabs(F) * 1.0e-9, To = T, Phi = w_theta;
    double yo, phix, ty, vy, Tmin, Tmax, betaT, betau, betap, theto, nft, sft;
    if (Fo < 1.0e-9) return 0.0;
    yo = sqrt(4.0 * Qo * Fo) / Phi;
    phix = Phi - sqrt(4.0 * Qo * Fo);
    ty = 1.0 + (yo * yo) * (1.0 - log(yo)) / 9.0;
    vy = 1.0 - (yo * yo) * (3.0 - log(yo)) / 3.0;
    Tmin = (hbar * Fo / (4.0 * kb * ty)) * sqrt(2.0 / (mo * Phi));
    Tmax = hbar * Fo / (kb * pi * sqrt(mo * Phi * yo));
    betaT = 1.0 / (kb * To);
    betau = (2.0 / (hbar * Fo)) * sqrt(2.0 * mo * Phi) * ty;
    betap = (pi / (hbar * Fo)) * sqrt(mo * Phi * yo);
    theto = (4.0 * sqrt(2.0 * mo * (Phi * Phi * Phi)) / (3.0 * hbar * Fo)) * vy;
    if (To < Tmin) { nft = betaT / betau; sft = theto; }
    else if (To > Tmax) { nft = betaT / betap; sft = betap * phix; }
    else {
        double Ap = 3.0 * (betap + betau) - 6.0 * theto / phix;
        double Bp = -2.0 * (betap + 2.0 * betau) + 6.0 * theto / phix;
        double Cp = betau - betaT;
        double po = (-Bp - sqrt(Bp * Bp - 4.0 * Ap * Cp)) / (2.0 * Ap);
        double Em = chem + po * phix;
        double theta = ((1.0 - po) * (1.0 - po)) * (2.0 * po + 1.0) * theto - phix * po * (1.0 - po) * ((1.0 - po) * betau - po * betap);
        nft = 1.0;
        sft = theta + betaT * (Em - chem);
    }
    return (Arld * Nns(nft, sft) * (To * To)) * 1.0e4;
}

/* ---- helpers -------------------------------------------------------------------------------- */
static void field_at(const orc_emission *E, const double pt[3], double out[3])
{
    const orc_store *s = E->store;
    orc_calc_field_at(E->p, s->nrPart, s->pos, s->charge, s->species, pt, out);
}

/* check_limits_metro_rec, src/mod_field_emission_v2.F90:1466-1516 (reflection at the edges) */
static void check_limits_fe(const orc_emission *E, double pos[3])
{
    const double x_max = E->emit_pos[0] + E->emit_dim[0], x_min = E->emit_pos[0];
    const double y_max = E->emit_pos[1] + E->emit_dim[1], y_min = E->emit_pos[1];
    if (pos[0] > x_max) pos[0] = x_max - (pos[0] - x_max);
    else if (pos[0] < x_min) pos[0] = (x_min - pos[0]) + x_min;
    if (pos[1] > y_max) pos[1] = y_max - (pos[1] - y_max);
    else if (pos[1] < y_min) pos[1] = (y_min - pos[1]) + y_min;
}
/* check_limits_metro_rec, src/mod_field_thermo_emission.F90:369-389 (a different rule!) */
static void check_limits_tfe(const orc_emission *E, double pos[3])
{
    double sx = (pos[0] - E->emit_pos[0]) / E->emit_dim[0];
    double sy = (pos[1] - E->emit_pos[1]) / E->emit_dim[1];
    if ((sx > 1.0) || (sx < 0.0)) sx = 1.0 - (sx - floor(sx));
    if ((sy > 1.0) || (sy < 0.0)) sy = 1.0 - (sy - floor(sy));
    pos[0] = sx * E->emit_dim[0] + E->emit_pos[0];
    pos[1] = sy * E->emit_dim[1] + E->emit_pos[1];
}

/* MH_std_update, src/mod_field_emission_v2.F90:603-612 with constants :70-76 */
static void mh_std_update(orc_emission *E, double rate)
{
    E->MH_std = E->MH_std * exp(0.025 * (rate - 0.35));
    if (E->MH_std > 0.1250) E->MH_std = 0.1250;
    else if (E->MH_std < 0.00005) E->MH_std = 0.00005;
}

/* integrand_cuba_fe_v for one point (src/mod_field_emission_v2.F90:668-745) and
 * integrand_cuba_simple (src/mod_field_thermo_emission.F90:394-446).  xx in the unit square. */
double orc_supply_integrand(const orc_emission *E, int kind, const double xx[2], double field_out[3])
{
    double pos[3], f[3], A = E->emit_dim[0] * E->emit_dim[1], ff = 0.0;
    pos[0] = E->emit_pos[0] + xx[0] * E->emit_dim[0];
    pos[1] = E->emit_pos[1] + xx[1] * E->emit_dim[1];
    pos[2] = 0.0;
    field_at(E, pos, f);
    if (field_out) { field_out[0] = f[0]; field_out[1] = f[1]; field_out[2] = f[2]; }
    if (f[2] < 0.0) {
        double w = orc_w_theta_xy(E, pos, NULL);
        if (kind == ORC_SUPPLY_FE) ff = orc_fn_elec_supply_v2(E->p, f[2], w);
        else ff = orc_kevin_jgtf_v2(f[2], E->T_temp, w) * (E->p->time_step / E_Q0);
    }
    return A * ff;
}

/* Stand-in for Cuba: n x n midpoint rule over the unit square. */
double orc_supply_grid(const orc_emission *E, int kind, int n, double F_avg[3])
{
    double sum = 0.0, fa[3] = {0, 0, 0};
    int i, j;
#pragma omp parallel for collapse(2) reduction(+ : sum) reduction(+ : fa[:3]) schedule(dynamic, 8)
    for (i = 0; i < n; ++i)
        for (j = 0; j < n; ++j) {
            double xx[2], f[3];
            xx[0] = (i + 0.5) / n; xx[1] = (j + 0.5) / n;
            sum += orc_supply_integrand(E, kind, xx, f);
            fa[0] += f[0]; fa[1] += f[1]; fa[2] += f[2];
        }
    if (F_avg) { F_avg[0] = fa[0] / ((double)n * n); F_avg[1] = fa[1] / ((double)n * n); F_avg[2] = fa[2] / ((double)n * n); }
    return sum / ((double)n * n);
}

/* ---- planar FE: serial chain, src/mod_field_emission_v2.F90:1122-1265 -------------------- */
int orc_mh_rectangle_J(orc_emission *E, orc_rng *r, double *df_out, double *F_out, double pos_out[3])
{
    const int ndim = 25 * 8, ndim_first = (int)lround(ndim * 0.25);
    int jump_a = 0, jump_r = 0, count = 0, i;
    double std[2], cur_pos[3], new_pos[3], field[3], sup_cur, sup_new, alpha;
    std[0] = E->emit_dim[0] * 0.10; std[1] = E->emit_dim[1] * 0.10;
    for (;;) {
        cur_pos[0] = orc_rng_uniform(r) * E->emit_dim[0] + E->emit_pos[0];
        cur_pos[1] = orc_rng_uniform(r) * E->emit_dim[1] + E->emit_pos[1];
        cur_pos[2] = 0.0;
        field_at(E, cur_pos, field);
        if (field[2] < 0.0) break;
        if (++count > 10000) {
            *F_out = 1.0; *df_out = -1.7976931348623157e308;
            pos_out[0] = E->emit_pos[0]; pos_out[1] = E->emit_pos[1]; pos_out[2] = 0.0;
            return -1;
        }
    }
    *F_out = field[2];
    sup_cur = orc_fn_elec_supply_log(E->p, field[2], orc_w_theta_xy(E, cur_pos, NULL));
    for (i = 1; i <= ndim; ++i) {
        if (i > ndim_first) { std[0] = E->emit_dim[0] * E->MH_std; std[1] = E->emit_dim[1] * E->MH_std; }
        orc_box_muller(r, cur_pos, std, new_pos);
        new_pos[2] = 0.0;
        check_limits_fe(E, new_pos);
        field_at(E, new_pos, field);
        if (field[2] >= 0.0) { if (i > ndim_first) jump_r++; continue; }
        sup_new = orc_fn_elec_supply_log(E->p, field[2], orc_w_theta_xy(E, new_pos, NULL));
        alpha = sup_new - sup_cur;
        if (sup_new >= sup_cur) {
            memcpy(cur_pos, new_pos, sizeof(cur_pos)); sup_cur = sup_new; *F_out = field[2];
            if (i > ndim_first) jump_a++;
        } else {
            double rnd = orc_rng_uniform(r);
            if (log(rnd) <= alpha) {
                memcpy(cur_pos, new_pos, sizeof(cur_pos)); sup_cur = sup_new; *F_out = field[2];
                if (i > ndim_first) jump_a++;
            } else if (i > ndim_first) jump_r++;
        }
    }
    if (jump_a + jump_r > 0) {
        E->a_rate = (double)jump_a / (double)(jump_r + jump_a);
        mh_std_update(E, E->a_rate);
    }
    memcpy(pos_out, cur_pos, sizeof(cur_pos));
    *df_out = orc_fn_escape_prob_log(E->p, *F_out, orc_w_theta_xy(E, cur_pos, NULL));
    return 0;
}

/* ---- planar FE: lock-step batch, src/mod_field_emission_v2.F90:1284-1458 ------------------ */
void orc_mh_rectangle_J_batch(orc_emission *E, orc_rng *r, int M, double *df_out, double *F_out, double *pos_out)
{
    const int ndim = 25 * 8, ndim_first = (int)lround(ndim * 0.25);
    const orc_store *s = E->store;
    int *act = (int *)malloc(sizeof(int) * (size_t)M), *ok = (int *)calloc((size_t)M, sizeof(int));
    double *cur = (double *)calloc(3 * (size_t)M, sizeof(double)), *w_pos = (double *)malloc(sizeof(double) * 3 * (size_t)M);
    double *w_field = (double *)malloc(sizeof(double) * 3 * (size_t)M), *sup_cur = (double *)malloc(sizeof(double) * (size_t)M);
    double std[2];
    int n_act = M, count = 0, i, k, mc;
    std[0] = E->emit_dim[0] * 0.10; std[1] = E->emit_dim[1] * 0.10;
    for (k = 0; k < M; ++k) act[k] = k;
    while (n_act > 0) {
        int old;
        for (k = 0; k < n_act; ++k) {
            w_pos[3 * k] = orc_rng_uniform(r) * E->emit_dim[0] + E->emit_pos[0];
            w_pos[3 * k + 1] = orc_rng_uniform(r) * E->emit_dim[1] + E->emit_pos[1];
            w_pos[3 * k + 2] = 0.0;
        }
        orc_calc_field_at_batch(E->p, s->nrPart, s->pos, s->charge, s->species, n_act, w_pos, w_field);
        old = n_act; n_act = 0;
        for (k = 0; k < old; ++k) {
            mc = act[k];
            if (w_field[3 * k + 2] < 0.0) {
                memcpy(&cur[3 * mc], &w_pos[3 * k], 3 * sizeof(double));
                F_out[mc] = w_field[3 * k + 2];
                sup_cur[mc] = orc_fn_elec_supply_log(E->p, w_field[3 * k + 2], orc_w_theta_xy(E, &w_pos[3 * k], NULL));
                ok[mc] = 1;
            } else act[n_act++] = mc;
        }
        count++;
        if ((count > 10000) && (n_act > 0)) {
            for (k = 0; k < n_act; ++k) {
                mc = act[k];
                F_out[mc] = 1.0; sup_cur[mc] = -1.7976931348623157e308;
                cur[3 * mc] = E->emit_pos[0]; cur[3 * mc + 1] = E->emit_pos[1]; cur[3 * mc + 2] = 0.0;
            }
            break;
        }
    }
    for (i = 1; i <= ndim; ++i) {
        int it_a = 0, it_r = 0;
        if (i > ndim_first) { std[0] = E->emit_dim[0] * E->MH_std; std[1] = E->emit_dim[1] * E->MH_std; }
        n_act = 0;
        for (mc = 0; mc < M; ++mc) {
            if (!ok[mc]) continue;
            act[n_act] = mc;
            orc_box_muller(r, &cur[3 * mc], std, &w_pos[3 * n_act]);
            w_pos[3 * n_act + 2] = 0.0;
            check_limits_fe(E, &w_pos[3 * n_act]);
            n_act++;
        }
        if (n_act == 0) break;
        orc_calc_field_at_batch(E->p, s->nrPart, s->pos, s->charge, s->species, n_act, w_pos, w_field);
        for (k = 0; k < n_act; ++k) {
            double sup_new, alpha;
            mc = act[k];
            if (w_field[3 * k + 2] >= 0.0) { it_r++; continue; }
            sup_new = orc_fn_elec_supply_log(E->p, w_field[3 * k + 2], orc_w_theta_xy(E, &w_pos[3 * k], NULL));
            alpha = sup_new - sup_cur[mc];
            if (sup_new >= sup_cur[mc]) {
                memcpy(&cur[3 * mc], &w_pos[3 * k], 3 * sizeof(double)); sup_cur[mc] = sup_new; F_out[mc] = w_field[3 * k + 2]; it_a++;
            } else {
                double rnd = orc_rng_uniform(r);
                if (log(rnd) <= alpha) {
                    memcpy(&cur[3 * mc], &w_pos[3 * k], 3 * sizeof(double)); sup_cur[mc] = sup_new; F_out[mc] = w_field[3 * k + 2]; it_a++;
                } else it_r++;
            }
        }
        if ((i > ndim_first) && (it_a + it_r > 0)) {
            E->a_rate = (double)it_a / (double)(it_a + it_r);
            mh_std_update(E, E->a_rate);
        }
    }
    memcpy(pos_out, cur, 3 * (size_t)M * sizeof(double));
    for (mc = 0; mc < M; ++mc) {
        if (ok[mc]) df_out[mc] = orc_fn_escape_prob_log(E->p, F_out[mc], orc_w_theta_xy(E, &cur[3 * mc], NULL));
        else df_out[mc] = -1.7976931348623157e308;
    }
    free(act); free(ok); free(cur); free(w_pos); free(w_field); free(sup_cur);
}

/* Do_Field_Emission_Planar_rectangle, src/mod_field_emission_v2.F90:261-395, with the supply
 * N_sup handed in (the quadrature is a separate concern).  Returns the number emitted. */
int orc_do_field_emission_planar(orc_emission *E, orc_rng *r, int step, double N_sup, int mh_batch, double *df_avg_out)
{
    int N_round = (int)lround(N_sup + E->residual), s, nrElecEmit = 0;
    double df_avg = 0.0, *mh_df = NULL, *mh_F = NULL, *mh_pos = NULL;
    E->residual = N_sup - N_round;
    if (mh_batch && N_round > 0) {
        mh_df = (double *)malloc(sizeof(double) * (size_t)N_round);
        mh_F = (double *)malloc(sizeof(double) * (size_t)N_round);
        mh_pos = (double *)malloc(sizeof(double) * 3 * (size_t)N_round);
        orc_mh_rectangle_J_batch(E, r, N_round, mh_df, mh_F, mh_pos);
    }
    for (s = 0; s < N_round; ++s) {
        double D_f, F, par_pos[3], par_vel[3] = {0, 0, 0}, rnd;
        int sec;
        if (mh_batch) { D_f = mh_df[s]; F = mh_F[s]; memcpy(par_pos, &mh_pos[3 * s], sizeof(par_pos)); }
        else orc_mh_rectangle_J(E, r, &D_f, &F, par_pos);
        if (F >= 0.0) D_f = -1.7976931348623157e308;
        df_avg += exp(D_f);
        rnd = orc_rng_uniform(r);
        if (log(rnd) <= D_f) {
            par_pos[2] = 1.0 * E_LEN;
            (void)orc_w_theta_xy(E, par_pos, &sec);
            orc_add_particle(E->store, E->p, par_pos, par_vel, ORC_SPECIES_ELEC, step, 1, -1, sec);
            nrElecEmit++;
        }
    }
    if (df_avg_out) *df_avg_out = (N_sup != 0.0) ? df_avg / N_sup : 0.0;
    free(mh_df); free(mh_F); free(mh_pos);
    return nrElecEmit;
}

/* ---- thermal-field: src/mod_field_thermo_emission.F90:198-364 and :136-192 ------------------- */
int orc_mh_rectangle_J_thermo(orc_emission *E, orc_rng *r, double pos_out[3])
{
    const int ndim = 25;
    int jump_a = 0, jump_r = 0, count = 0, i;
    double std[2], cur_pos[3], new_pos[3], field[3], df_cur, df_new, cur_w, new_w, alpha, J;
    std[0] = E->emit_dim[0] * E->MH_std; std[1] = E->emit_dim[1] * E->MH_std;
    for (;;) {
        cur_pos[0] = orc_rng_uniform(r) * E->emit_dim[0] + E->emit_pos[0];
        cur_pos[1] = orc_rng_uniform(r) * E->emit_dim[1] + E->emit_pos[1];
        cur_pos[2] = 0.0;
        field_at(E, cur_pos, field);
        cur_w = orc_w_theta_xy(E, cur_pos, NULL);
        if (field[2] < 0.0) break;
        if (++count > 10000) { pos_out[0] = E->emit_pos[0]; pos_out[1] = E->emit_pos[1]; pos_out[2] = 0.0; return -1; }
    }
    J = orc_kevin_jgtf_v2(field[2], E->T_temp, cur_w);
    df_cur = log(J > 2.2250738585072014e-308 ? J : 2.2250738585072014e-308);
    for (i = 1; i <= ndim; ++i) {
        orc_box_muller(r, cur_pos, std, new_pos);
        new_pos[2] = 0.0;
        check_limits_tfe(E, new_pos);
        field_at(E, new_pos, field);
        new_w = orc_w_theta_xy(E, new_pos, NULL);
        if (field[2] > 0.0) { jump_r++; continue; }
        J = orc_kevin_jgtf_v2(field[2], E->T_temp, new_w);
        df_new = log(J > 2.2250738585072014e-308 ? J : 2.2250738585072014e-308);
        alpha = df_new - df_cur;
        if (df_new >= df_cur) { memcpy(cur_pos, new_pos, sizeof(cur_pos)); df_cur = df_new; cur_w = new_w; jump_a++; }
        else {
            double rnd = orc_rng_uniform(r);
            if (log(rnd) <= alpha) { memcpy(cur_pos, new_pos, sizeof(cur_pos)); df_cur = df_new; cur_w = new_w; jump_a++; }
            else jump_r++;
        }
    }
    if ((jump_a + jump_r) > 0) {
        E->a_rate = (double)jump_a / (double)(jump_r + jump_a);
        E->MH_std = E->MH_std * exp(0.025 * (E->a_rate - 0.35));
        if (E->MH_std > 0.1250) E->MH_std = 0.1250; else if (E->MH_std < 0.005) E->MH_std = 0.005;
    }
    memcpy(pos_out, cur_pos, sizeof(cur_pos));
    return 0;
}

int orc_do_field_thermo_emission_planar(orc_emission *E, orc_rng *r, int step, double N_sup)
{
    int N_round = orc_rand_poisson(r, N_sup), i, nrElecEmit = 0;
    for (i = 0; i < N_round; ++i) {
        double par_pos[3], par_vel[3];
        int sec;
        if (orc_mh_rectangle_J_thermo(E, r, par_pos) < 0) continue;
        par_pos[2] = 1.0 * E_LEN;
        orc_get_mb_velocity(r, E->T_temp, par_vel);
        (void)orc_w_theta_xy(E, par_pos, &sec);
        orc_add_particle(E->store, E->p, par_pos, par_vel, ORC_SPECIES_ELEC, step, 1, -1, sec);
        nrElecEmit++;
    }
    return nrElecEmit;
}

/* Get_Laser_Energy, src/mod_photo_emission.f90:850-866: two Box-Muller pairs, |second value of the second pair| */
double orc_get_laser_energy(orc_rng *r, double laser_energy, double laser_variation)
{
    double mean[2], std[2], a[2], b[2];
    mean[0] = mean[1] = laser_energy;
    std[0] = std[1] = laser_variation;
    orc_box_muller(r, mean, std, a);
    orc_box_muller(r, mean, std, b);
    return fabs(b[1]);
}

/* ---- photo emission: src/mod_photo_emission.f90:603-686 -------------------------------------- */
int orc_do_photo_emission_rectangle(orc_emission *E, orc_rng *r, int step, double p_eV, int photon_mode, int max_elec_emit)
{
    const int MAX_EMISSION_TRY = 100;
    int nrTry = 0, nrElecEmit = 0;
    while (nrTry <= MAX_EMISSION_TRY) {
        double par_pos[3], par_vel[3] = {0, 0, 0}, field[3];
        if ((nrElecEmit >= max_elec_emit) && (max_elec_emit != -1)) break;
        if (E->store->nrElec >= E->store->capacity - 1) break;
        par_pos[0] = E->emit_pos[0] + E->emit_dim[0] * orc_rng_uniform(r);
        par_pos[1] = E->emit_pos[1] + E->emit_dim[1] * orc_rng_uniform(r);
        nrTry++;
        par_pos[2] = 0.0;
        if (orc_w_theta_xy(E, par_pos, NULL) <= p_eV) {
            field_at(E, par_pos, field);
            if (field[2] < 0.0) {
                par_pos[2] = 1.0 * E_LEN;
                field_at(E, par_pos, field);
                if (field[2] < 0.0) {
                    if (photon_mode == 2) par_vel[2] = sqrt((2.0 * ((p_eV - orc_w_theta_xy(E, par_pos, NULL)) * E_Q0)) / E_M0);
                    orc_add_particle(E->store, E->p, par_pos, par_vel, ORC_SPECIES_ELEC, step, 1, -1, 1);
                    nrElecEmit++;
                    nrTry = 0;
                }
            }
        }
    }
    return nrElecEmit;
}

/* ---- hyperboloid tip: src/mod_emission_tip.f90:417-534, :1213-1390 ------------------------------- */
static double tip_normal_field(const orc_emission *E, const double pos[3])
{
    double f[3];
    field_at(E, pos, f);
    return orc_field_normal(E->p, pos, f);
}

/* The 100 x 100 (xi, phi) midpoint rule of Do_Field_Emission_Tip_OLDCODE, :431-481 */
double orc_tip_supply_grid(const orc_emission *E, int nr_xi, int nr_phi, double *F_avg_out)
{
    const orc_params *p = E->p;
    const double len_phi = 2.0 * E_PI / nr_phi, len_xi = (p->max_xi - 1.0) / nr_xi, w_theta = 4.7;
    double n_s = 0.0, F_avg = 0.0;
    int i, j;
#pragma omp parallel for collapse(2) reduction(+ : n_s, F_avg) schedule(dynamic, 8)
    for (i = 1; i <= nr_xi; ++i)
        for (j = 1; j <= nr_phi; ++j) {
            double xi_c = 1.0 + (i - 0.5) * len_xi, phi_c = (j - 0.5) * len_phi, pos[3], F, n_add = 0.0;
            orc_xyz_corr(p, xi_c, p->eta_1, phi_c, pos);
            F = tip_normal_field(E, pos);
            F_avg += F;
            if (F < 0.0) {
                double A_f = orc_tip_area(p, 1.0 + (i - 1.0) * len_xi, 1.0 + (i + 0.0) * len_xi, (j - 1.0) * len_phi, (j + 0.0) * len_phi);
                n_add = orc_tip_elec_supply(p, A_f, F, w_theta);
            }
            n_s += n_add;
        }
    if (F_avg_out) *F_avg_out = F_avg / ((double)nr_phi * nr_xi);
    return n_s;
}

static double tip_target_log(const orc_emission *E, double eta_f, double xi)
{
    /* Tip_fe_target_log :1213-1219 with Elec_Supply_tip :1720-1728 */
    const orc_params *p = E->p;
    const double w_theta = 4.7;
    double t = orc_tip_t_y(p, eta_f, w_theta);
    orc_constants k;
    double sup;
    orc_get_constants(&k);
    sup = (p->time_step / E_Q0) * k.a_FN / ((t * t) * w_theta) * (eta_f * eta_f);
    return log(sup > 2.2250738585072014e-308 ? sup : 2.2250738585072014e-308) + 0.5 * log(xi * xi - p->eta_1 * p->eta_1);
}

int orc_metro_algo_tip_v3(orc_emission *E, orc_rng *r, int ndim, double *xi_out, double *phi_out, double *eta_f_out,
                          double *df_cur, double par_pos[3])
{
    const orc_params *p = E->p;
    const double w_theta = 4.7;
    const int ndim_first = (int)lround(ndim * 0.25);
    int acc = 0, rej = 0, count = 0, i;
    double std[2], step2[2], zero[2] = {0.0, 0.0}, cur_pos[3], new_pos[3], xi, phi, eta_f, sup_cur;
    if (E->MH_std_tip > 0.125) E->MH_std_tip = 0.125; else if (E->MH_std_tip < 0.0005) E->MH_std_tip = 0.0005;
    std[0] = (p->max_xi - 1.0) * 0.10; std[1] = 2.0 * E_PI * 0.10;
    for (;;) {
        step2[0] = orc_rng_uniform(r); step2[1] = orc_rng_uniform(r);
        xi = 1.0 + (p->max_xi - 1.0) * step2[0];
        phi = 2.0 * E_PI * step2[1];
        orc_xyz_corr(p, xi, p->eta_1, phi, cur_pos);
        eta_f = tip_normal_field(E, cur_pos);
        if (eta_f < 0.0) break;
        if (++count > 10000) {
            *xi_out = 1.0; *phi_out = 0.0; orc_xyz_corr(p, 1.0, p->eta_1, 0.0, par_pos); *eta_f_out = 1.0; *df_cur = 0.0;
            return -1;
        }
    }
    sup_cur = tip_target_log(E, eta_f, xi);
    for (i = 1; i <= ndim; ++i) {
        double new_xi, new_phi, new_eta_f, sup_new, alpha;
        if (i > ndim_first) { std[0] = (p->max_xi - 1.0) * E->MH_std_tip; std[1] = 2.0 * E_PI * E->MH_std_tip; }
        orc_box_muller(r, zero, std, step2);
        new_xi = xi + step2[0];
        new_phi = fmod(phi + step2[1], 2.0 * E_PI);
        if (new_phi < 0.0) new_phi += 2.0 * E_PI; /* Fortran modulo() */
        if (new_xi > p->max_xi) new_xi = 2.0 * p->max_xi - new_xi;
        if (new_xi < 1.0) new_xi = 2.0 - new_xi;
        if ((new_xi < 1.0) || (new_xi > p->max_xi)) { if (i > ndim_first) rej++; continue; }
        orc_xyz_corr(p, new_xi, p->eta_1, new_phi, new_pos);
        new_eta_f = tip_normal_field(E, new_pos);
        if (new_eta_f >= 0.0) { if (i > ndim_first) rej++; continue; }
        sup_new = tip_target_log(E, new_eta_f, new_xi);
        alpha = sup_new - sup_cur;
        if (sup_new >= sup_cur || log(orc_rng_uniform(r)) <= alpha) {
            memcpy(cur_pos, new_pos, sizeof(cur_pos)); xi = new_xi; phi = new_phi; eta_f = new_eta_f; sup_cur = sup_new;
            if (i > ndim_first) acc++;
        } else if (i > ndim_first) rej++;
    }
    if (acc + rej > 0) {
        E->a_rate = (double)acc / (double)(acc + rej);
        E->MH_std_tip = E->MH_std_tip * exp(0.025 * (E->a_rate - 0.35));
        if (E->MH_std_tip > 0.125) E->MH_std_tip = 0.125; else if (E->MH_std_tip < 0.0005) E->MH_std_tip = 0.0005;
    }
    memcpy(par_pos, cur_pos, sizeof(cur_pos));
    *xi_out = xi; *phi_out = phi; *eta_f_out = eta_f;
    *df_cur = orc_tip_escape_prob(p, eta_f, w_theta);
    return 0;
}

/* Do_Field_Emission_Tip_OLDCODE, :417-534 (n_s handed in or computed by the caller) */
int orc_do_field_emission_tip(orc_emission *E, orc_rng *r, int step, double n_s)
{
    int n_r = (int)lround(n_s), s, nrElecEmit = 0;
    double *rnd;
    if (n_r < 0) return -1;
    rnd = (double *)malloc(sizeof(double) * (size_t)(n_r > 0 ? n_r : 1));
    for (s = 0; s < n_r; ++s) rnd[s] = orc_rng_uniform(r);
    for (s = 0; s < n_r; ++s) {
        double xi, phi, F, D_f, par_pos[3], nrm[3], par_vel[3] = {0, 0, 0};
        orc_metro_algo_tip_v3(E, r, 80, &xi, &phi, &F, &D_f, par_pos);
        if ((F < 0.0) && (rnd[s] <= D_f)) {
            orc_surface_normal(E->p, par_pos, nrm);
            par_pos[0] += nrm[0] * E_LEN; par_pos[1] += nrm[1] * E_LEN; par_pos[2] += nrm[2] * E_LEN;
            orc_add_particle(E->store, E->p, par_pos, par_vel, ORC_SPECIES_ELEC, step, 1, -1, 1);
            nrElecEmit++;
        }
    }
    free(rnd);
    return nrElecEmit;
}

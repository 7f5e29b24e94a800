/*
 * rumdeed_oracle.c -- CPU restatement of RUMDEED's per-timestep hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see rumdeed_oracle.h).  Plain C99 + OpenMP.
 * Each function follows the cited Fortran statement by statement, including
 * the reference's quirks (softening added to r, index-ordered image roles,
 * q_0/(4 pi eps0) inside the tip image term).  Compile with
 * -ffp-contract=off so that the arithmetic is the one written in the source.
 */
#include "rumdeed_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------------------
 * Constants: src/mod_global.F90:26-75 and :333.  epsilon_0 is DERIVED from
 * mu_0 and c exactly as the reference does (not the CODATA literal).
 */
#define ORC_PI 3.141592653589793238462643383279502884197169399375105820974944592307816406286
static const double K_H = 6.62607015e-34;
static const double K_KB = 1.380649e-23;
static const double K_C = 299792458.0;
static const double K_MU0 = 1.25663706212e-6;
static const double K_MU = 1.66053906660e-27;
static const double K_M0 = 9.1093837015e-31;
static const double K_Q0 = 1.602176634e-19;
static const double K_LEN = 1.0e-9;
static const double K_TIME = 1.0e-12;

static inline double k_eps0(void) { return 1.0 / (K_MU0 * (K_C * K_C)); }
static inline double k_div_fac_c(void) { return 1.0 / (4.0 * ORC_PI * k_eps0() * 1.0); }
static inline double k_hbar(void) { return K_H / (2.0 * ORC_PI); }
static inline double k_mN2(void) { return 28.0134 * K_MU; }
static inline double k_mN2p(void) { return k_mN2() - K_M0; }
static inline double k_aFN(void) { return (K_Q0 * K_Q0) / (16.0 * (ORC_PI * ORC_PI) * k_hbar()); }
static inline double k_bFN(void) { return -4.0 / (3.0 * k_hbar()) * sqrt(2.0 * K_M0 * K_Q0); }
static inline double k_lconst(void) { return K_Q0 / (4.0 * ORC_PI * k_eps0()); }

void orc_get_constants(orc_constants *c)
{
    c->pi = ORC_PI; c->h = K_H; c->k_b = K_KB; c->c = K_C; c->mu_0 = K_MU0;
    c->epsilon_0 = k_eps0(); c->m_u = K_MU; c->h_bar = k_hbar(); c->m_0 = K_M0; c->q_0 = K_Q0;
    c->m_N2 = k_mN2(); c->m_N2p = k_mN2p(); c->length_scale = K_LEN; c->time_scale = K_TIME;
    c->div_fac_c = k_div_fac_c();
    c->a_FN = k_aFN(); c->b_FN = k_bFN(); c->l_const = k_lconst();
}

/* Sample_Elec_Position, src/mod_pair.F90:990-1011 (the per-row scan; the file writer is host-side I/O). */
void orc_nearest_elec(int n, const double *pos, const int *species, double *dist_out, int *id_out)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
        double best = 1000.0; /* particles_nearest_dist = 1000.0d0, :988 */
        int id = -1;
        if (species[i] == ORC_SPECIES_ELEC) {
            for (int j = 0; j < n; ++j) {
                if (j == i || species[j] != ORC_SPECIES_ELEC) continue;
                const double dx = pos[3 * i] - pos[3 * j], dy = pos[3 * i + 1] - pos[3 * j + 1], dz = pos[3 * i + 2] - pos[3 * j + 2];
                const double dist = sqrt(dx * dx + dy * dy + dz * dz); /* norm2, :1002 */
                if (dist < best) { best = dist; id = j; }
            }
        }
        dist_out[i] = best;
        id_out[i] = id;
    }
}

/* bench.py's CPU legs set the OpenMP thread count explicitly (torchrun exports OMP_NUM_THREADS=1) */
void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---------------------------------------------------------------------------
 * Parameter set-up.
 */
static void params_common(orc_params *p, double V_s, const double box_dim[3], double time_step, int image_charge)
{
    memset(p, 0, sizeof(*p));
    p->V_s = V_s;
    p->box_dim[0] = box_dim[0]; p->box_dim[1] = box_dim[1]; p->box_dim[2] = box_dim[2];
    p->time_step = time_step;
    p->image_charge = image_charge ? 1 : 0;
    p->planes_N = 0;
}

/* src/main.F90 Init (d = box_dim(3)) and src/mod_verlet.F90:2050-2052 (E_z = -V_d/d). */
void orc_params_planar(orc_params *p, double V_s, double d, const double box_dim[3], double time_step,
                       int image_charge, int N_ic_max)
{
    params_common(p, V_s, box_dim, time_step, image_charge);
    p->geometry = ORC_GEOM_PLANAR;
    p->N_ic_max = N_ic_max;
    p->d = d;
    p->E_z = -1.0 * V_s / d;
}

/* src/mod_emission_tip.f90:105-125 */
void orc_params_tip(orc_params *p, double V_s, double d_tip, double R_base, double h_tip,
                    const double box_dim[3], double time_step, int image_charge)
{
    const double eta_2 = 0.0; /* src/mod_hyperboloid_tip.f90:15 */
    params_common(p, V_s, box_dim, time_step, image_charge);
    p->geometry = ORC_GEOM_TIP;
    p->N_ic_max = 0;
    p->d_tip = d_tip; p->R_base = R_base; p->h_tip = h_tip;
    p->d = d_tip + h_tip;
    p->E_z = -1.0 * V_s / p->d; /* Set_Voltage, src/mod_verlet.F90:2052 (unused by the tip field) */
    p->max_xi = h_tip / d_tip + 1.0;
    p->a_foci = sqrt((d_tip * d_tip) * (R_base * R_base) / (h_tip * h_tip + 2 * d_tip * h_tip) + d_tip * d_tip);
    p->eta_1 = -1.0 * d_tip / p->a_foci;
    p->theta_tip = acos(d_tip / p->a_foci);
    p->r_tip = p->a_foci * sin(p->theta_tip) * tan(p->theta_tip);
    p->shift_z = fabs(p->a_foci * p->eta_1 * p->max_xi);
    {
        double lg = log((1.0 + p->eta_1) / (1.0 - p->eta_1) * (1.0 - eta_2) / (1.0 + eta_2));
        p->pre_fac_E_tip_unit_voltage = 2.0 * 1.0 / (p->a_foci * lg);
        p->pre_fac_E_tip = 2.0 * V_s / (p->a_foci * lg);
    }
}

/* ---------------------------------------------------------------------------
 * Planar image charge series.
 * src/mod_verlet.F90:1924-1979 (Force_Image_charges_v2) ==
 * src/acc_ic_planar_series.inc:20-63.  Returned WITHOUT the charge prefactor.
 */
static inline void ic_series(int Nic, double d_loc, double z_a, double z_b,
                             double diff_x, double diff_y, double dxy2,
                             double *ic_x, double *ic_y, double *ic_z)
{
    const double soft = K_LEN * K_LEN; /* length_scale**2 = 1e-18 m, added to r */
    double z_ic, diff_z, r, inv_r3, x, y, z;
    int n;

    /* n = 0: opposite charge partner below the cathode */
    z_ic = -1.0 * z_b;
    diff_z = z_a - z_ic;
    r = sqrt(dxy2 + diff_z * diff_z) + soft;
    inv_r3 = 1.0 / (r * r * r);
    x = -diff_x * inv_r3;
    y = -diff_y * inv_r3;
    z = -diff_z * inv_r3;

    for (n = 1; n <= Nic; ++n) {
        /* opposite charge, +n */
        z_ic = 2.0 * n * d_loc - z_b;
        diff_z = z_a - z_ic;
        r = sqrt(dxy2 + diff_z * diff_z) + soft;
        inv_r3 = 1.0 / (r * r * r);
        x = x - diff_x * inv_r3; y = y - diff_y * inv_r3; z = z - diff_z * inv_r3;
        /* opposite charge, -n */
        z_ic = -2.0 * n * d_loc - z_b;
        diff_z = z_a - z_ic;
        r = sqrt(dxy2 + diff_z * diff_z) + soft;
        inv_r3 = 1.0 / (r * r * r);
        x = x - diff_x * inv_r3; y = y - diff_y * inv_r3; z = z - diff_z * inv_r3;
        /* same charge, +n */
        z_ic = 2.0 * n * d_loc + z_b;
        diff_z = z_a - z_ic;
        r = sqrt(dxy2 + diff_z * diff_z) + soft;
        inv_r3 = 1.0 / (r * r * r);
        x = x + diff_x * inv_r3; y = y + diff_y * inv_r3; z = z + diff_z * inv_r3;
        /* same charge, -n */
        z_ic = -2.0 * n * d_loc + z_b;
        diff_z = z_a - z_ic;
        r = sqrt(dxy2 + diff_z * diff_z) + soft;
        inv_r3 = 1.0 / (r * r * r);
        x = x + diff_x * inv_r3; y = y + diff_y * inv_r3; z = z + diff_z * inv_r3;
    }
    *ic_x = x; *ic_y = y; *ic_z = z;
}

void orc_force_image_charges_v2(const orc_params *p, const double pos_1[3], const double pos_2[3], double out[3])
{
    if (!p->image_charge) { out[0] = out[1] = out[2] = 0.0; return; }
    {
        double dx = pos_1[0] - pos_2[0];
        double dy = pos_1[1] - pos_2[1];
        /* The function form sums diff**2 over x,y,z (sum(diff**2)); the .inc form reuses dxy2.
         * (dx*dx + dy*dy) + dz*dz is the same association in both. */
        double dxy2 = dx * dx + dy * dy;
        ic_series(p->N_ic_max, p->d, pos_1[2], pos_2[2], dx, dy, dxy2, &out[0], &out[1], &out[2]);
    }
}

/* ---------------------------------------------------------------------------
 * Hyperboloid tip: prolate spheroidal coordinates, vacuum field, sphere image.
 * src/mod_hyperboloid_tip.f90:25-210 ; src/acc_tip_*.inc
 */
double orc_xi_coor(const orc_params *p, double x, double y, double z)
{
    double a = p->a_foci, s = p->shift_z;
    return 1.0 / (2.0 * a) * (sqrt(x * x + y * y + (z + a - s) * (z + a - s)) + sqrt(x * x + y * y + (z - a - s) * (z - a - s)));
}
double orc_eta_coor(const orc_params *p, double x, double y, double z)
{
    double a = p->a_foci, s = p->shift_z;
    return 1.0 / (2.0 * a) * (sqrt(x * x + y * y + (z + a - s) * (z + a - s)) - sqrt(x * x + y * y + (z - a - s) * (z - a - s)));
}
double orc_phi_coor(double x, double y)
{
    if ((fabs(x) < 1.0e-18) && (fabs(y) < 1.0e-18)) return 0.0;
    return atan2(y, x);
}
void orc_xyz_corr(const orc_params *p, double xi, double eta, double phi, double out[3])
{
    double xy = p->a_foci * sqrt((xi * xi - 1.0) * (1.0 - eta * eta));
    out[0] = xy * cos(phi);
    out[1] = xy * sin(phi);
    out[2] = p->a_foci * xi * eta + p->shift_z;
}
/* src/mod_hyperboloid_tip.f90:78-99 */
void orc_surface_normal(const orc_params *p, const double pos[3], double out[3])
{
    double eta_fac = p->eta_1 / sqrt(1 - p->eta_1 * p->eta_1);
    double div_fac = -1.0 / sqrt(pos[0] * pos[0] + pos[1] * pos[1] + (p->a_foci * p->a_foci) * (1 - p->eta_1 * p->eta_1));
    double nx = eta_fac * pos[0] * div_fac, ny = eta_fac * pos[1] * div_fac, nz = 1.0;
    double nrm = sqrt(nx * nx + ny * ny + nz * nz);
    out[0] = nx / nrm; out[1] = ny / nrm; out[2] = nz / nrm;
}
double orc_field_normal(const orc_params *p, const double pos[3], const double field[3])
{
    double u[3];
    orc_surface_normal(p, pos, u);
    return u[0] * field[0] + u[1] * field[1] + u[2] * field[2];
}
/* src/mod_hyperboloid_tip.f90:156-163 */
double orc_tip_area(const orc_params *p, double xi_1, double xi_2, double phi_1, double phi_2)
{
    double e2 = p->eta_1 * p->eta_1;
    double fac_1 = xi_1 * sqrt(xi_1 * xi_1 - e2) - e2 * log(xi_1 + sqrt(xi_1 * xi_1 - e2));
    double fac_2 = xi_2 * sqrt(xi_2 * xi_2 - e2) - e2 * log(xi_2 + sqrt(xi_2 * xi_2 - e2));
    return 0.5 * (p->a_foci * p->a_foci) * sqrt(1.0 - e2) * (phi_2 - phi_1) * (fac_2 - fac_1);
}

/* src/mod_hyperboloid_tip.f90:115-154 == src/acc_tip_field_E.inc:13-34 */
void orc_field_E_hyperboloid(const orc_params *p, const double pos[3], double out[3])
{
    double xi = orc_xi_coor(p, pos[0], pos[1], pos[2]);
    double eta = orc_eta_coor(p, pos[0], pos[1], pos[2]);
    double phi = orc_phi_coor(pos[0], pos[1]);
    double pre = p->pre_fac_E_tip * 1.0 / (xi * xi - eta * eta);
    double fac_xy;
    if (fabs(xi - 1.0) < 1.0e-6) fac_xy = 0.0;
    else fac_xy = eta * sqrt((xi * xi - 1.0) / (1.0 - eta * eta));
    out[0] = -1.0 * pre * fac_xy * cos(phi);
    out[1] = -1.0 * pre * fac_xy * sin(phi);
    out[2] = pre * xi;
}

void orc_field_E_planar(const orc_params *p, const double pos[3], double out[3])
{
    (void)pos;
    out[0] = 0.0; out[1] = 0.0; out[2] = p->E_z; /* src/mod_verlet.F90:1983-1999 */
}

/* src/mod_hyperboloid_tip.f90:168-210 : pos_1 is imaged in the sphere, the
 * field of pos_1 and of its image is evaluated at pos_2.  Carries
 * q_0/(4 pi eps0) itself (the caller multiplies by its own prefactor again). */
void orc_sphere_ic_field(const orc_params *p, const double pos_1[3], const double pos_2[3], double out[3])
{
    if (!p->image_charge) { out[0] = out[1] = out[2] = 0.0; return; }
    {
        double z_0 = p->h_tip - p->r_tip, R = p->r_tip;
        double x_a = pos_1[0], y_a = pos_1[1], z_a = pos_1[2];
        double x = pos_2[0], y = pos_2[1], z = pos_2[2];
        double dis_a = sqrt(x_a * x_a + y_a * y_a + (z_a - z_0) * (z_a - z_0));
        double zz = (z_a - z_0) * (z_a - z_0);
        double z_b = z_0 + (R * R) / (sqrt(1 + (x_a * x_a) / zz + (y_a * y_a) / zz) * dis_a);
        double x_b = (z_b - z_0) * x_a / (z_a - z_0);
        double y_b = (z_b - z_0) * y_a / (z_a - z_0);
        double tmp_dis_a = pow((x - x_a) * (x - x_a) + (y - y_a) * (y - y_a) + (z - z_a) * (z - z_a), 3.0 / 2.0);
        double tmp_dis_b = pow((x - x_b) * (x - x_b) + (y - y_b) * (y - y_b) + (z - z_b) * (z - z_b), 3.0 / 2.0);
        double pre = 1.0 * K_Q0 / (4.0 * ORC_PI * k_eps0());
        out[0] = pre * ((x_a - x) / tmp_dis_a - (R * (x_b - x)) / (dis_a * tmp_dis_b));
        out[1] = pre * ((y_a - y) / tmp_dis_a - (R * (y_b - y)) / (dis_a * tmp_dis_b));
        out[2] = pre * ((z_a - z) / tmp_dis_a - (R * (z_b - z)) / (dis_a * tmp_dis_b));
    }
}

/* ptr_field_E / ptr_Image_Charge_effect / ptr_E_zunit dispatch, src/mod_global.F90:445-506 */
void orc_field_E(const orc_params *p, const double pos[3], double out[3])
{
    if (p->geometry == ORC_GEOM_TIP) orc_field_E_hyperboloid(p, pos, out);
    else orc_field_E_planar(p, pos, out);
}
void orc_image_charge_effect(const orc_params *p, const double pos_1[3], const double pos_2[3], double out[3])
{
    if (p->geometry == ORC_GEOM_TIP) orc_sphere_ic_field(p, pos_1, pos_2, out);
    else orc_force_image_charges_v2(p, pos_1, pos_2, out);
}
/* src/mod_field_emission_v2.F90:158-166 ; src/mod_emission_tip.f90:133-141 */
void orc_E_zunit(const orc_params *p, const double pos[3], double out[3])
{
    if (p->geometry == ORC_GEOM_TIP) {
        orc_field_E_hyperboloid(p, pos, out);
        out[0] = out[0] * p->pre_fac_E_tip_unit_voltage / p->pre_fac_E_tip;
        out[1] = out[1] * p->pre_fac_E_tip_unit_voltage / p->pre_fac_E_tip;
        out[2] = out[2] * p->pre_fac_E_tip_unit_voltage / p->pre_fac_E_tip;
    } else {
        out[0] = 0.0; out[1] = 0.0; out[2] = -1.0 / p->d;
    }
}

/* ---------------------------------------------------------------------------
 * Generic pair loop: src/mod_verlet.F90:625-751.  Serial (the summation order
 * of the reference's serial build).  Adds into acc.
 */
void orc_accel_generic(const orc_params *p, int n, const double *pos, const double *q, const double *m,
                       const int *species, double *acc)
{
    const double soft = K_LEN * K_LEN;
    const double dfc = k_div_fac_c();
    double *inv_mass = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *sum = (double *)calloc((size_t)(3 * (n > 0 ? n : 1)), sizeof(double));
    int i, j, k;
    for (i = 0; i < n; ++i) inv_mass[i] = 1.0 / m[i];
    for (i = 0; i < n; ++i) {
        double pos_1[3], force_E[3], im_1, q_1, qd_1;
        if (species && species[i] == ORC_SPECIES_ATOM) continue;
        pos_1[0] = pos[3 * i]; pos_1[1] = pos[3 * i + 1]; pos_1[2] = pos[3 * i + 2];
        im_1 = inv_mass[i];
        q_1 = q[i];
        qd_1 = q_1 * dfc;
        orc_field_E(p, pos_1, force_E);
        force_E[0] = q_1 * force_E[0]; force_E[1] = q_1 * force_E[1]; force_E[2] = q_1 * force_E[2];
        for (j = i + 1; j < n; ++j) {
            double pos_2[3], diff[3], force_c[3], force_ic[3], force_ic_N[3], im_2, q_2, pre_fac_c, r, inv_r3;
            if (species && species[j] == ORC_SPECIES_ATOM) continue;
            pos_2[0] = pos[3 * j]; pos_2[1] = pos[3 * j + 1]; pos_2[2] = pos[3 * j + 2];
            im_2 = inv_mass[j];
            q_2 = q[j];
            pre_fac_c = qd_1 * q_2;
            diff[0] = pos_1[0] - pos_2[0]; diff[1] = pos_1[1] - pos_2[1]; diff[2] = pos_1[2] - pos_2[2];
            r = sqrt(diff[0] * diff[0] + diff[1] * diff[1] + diff[2] * diff[2]) + soft;
            inv_r3 = 1.0 / (r * r * r);
            for (k = 0; k < 3; ++k) force_c[k] = (pre_fac_c * inv_r3) * diff[k];
            orc_image_charge_effect(p, pos_1, pos_2, force_ic);
            for (k = 0; k < 3; ++k) force_ic[k] = pre_fac_c * force_ic[k];
            force_ic_N[0] = -1.0 * force_ic[0];
            force_ic_N[1] = -1.0 * force_ic[1];
            force_ic_N[2] = +1.0 * force_ic[2];
            for (k = 0; k < 3; ++k) sum[3 * j + k] = sum[3 * j + k] + im_2 * (force_ic_N[k] - force_c[k]);
            for (k = 0; k < 3; ++k) sum[3 * i + k] = sum[3 * i + k] + im_1 * (force_c[k] + force_ic[k]);
        }
        /* force_ic_self = 0 in the reference */
        for (k = 0; k < 3; ++k) sum[3 * i + k] = sum[3 * i + k] + force_E[k] * im_1 + 0.0 * im_1;
    }
    for (i = 0; i < 3 * n; ++i) acc[i] = acc[i] + sum[i];
    free(sum);
    free(inv_mass);
}

/* ---------------------------------------------------------------------------
 * Planar specialised pair loop: src/mod_verlet.F90:763-884.
 * OpenMP schedule(dynamic,1) with a per-thread 3N reduction array, like the
 * reference's REDUCTION(+:accel_sum).  rows i0<=i<i1 step i_stride.
 */
long long orc_accel_planar_rows(const orc_params *p, int n, const double *pos, const double *q, const double *m,
                                const int *species, double *acc, int i0, int i1, int i_stride)
{
    const double soft = K_LEN * K_LEN;
    const double dfc = k_div_fac_c();
    const double Ez_loc = p->E_z, d_loc = p->d;
    const int Nic = p->N_ic_max, do_ic = p->image_charge;
    long long pairs = 0;
    int i;
    double *inv_mass, *accel_sum;
    int nthreads = 1;
    if (n <= 0) return 0;
    if (i_stride < 1) i_stride = 1;
    inv_mass = (double *)malloc(sizeof(double) * (size_t)n);
#pragma omp parallel for
    for (i = 0; i < n; ++i) inv_mass[i] = 1.0 / m[i];
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    accel_sum = (double *)calloc((size_t)nthreads * 3 * (size_t)n, sizeof(double));

#pragma omp parallel reduction(+ : pairs)
    {
        int tid = 0;
        double *my;
        int ii, j;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        my = accel_sum + (size_t)tid * 3 * (size_t)n;
#pragma omp for schedule(dynamic, 1)
        for (ii = i0; ii < i1; ii += i_stride) {
            double x_1, y_1, z_1, q_1, qd_1, im_1, a_x = 0.0, a_y = 0.0, a_z = 0.0;
            if (species && species[ii] == ORC_SPECIES_ATOM) continue;
            x_1 = pos[3 * ii]; y_1 = pos[3 * ii + 1]; z_1 = pos[3 * ii + 2];
            q_1 = q[ii];
            qd_1 = q_1 * dfc;
            im_1 = inv_mass[ii];
            for (j = ii + 1; j < n; ++j) {
                double x_2, y_2, z_2, q_2, im_2, pre_fac_c, diff_x, diff_y, diff_z, dxy2, r, inv_r3;
                double fc_x, fc_y, fc_z, ic_x, ic_y, ic_z;
                if (species && species[j] == ORC_SPECIES_ATOM) continue;
                x_2 = pos[3 * j]; y_2 = pos[3 * j + 1]; z_2 = pos[3 * j + 2];
                q_2 = q[j];
                im_2 = inv_mass[j];
                pre_fac_c = qd_1 * q_2;
                diff_x = x_1 - x_2; diff_y = y_1 - y_2; diff_z = z_1 - z_2;
                dxy2 = diff_x * diff_x + diff_y * diff_y;
                r = sqrt(dxy2 + diff_z * diff_z) + soft;
                inv_r3 = 1.0 / (r * r * r);
                fc_x = (pre_fac_c * inv_r3) * diff_x;
                fc_y = (pre_fac_c * inv_r3) * diff_y;
                fc_z = (pre_fac_c * inv_r3) * diff_z;
                if (do_ic) {
                    ic_series(Nic, d_loc, z_1, z_2, diff_x, diff_y, dxy2, &ic_x, &ic_y, &ic_z);
                } else {
                    ic_x = 0.0; ic_y = 0.0; ic_z = 0.0;
                }
                a_x = a_x + fc_x + pre_fac_c * ic_x;
                a_y = a_y + fc_y + pre_fac_c * ic_y;
                a_z = a_z + fc_z + pre_fac_c * ic_z;
                my[3 * j] = my[3 * j] + im_2 * (-pre_fac_c * ic_x - fc_x);
                my[3 * j + 1] = my[3 * j + 1] + im_2 * (-pre_fac_c * ic_y - fc_y);
                my[3 * j + 2] = my[3 * j + 2] + im_2 * (pre_fac_c * ic_z - fc_z);
                pairs += 1;
            }
            my[3 * ii] = my[3 * ii] + a_x * im_1;
            my[3 * ii + 1] = my[3 * ii + 1] + a_y * im_1;
            my[3 * ii + 2] = my[3 * ii + 2] + (a_z + q_1 * Ez_loc) * im_1;
        }
    }
    /* fold the per-thread sums (thread order = the OpenMP reduction order) */
#pragma omp parallel for
    for (i = 0; i < 3 * n; ++i) {
        double s = 0.0;
        int t;
        for (t = 0; t < nthreads; ++t) s += accel_sum[(size_t)t * 3 * (size_t)n + i];
        acc[i] = acc[i] + s;
    }
    free(accel_sum);
    free(inv_mass);
    return pairs;
}

void orc_accel_planar(const orc_params *p, int n, const double *pos, const double *q, const double *m,
                      const int *species, double *acc)
{
    (void)orc_accel_planar_rows(p, n, pos, q, m, species, acc, 0, n, 1);
}

/* ---------------------------------------------------------------------------
 * Gather formulation (the OpenACC kernels): src/mod_verlet.F90:1217-1429.
 * OVERWRITES acc.  Image roles by index: a = lower index, b = higher index.
 */
static void tip_image_point(double z_0_tip, double Rsp_tip, double x_a, double y_a, double z_a,
                            double *dis_a, double *x_im, double *y_im, double *z_im)
{
    /* src/acc_tip_image_point.inc:14-18 */
    double zz = (z_a - z_0_tip) * (z_a - z_0_tip);
    *dis_a = sqrt(x_a * x_a + y_a * y_a + zz);
    *z_im = z_0_tip + (Rsp_tip * Rsp_tip) / (sqrt(1 + (x_a * x_a) / zz + (y_a * y_a) / zz) * (*dis_a));
    *x_im = (*z_im - z_0_tip) * x_a / (z_a - z_0_tip);
    *y_im = (*z_im - z_0_tip) * y_a / (z_a - z_0_tip);
}
static void tip_ic_force(double Rsp_tip, double x_a, double y_a, double z_a, double x_b, double y_b, double z_b,
                         double dis_a, double x_im, double y_im, double z_im,
                         double *ic_x, double *ic_y, double *ic_z)
{
    /* src/acc_tip_ic_force.inc:17-22 */
    double tmp_dis_a = pow((x_b - x_a) * (x_b - x_a) + (y_b - y_a) * (y_b - y_a) + (z_b - z_a) * (z_b - z_a), 3.0 / 2.0);
    double tmp_dis_b = pow((x_b - x_im) * (x_b - x_im) + (y_b - y_im) * (y_b - y_im) + (z_b - z_im) * (z_b - z_im), 3.0 / 2.0);
    double pre = 1.0 * K_Q0 / (4.0 * ORC_PI * k_eps0());
    *ic_x = pre * ((x_a - x_b) / tmp_dis_a - (Rsp_tip * (x_im - x_b)) / (dis_a * tmp_dis_b));
    *ic_y = pre * ((y_a - y_b) / tmp_dis_a - (Rsp_tip * (y_im - y_b)) / (dis_a * tmp_dis_b));
    *ic_z = pre * ((z_a - z_b) / tmp_dis_a - (Rsp_tip * (z_im - z_b)) / (dis_a * tmp_dis_b));
}

void orc_accel_gather(const orc_params *p, int n, const double *pos, const double *q, const double *m, double *acc)
{
    const double soft = K_LEN * K_LEN;
    const double dfc = k_div_fac_c();
    const double Ez_loc = p->E_z, d_loc = p->d;
    const int Nic = p->N_ic_max, do_ic = p->image_charge;
    const double z_0_tip = p->h_tip - p->r_tip, Rsp_tip = p->r_tip;
    int i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < n; ++i) {
        double x_1 = pos[3 * i], y_1 = pos[3 * i + 1], z_1 = pos[3 * i + 2];
        double q_1 = q[i], qd_1 = q_1 * dfc, a_x = 0.0, a_y = 0.0, a_z = 0.0, im_1;
        int j;
        for (j = 0; j < n; ++j) {
            double x_2, y_2, z_2, q_2, pre_fac_c, diff_x, diff_y, diff_z, dxy2, r, inv_r3, ic_x, ic_y, ic_z;
            if (j == i) continue;
            x_2 = pos[3 * j]; y_2 = pos[3 * j + 1]; z_2 = pos[3 * j + 2];
            q_2 = q[j];
            pre_fac_c = qd_1 * q_2;
            diff_x = x_1 - x_2; diff_y = y_1 - y_2; diff_z = z_1 - z_2;
            if (p->geometry == ORC_GEOM_PLANAR) {
                dxy2 = diff_x * diff_x + diff_y * diff_y;
                r = sqrt(dxy2 + diff_z * diff_z) + soft;
                inv_r3 = 1.0 / (r * r * r);
                a_x = a_x + pre_fac_c * inv_r3 * diff_x;
                a_y = a_y + pre_fac_c * inv_r3 * diff_y;
                a_z = a_z + pre_fac_c * inv_r3 * diff_z;
                if (do_ic) {
                    double z_a, z_b;
                    if (j > i) { z_a = z_1; z_b = z_2; } else { z_a = z_2; z_b = z_1; }
                    ic_series(Nic, d_loc, z_a, z_b, diff_x, diff_y, dxy2, &ic_x, &ic_y, &ic_z);
                    a_x = a_x + pre_fac_c * ic_x;
                    a_y = a_y + pre_fac_c * ic_y;
                    a_z = a_z + pre_fac_c * ic_z;
                }
            } else {
                r = sqrt(diff_x * diff_x + diff_y * diff_y + diff_z * diff_z) + soft;
                inv_r3 = 1.0 / (r * r * r);
                a_x = a_x + pre_fac_c * inv_r3 * diff_x;
                a_y = a_y + pre_fac_c * inv_r3 * diff_y;
                a_z = a_z + pre_fac_c * inv_r3 * diff_z;
                if (do_ic) {
                    double x_a, y_a, z_a, x_b, y_b, z_b, sgn_xy, dis_a, x_im, y_im, z_im;
                    if (j > i) { x_a = x_1; y_a = y_1; z_a = z_1; x_b = x_2; y_b = y_2; z_b = z_2; sgn_xy = 1.0; }
                    else       { x_a = x_2; y_a = y_2; z_a = z_2; x_b = x_1; y_b = y_1; z_b = z_1; sgn_xy = -1.0; }
                    tip_image_point(z_0_tip, Rsp_tip, x_a, y_a, z_a, &dis_a, &x_im, &y_im, &z_im);
                    tip_ic_force(Rsp_tip, x_a, y_a, z_a, x_b, y_b, z_b, dis_a, x_im, y_im, z_im, &ic_x, &ic_y, &ic_z);
                    a_x = a_x + pre_fac_c * sgn_xy * ic_x;
                    a_y = a_y + pre_fac_c * sgn_xy * ic_y;
                    a_z = a_z + pre_fac_c * ic_z;
                }
            }
        }
        im_1 = 1.0 / m[i];
        if (p->geometry == ORC_GEOM_PLANAR) {
            acc[3 * i] = a_x * im_1;
            acc[3 * i + 1] = a_y * im_1;
            acc[3 * i + 2] = (a_z + q_1 * Ez_loc) * im_1;
        } else {
            double pt[3], fE[3];
            pt[0] = x_1; pt[1] = y_1; pt[2] = z_1;
            orc_field_E_hyperboloid(p, pt, fE);
            acc[3 * i] = (a_x + q_1 * fE[0]) * im_1;
            acc[3 * i + 1] = (a_y + q_1 * fE[1]) * im_1;
            acc[3 * i + 2] = (a_z + q_1 * fE[2]) * im_1;
        }
    }
}

/* ---------------------------------------------------------------------------
 * Higher-precision truth: the same formulas evaluated in long double (x87
 * 80-bit, 64-bit mantissa) with Neumaier-compensated accumulation.  Used to
 * bound the rounding error of BOTH summation orders (scatter and gather) and
 * of the CUDA kernels.
 */
typedef struct { long double s, c; } ksum;
static inline void kadd(ksum *k, long double v)
{
    long double t = k->s + v;
    if (fabsl(k->s) >= fabsl(v)) k->c += (k->s - t) + v; else k->c += (v - t) + k->s;
    k->s = t;
}
static inline long double kval(const ksum *k) { return k->s + k->c; }

static inline long double ld_inv_r3(long double s2)
{
    long double r = sqrtl(s2) + (long double)K_LEN * (long double)K_LEN;
    return 1.0L / (r * r * r);
}
static void ic_series_ld(int Nic, long double d_loc, long double z_a, long double z_b,
                         long double dx, long double dy, long double dxy2, ksum *sx, ksum *sy, ksum *sz, long double pre)
{
    int n, t;
    long double dz = z_a + z_b, w = ld_inv_r3(dxy2 + dz * dz);
    kadd(sx, -pre * dx * w); kadd(sy, -pre * dy * w); kadd(sz, -pre * dz * w);
    for (n = 1; n <= Nic; ++n) {
        long double zic[4]; long double sg[4];
        zic[0] = 2.0L * n * d_loc - z_b;  sg[0] = -1.0L;
        zic[1] = -2.0L * n * d_loc - z_b; sg[1] = -1.0L;
        zic[2] = 2.0L * n * d_loc + z_b;  sg[2] = +1.0L;
        zic[3] = -2.0L * n * d_loc + z_b; sg[3] = +1.0L;
        for (t = 0; t < 4; ++t) {
            dz = z_a - zic[t];
            w = ld_inv_r3(dxy2 + dz * dz);
            kadd(sx, sg[t] * pre * dx * w); kadd(sy, sg[t] * pre * dy * w); kadd(sz, sg[t] * pre * dz * w);
        }
    }
}
static void sphere_ic_ld(const orc_params *p, const long double a[3], const long double b[3], long double out[3])
{
    long double z_0 = (long double)p->h_tip - (long double)p->r_tip, R = (long double)p->r_tip;
    long double zz = (a[2] - z_0) * (a[2] - z_0);
    long double dis_a = sqrtl(a[0] * a[0] + a[1] * a[1] + zz);
    long double z_im = z_0 + (R * R) / (sqrtl(1 + (a[0] * a[0]) / zz + (a[1] * a[1]) / zz) * dis_a);
    long double x_im = (z_im - z_0) * a[0] / (a[2] - z_0);
    long double y_im = (z_im - z_0) * a[1] / (a[2] - z_0);
    long double ta = (b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1]) + (b[2] - a[2]) * (b[2] - a[2]);
    long double tb = (b[0] - x_im) * (b[0] - x_im) + (b[1] - y_im) * (b[1] - y_im) + (b[2] - z_im) * (b[2] - z_im);
    long double eps0 = 1.0L / ((long double)K_MU0 * ((long double)K_C * (long double)K_C));
    long double pre = (long double)K_Q0 / (4.0L * 3.141592653589793238462643383279502884L * eps0);
    ta = ta * sqrtl(ta); tb = tb * sqrtl(tb);
    out[0] = pre * ((a[0] - b[0]) / ta - (R * (x_im - b[0])) / (dis_a * tb));
    out[1] = pre * ((a[1] - b[1]) / ta - (R * (y_im - b[1])) / (dis_a * tb));
    out[2] = pre * ((a[2] - b[2]) / ta - (R * (z_im - b[2])) / (dis_a * tb));
}
static void field_E_hyperboloid_ld(const orc_params *p, const long double pos[3], long double out[3])
{
    long double a = p->a_foci, s = p->shift_z;
    long double x = pos[0], y = pos[1], z = pos[2];
    long double r_p = sqrtl(x * x + y * y + (z + a - s) * (z + a - s));
    long double r_m = sqrtl(x * x + y * y + (z - a - s) * (z - a - s));
    long double xi = (r_p + r_m) / (2.0L * a), eta = (r_p - r_m) / (2.0L * a), phi, pre, fac_xy;
    if ((fabsl(x) < 1.0e-18L) && (fabsl(y) < 1.0e-18L)) phi = 0.0L; else phi = atan2l(y, x);
    pre = (long double)p->pre_fac_E_tip / (xi * xi - eta * eta);
    if (fabsl(xi - 1.0L) < 1.0e-6L) fac_xy = 0.0L; else fac_xy = eta * sqrtl((xi * xi - 1.0L) / (1.0L - eta * eta));
    out[0] = -pre * fac_xy * cosl(phi);
    out[1] = -pre * fac_xy * sinl(phi);
    out[2] = pre * xi;
}

/* rows == NULL: the contiguous rows i0 .. i0+nrows-1; else the listed rows (any order) */
static void accel_gather_ld_rows(const orc_params *p, int n, const double *pos, const double *q, const double *m,
                                 int i0, int nrows, const int *rows, double *acc_out)
{
    const long double eps0 = 1.0L / ((long double)K_MU0 * ((long double)K_C * (long double)K_C));
    const long double dfc = 1.0L / (4.0L * 3.141592653589793238462643383279502884L * eps0);
    int r;
#pragma omp parallel for schedule(dynamic, 4)
    for (r = 0; r < nrows; ++r) {
        const int i = rows ? rows[r] : i0 + r;
        long double x_1 = pos[3 * i], y_1 = pos[3 * i + 1], z_1 = pos[3 * i + 2];
        long double q_1 = q[i], qd_1 = q_1 * dfc, im_1;
        ksum sx = {0, 0}, sy = {0, 0}, sz = {0, 0};
        int j;
        for (j = 0; j < n; ++j) {
            long double x_2, y_2, z_2, pre, dx, dy, dz, dxy2, w;
            if (j == i) continue;
            x_2 = pos[3 * j]; y_2 = pos[3 * j + 1]; z_2 = pos[3 * j + 2];
            pre = qd_1 * (long double)q[j];
            dx = x_1 - x_2; dy = y_1 - y_2; dz = z_1 - z_2;
            dxy2 = dx * dx + dy * dy;
            w = ld_inv_r3(dxy2 + dz * dz);
            kadd(&sx, pre * w * dx); kadd(&sy, pre * w * dy); kadd(&sz, pre * w * dz);
            if (!p->image_charge) continue;
            if (p->geometry == ORC_GEOM_PLANAR) {
                long double z_a, z_b;
                if (j > i) { z_a = z_1; z_b = z_2; } else { z_a = z_2; z_b = z_1; }
                ic_series_ld(p->N_ic_max, (long double)p->d, z_a, z_b, dx, dy, dxy2, &sx, &sy, &sz, pre);
            } else {
                long double a[3], b[3], ic[3], sgn;
                if (j > i) { a[0] = x_1; a[1] = y_1; a[2] = z_1; b[0] = x_2; b[1] = y_2; b[2] = z_2; sgn = 1.0L; }
                else       { a[0] = x_2; a[1] = y_2; a[2] = z_2; b[0] = x_1; b[1] = y_1; b[2] = z_1; sgn = -1.0L; }
                sphere_ic_ld(p, a, b, ic);
                kadd(&sx, pre * sgn * ic[0]); kadd(&sy, pre * sgn * ic[1]); kadd(&sz, pre * ic[2]);
            }
        }
        im_1 = 1.0L / (long double)m[i];
        if (p->geometry == ORC_GEOM_PLANAR) {
            kadd(&sz, q_1 * (long double)p->E_z);
        } else {
            long double pt[3], fE[3];
            pt[0] = x_1; pt[1] = y_1; pt[2] = z_1;
            field_E_hyperboloid_ld(p, pt, fE);
            kadd(&sx, q_1 * fE[0]); kadd(&sy, q_1 * fE[1]); kadd(&sz, q_1 * fE[2]);
        }
        acc_out[3 * r] = (double)(kval(&sx) * im_1);
        acc_out[3 * r + 1] = (double)(kval(&sy) * im_1);
        acc_out[3 * r + 2] = (double)(kval(&sz) * im_1);
    }
}
void orc_accel_gather_ld(const orc_params *p, int n, const double *pos, const double *q, const double *m,
                         int i0, int i1, double *acc_out)
{
    accel_gather_ld_rows(p, n, pos, q, m, i0, i1 - i0, NULL, acc_out);
}
void orc_accel_gather_ld_rows(const orc_params *p, int n, const double *pos, const double *q, const double *m,
                              int nrows, const int *rows, double *acc_out)
{
    accel_gather_ld_rows(p, n, pos, q, m, 0, nrows, rows, acc_out);
}

/* ---------------------------------------------------------------------------
 * Field at a point: src/mod_verlet.F90:1466-1529.  Serial j order.
 */
void orc_calc_field_at(const orc_params *p, int n, const double *pos, const double *q, const int *species,
                       const double pt[3], double out[3])
{
    const double soft = K_LEN * K_LEN;
    const double dfc = k_div_fac_c();
    double tot[3];
    int j, k;
    orc_field_E(p, pt, tot);
    for (j = 0; j < n; ++j) {
        double pos_2[3], diff[3], force_c[3], force_ic[3], q_2, pre_fac_c, r, inv_r3;
        if (species && species[j] == ORC_SPECIES_ATOM) continue;
        pos_2[0] = pos[3 * j]; pos_2[1] = pos[3 * j + 1]; pos_2[2] = pos[3 * j + 2];
        q_2 = q[j];
        pre_fac_c = q_2 * dfc;
        for (k = 0; k < 3; ++k) diff[k] = pt[k] - pos_2[k];
        r = sqrt(diff[0] * diff[0] + diff[1] * diff[1] + diff[2] * diff[2]) + soft;
        inv_r3 = 1.0 / (r * r * r);
        for (k = 0; k < 3; ++k) force_c[k] = diff[k] * inv_r3;
        orc_image_charge_effect(p, pt, pos_2, force_ic);
        for (k = 0; k < 3; ++k) tot[k] = tot[k] + pre_fac_c * (force_c[k] + force_ic[k]);
    }
    out[0] = tot[0]; out[1] = tot[1]; out[2] = tot[2];
}

void orc_calc_field_at_batch(const orc_params *p, int n, const double *pos, const double *q, const int *species,
                             int M, const double *pts, double *out)
{
    int k;
#pragma omp parallel for schedule(dynamic, 1)
    for (k = 0; k < M; ++k) orc_calc_field_at(p, n, pos, q, species, pts + 3 * k, out + 3 * k);
}

void orc_calc_field_at_ld(const orc_params *p, int n, const double *pos, const double *q, const int *species,
                          const double pt[3], double out[3])
{
    const long double eps0 = 1.0L / ((long double)K_MU0 * ((long double)K_C * (long double)K_C));
    const long double dfc = 1.0L / (4.0L * 3.141592653589793238462643383279502884L * eps0);
    ksum sx = {0, 0}, sy = {0, 0}, sz = {0, 0};
    long double x_1 = pt[0], y_1 = pt[1], z_1 = pt[2];
    int j;
    if (p->geometry == ORC_GEOM_PLANAR) {
        kadd(&sz, (long double)p->E_z);
    } else {
        long double q3[3], fE[3];
        q3[0] = x_1; q3[1] = y_1; q3[2] = z_1;
        field_E_hyperboloid_ld(p, q3, fE);
        kadd(&sx, fE[0]); kadd(&sy, fE[1]); kadd(&sz, fE[2]);
    }
    for (j = 0; j < n; ++j) {
        long double x_2, y_2, z_2, pre, dx, dy, dz, dxy2, w;
        if (species && species[j] == ORC_SPECIES_ATOM) continue;
        x_2 = pos[3 * j]; y_2 = pos[3 * j + 1]; z_2 = pos[3 * j + 2];
        pre = (long double)q[j] * dfc;
        dx = x_1 - x_2; dy = y_1 - y_2; dz = z_1 - z_2;
        dxy2 = dx * dx + dy * dy;
        w = ld_inv_r3(dxy2 + dz * dz);
        kadd(&sx, pre * w * dx); kadd(&sy, pre * w * dy); kadd(&sz, pre * w * dz);
        if (!p->image_charge) continue;
        if (p->geometry == ORC_GEOM_PLANAR) {
            ic_series_ld(p->N_ic_max, (long double)p->d, z_1, z_2, dx, dy, dxy2, &sx, &sy, &sz, pre);
        } else {
            long double a[3], b[3], ic[3];
            a[0] = x_1; a[1] = y_1; a[2] = z_1; b[0] = x_2; b[1] = y_2; b[2] = z_2;
            sphere_ic_ld(p, a, b, ic);
            kadd(&sx, pre * ic[0]); kadd(&sy, pre * ic[1]); kadd(&sz, pre * ic[2]);
        }
    }
    out[0] = (double)kval(&sx); out[1] = (double)kval(&sy); out[2] = (double)kval(&sz);
}

/* ---------------------------------------------------------------------------
 * Particle store.  src/mod_global.F90:128-172 (arrays), src/mod_pair.F90.
 */
orc_store *orc_store_new(int capacity)
{
    orc_store *s = (orc_store *)calloc(1, sizeof(orc_store));
    size_t c = (size_t)(capacity > 0 ? capacity : 1);
    int i;
    s->capacity = capacity;
    s->pos = (double *)calloc(3 * c, sizeof(double));
    s->prev_pos = (double *)calloc(3 * c, sizeof(double));
    s->vel = (double *)calloc(3 * c, sizeof(double));
    s->acc = (double *)calloc(3 * c, sizeof(double));
    s->acc_prev = (double *)calloc(3 * c, sizeof(double));
    s->acc_prev2 = (double *)calloc(3 * c, sizeof(double));
    s->charge = (double *)calloc(c, sizeof(double));
    s->mass = (double *)calloc(c, sizeof(double));
    s->species = (int *)calloc(c, sizeof(int));
    s->step = (int *)calloc(c, sizeof(int));
    s->emitter = (int *)calloc(c, sizeof(int));
    s->section = (int *)calloc(c, sizeof(int));
    s->life = (int *)calloc(c, sizeof(int));
    s->id = (int *)calloc(c, sizeof(int));
    s->mask = (int *)calloc(c, sizeof(int));
    for (i = 0; i < capacity; ++i) s->mask[i] = 1;
    s->cap_events = 64;
    s->events = (orc_event *)calloc((size_t)s->cap_events, sizeof(orc_event));
    s->ramo_current_emit = (double *)calloc((size_t)ORC_MAX_SECTIONS * ORC_MAX_EMITTERS, sizeof(double));
    return s;
}

void orc_store_free(orc_store *s)
{
    if (!s) return;
    free(s->pos); free(s->prev_pos); free(s->vel); free(s->acc); free(s->acc_prev); free(s->acc_prev2);
    free(s->charge); free(s->mass); free(s->species); free(s->step); free(s->emitter); free(s->section);
    free(s->life); free(s->id); free(s->mask); free(s->events); free(s->ramo_current_emit);
    free(s);
}

void orc_store_clear_events(orc_store *s) { s->n_events = 0; }

static void push_event(orc_store *s, int kind, int plane, int i)
{
    orc_event *e;
    if (s->n_events == s->cap_events) {
        s->cap_events *= 2;
        s->events = (orc_event *)realloc(s->events, (size_t)s->cap_events * sizeof(orc_event));
    }
    e = &s->events[s->n_events++];
    e->kind = kind; e->plane = plane; e->index = i;
    e->x = s->pos[3 * i] / K_LEN; e->y = s->pos[3 * i + 1] / K_LEN;
    e->vx = s->vel[3 * i]; e->vy = s->vel[3 * i + 1]; e->vz = s->vel[3 * i + 2];
    e->emit = s->emitter[i]; e->sec = s->section[i]; e->id = s->id[i];
}

/* src/mod_pair.F90:29-159 */
int orc_add_particle(orc_store *s, const orc_params *p, const double pos[3], const double vel[3],
                     int species, int step, int emit, int life, int sec)
{
    int k = s->nrPart, c;
    if (k + 1 > s->capacity) { s->nrPart_dropped += 1; return -1; }
    for (c = 0; c < 3; ++c) {
        s->pos[3 * k + c] = pos[c];
        s->prev_pos[3 * k + c] = -1.0 * K_LEN;
        s->acc[3 * k + c] = 0.0; s->acc_prev[3 * k + c] = 0.0; s->acc_prev2[3 * k + c] = 0.0;
        s->vel[3 * k + c] = vel[c];
    }
    s->step[k] = step; s->mask[k] = 1; s->species[k] = species; s->emitter[k] = emit;
    s->section[k] = sec; s->life[k] = life; s->id[k] = s->nrID;
    if (species == ORC_SPECIES_ELEC) { s->nrElec += 1; s->charge[k] = -1.0 * K_Q0; s->mass[k] = 1.0 * K_M0; }
    else if (species == ORC_SPECIES_ION) { s->nrIon += 1; s->charge[k] = +1.0 * K_Q0; s->mass[k] = k_mN2p(); }
    else if (species == ORC_SPECIES_ATOM) { s->nrAtom += 1; s->charge[k] = 0.0; s->mass[k] = k_mN2(); }
    else return -2;
    /* seed the Beeman history with the vacuum-field acceleration, :133-139 */
    if (p) {
        double fE[3], f = s->charge[k] / s->mass[k];
        orc_field_E(p, pos, fE);
        for (c = 0; c < 3; ++c) {
            double a = f * fE[c];
            s->acc[3 * k + c] = a; s->acc_prev[3 * k + c] = a; s->acc_prev2[3 * k + c] = a;
        }
    }
    s->nrPart = s->nrElec + s->nrIon + s->nrAtom;
    s->charge_rev += 1;
    s->nrID += 1;
    return k;
}

/* src/mod_pair.F90:169-339 */
void orc_mark_particle_remove(orc_store *s, int i, int reason)
{
    int sp;
    if (!s->mask[i]) return; /* already marked */
    sp = s->species[i];
    if (sp != ORC_SPECIES_ELEC && sp != ORC_SPECIES_ION && sp != ORC_SPECIES_ATOM) return;
    s->mask[i] = 0;
    s->charge[i] = 0.0;
    s->charge_rev += 1;
    s->nrPart_remove += 1;
    if (sp == ORC_SPECIES_ELEC) {
        s->nrElec_remove += 1;
        if (reason == ORC_REMOVE_TOP) { s->nrPart_remove_top += 1; s->nrElec_remove_top += 1; push_event(s, 1, -1, i); }
        else if (reason == ORC_REMOVE_BOT) { s->nrPart_remove_bot += 1; s->nrElec_remove_bot += 1; push_event(s, 2, -1, i); }
    } else if (sp == ORC_SPECIES_ION) {
        s->nrIon_remove += 1;
        if (reason == ORC_REMOVE_TOP) { s->nrPart_remove_top += 1; s->nrIon_remove_top += 1; }
        else if (reason == ORC_REMOVE_BOT) { s->nrPart_remove_bot += 1; s->nrIon_remove_bot += 1; }
    } else {
        s->nrAtom_remove += 1;
    }
}

/* compact_array_*: src/mod_pair.F90:1180-1223 ; record_lifetime :1135-1160 */
static void compact_d3(double *A, const int *mask, int k, int m)
{
    int i, j = k;
    for (i = k; i < m; ++i) if (mask[i]) { A[3 * j] = A[3 * i]; A[3 * j + 1] = A[3 * i + 1]; A[3 * j + 2] = A[3 * i + 2]; ++j; }
}
static void compact_d1(double *A, const int *mask, int k, int m)
{
    int i, j = k;
    for (i = k; i < m; ++i) if (mask[i]) { A[j] = A[i]; ++j; }
}
static void compact_i1(int *A, const int *mask, int k, int m)
{
    int i, j = k;
    for (i = k; i < m; ++i) if (mask[i]) { A[j] = A[i]; ++j; }
}

/* src/mod_pair.F90:352-562 */
void orc_remove_particles(orc_store *s, int step)
{
    if ((s->nrPart_remove > 0) && (s->nrPart > 0)) {
        if ((s->nrPart - s->nrPart_remove) > 0) {
            int m = s->nrPart, k = 0, i, j;
            while (k < m && s->mask[k]) ++k; /* First_Dead_Index */
            compact_d3(s->pos, s->mask, k, m);
            compact_d3(s->prev_pos, s->mask, k, m);
            compact_d3(s->vel, s->mask, k, m);
            compact_d3(s->acc, s->mask, k, m);
            compact_d3(s->acc_prev, s->mask, k, m);
            compact_d3(s->acc_prev2, s->mask, k, m);
            /* record_lifetime compacts particles_step and bins the dead ones (before species is compacted) */
            j = k;
            for (i = k; i < m; ++i) {
                if (s->mask[i]) { s->step[j] = s->step[i]; ++j; }
                else {
                    int lt = step - s->step[i], sp = s->species[i];
                    if (lt <= 0) lt = 1;
                    if (lt > ORC_MAX_LIFE_TIME) lt = ORC_MAX_LIFE_TIME;
                    if (sp >= 1 && sp <= 3) s->life_time[lt][sp] += 1;
                }
            }
            compact_i1(s->species, s->mask, k, m);
            compact_d1(s->mass, s->mask, k, m);
            compact_d1(s->charge, s->mask, k, m);
            compact_i1(s->emitter, s->mask, k, m);
            compact_i1(s->section, s->mask, k, m);
            compact_i1(s->life, s->mask, k, m);
            compact_i1(s->id, s->mask, k, m);
        }
        s->nrElec -= s->nrElec_remove;
        s->nrIon -= s->nrIon_remove;
        s->nrAtom -= s->nrAtom_remove;
        if (s->nrElec < 0) s->nrElec = 0;
        if (s->nrIon < 0) s->nrIon = 0;
        s->nrPart = s->nrElec + s->nrIon + s->nrAtom;
        s->charge_rev += 1;
        {
            int i, lim = s->nrPart + s->nrPart_remove;
            if (lim > s->capacity) lim = s->capacity;
            for (i = 0; i < lim; ++i) s->mask[i] = 1;
        }
        s->nrPart_remove = 0; s->nrElec_remove = 0; s->nrIon_remove = 0; s->nrAtom_remove = 0;
        s->nrPart_remove_top = 0; s->nrPart_remove_bot = 0; s->nrElec_remove_top = 0; s->nrElec_remove_bot = 0;
        s->nrIon_remove_top = 0; s->nrIon_remove_bot = 0;
    }
}

/* src/mod_verlet.F90:197-232 with ptr_Check_Boundary (:325-338 planar,
 * src/mod_emission_tip.f90:1627-1647 tip) and Check_Planes (:343-367).
 * Serial i order = the order of the absorb / plane records. */
void orc_update_position(orc_store *s, const orc_params *p)
{
    const double dt = p->time_step, dt2 = p->time_step * p->time_step;
    int i, c, k;
    for (i = 0; i < s->nrPart; ++i) {
        double z;
        if (s->species[i] == ORC_SPECIES_ATOM) continue;
        for (c = 0; c < 3; ++c) {
            int e = 3 * i + c;
            s->prev_pos[e] = s->pos[e];
            s->pos[e] = s->pos[e] + s->vel[e] * dt + 1.0 / 6.0 * (4.0 * s->acc[e] - s->acc_prev[e]) * dt2;
            s->acc_prev2[e] = s->acc_prev[e];
            s->acc_prev[e] = s->acc[e];
            s->acc[e] = 0.0;
        }
        z = s->pos[3 * i + 2];
        if (z < 0.0) orc_mark_particle_remove(s, i, ORC_REMOVE_BOT);
        else if (z > p->box_dim[2]) orc_mark_particle_remove(s, i, ORC_REMOVE_TOP);
        if (p->geometry == ORC_GEOM_TIP) {
            double eta = orc_eta_coor(p, s->pos[3 * i], s->pos[3 * i + 1], z);
            if (eta < p->eta_1) orc_mark_particle_remove(s, i, ORC_REMOVE_BOT);
        }
        for (k = 0; k < p->planes_N; ++k) {
            double zp = p->planes_z[k];
            if (zp > 0.0) {
                double z_cur = s->pos[3 * i + 2], z_prev = s->prev_pos[3 * i + 2];
                if ((z_cur > zp) && (z_prev < zp)) push_event(s, 3, k, i);
            }
        }
    }
}

/* src/mod_verlet.F90:597-620 (CPU branch) */
void orc_update_acceleration(orc_store *s, const orc_params *p)
{
    if (p->geometry == ORC_GEOM_PLANAR) orc_accel_planar(p, s->nrPart, s->pos, s->charge, s->mass, s->species, s->acc);
    else orc_accel_generic(p, s->nrPart, s->pos, s->charge, s->mass, s->species, s->acc);
}

/* src/mod_verlet.F90:449-509 and :428-447 */
void orc_update_velocity(orc_store *s, const orc_params *p)
{
    const double dt = p->time_step;
    int i, c;
    for (c = 0; c < 4; ++c) s->ramo_current[c] = 0.0;
    for (i = 0; i < ORC_MAX_SECTIONS * ORC_MAX_EMITTERS; ++i) s->ramo_current_emit[i] = 0.0; /* src/mod_verlet.F90:155 */
    for (c = 0; c < 3; ++c) { s->avg_part_vel[c] = 0.0; s->avg_elec_vel[c] = 0.0; s->avg_ion_vel[c] = 0.0; }
    for (i = 0; i < s->nrPart; ++i) {
        double E_zu[3], EzV, qq;
        int sp = s->species[i];
        if (sp == ORC_SPECIES_ATOM) continue;
        for (c = 0; c < 3; ++c) {
            int e = 3 * i + c;
            s->vel[e] = s->vel[e] + 1.0 / 6.0 * (2.0 * s->acc[e] + 5.0 * s->acc_prev[e] - s->acc_prev2[e]) * dt;
        }
        qq = s->charge[i];
        orc_E_zunit(p, &s->pos[3 * i], E_zu);
        EzV = s->vel[3 * i] * E_zu[0] + s->vel[3 * i + 1] * E_zu[1] + s->vel[3 * i + 2] * E_zu[2];
        if (sp >= 0 && sp < 4) s->ramo_current[sp] = s->ramo_current[sp] + qq * EzV;
        {   /* ramo_current_emit(sec, emit), src/mod_verlet.F90:489-492 (an index outside the array is undefined
             * behaviour in the reference; such particles are skipped here) */
            int sec = s->section[i], emit = s->emitter[i];
            if (sec >= 1 && sec <= ORC_MAX_SECTIONS && emit >= 1 && emit <= ORC_MAX_EMITTERS) {
                double *r = &s->ramo_current_emit[(size_t)(emit - 1) * ORC_MAX_SECTIONS + (sec - 1)];
                *r = *r + qq * EzV;
            }
        }
        for (c = 0; c < 3; ++c) {
            if (sp == ORC_SPECIES_ELEC) s->avg_elec_vel[c] = s->avg_elec_vel[c] + s->vel[3 * i + c];
            else if (sp == ORC_SPECIES_ION) s->avg_ion_vel[c] = s->avg_ion_vel[c] + s->vel[3 * i + c];
            s->avg_part_vel[c] = s->avg_part_vel[c] + s->vel[3 * i + c];
        }
    }
    for (c = 0; c < 3; ++c) {
        if (s->nrPart != 0) s->avg_part_vel[c] = s->avg_part_vel[c] / s->nrPart;
        if (s->nrElec != 0) s->avg_elec_vel[c] = s->avg_elec_vel[c] / s->nrElec;
        if (s->nrIon != 0) s->avg_ion_vel[c] = s->avg_ion_vel[c] / s->nrIon;
    }
}

/* src/mod_verlet.F90:123-162 */
void orc_step(orc_store *s, const orc_params *p)
{
    orc_update_position(s, p);
    orc_update_acceleration(s, p);
    orc_update_velocity(s, p);
}

/* ---------------------------------------------------------------------------
 * Fowler-Nordheim helpers.  src/mod_field_emission_v2.F90:515-625.
 * The work function value w_theta(x,y) is passed in by the caller.
 */
double orc_fn_v_y(const orc_params *p, double F, double w_theta)
{
    if (p->image_charge) {
        double l = k_lconst() * (-1.0 * F) / (w_theta * w_theta);
        if (l > 1.0) l = 1.0;
        return 1.0 - l + 1.0 / 6.0 * l * log(l);
    }
    return 1.0;
}
double orc_fn_t_y(const orc_params *p, double F, double w_theta)
{
    if (p->image_charge) {
        double l = k_lconst() * (-1.0 * F) / (w_theta * w_theta);
        if (l > 1.0) l = 1.0;
        return 1.0 + l * (1.0 / 9.0 - 1.0 / 18.0 * log(l));
    }
    return 1.0;
}
double orc_fn_escape_prob_log(const orc_params *p, double F, double w_theta)
{
    double sw = sqrt(w_theta);
    return k_bFN() * (sw * sw * sw) * orc_fn_v_y(p, F, w_theta) / (-1.0 * F);
}
double orc_fn_elec_supply_log(const orc_params *p, double F, double w_theta)
{
    return 2.0 * log(-1.0 * F) - 2.0 * log(orc_fn_t_y(p, F, w_theta)) - log(w_theta);
}
double orc_fn_elec_supply_v2(const orc_params *p, double F, double w_theta)
{
    double t = orc_fn_t_y(p, F, w_theta);
    double time_step_div_q0 = p->time_step / K_Q0;
    return time_step_div_q0 * k_aFN() / ((t * t) * w_theta) * (F * F);
}
/* src/mod_emission_tip.f90:1657-1764 */
double orc_tip_v_y(const orc_params *p, double F, double w_theta) { return orc_fn_v_y(p, F, w_theta); }
double orc_tip_t_y(const orc_params *p, double F, double w_theta) { return orc_fn_t_y(p, F, w_theta); }
double orc_tip_escape_prob(const orc_params *p, double F, double w_theta)
{
    double sw = sqrt(w_theta);
    return exp(k_bFN() * (sw * sw * sw) * orc_tip_v_y(p, F, w_theta) / fabs(F));
}
double orc_tip_elec_supply(const orc_params *p, double A, double F, double w_theta)
{
    double t = orc_tip_t_y(p, F, w_theta);
    return A * k_aFN() * (F * F) * p->time_step / (K_Q0 * w_theta * (t * t));
}

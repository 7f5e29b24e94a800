/*
 * rumdeed_oracle_collisions.c -- CPU restatement of RUMDEED's electron / N2 collision step.
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE: see rumdeed_oracle_collisions.h.
 *
 * Follows src/mod_collisions.F90 and src/mod_polynomialroots.F90 statement by statement (file:line per function);
 * the polynomial solver keeps the original's GO TO structure as labelled blocks so that every branch can be
 * compared with the Fortran by eye.
 */
#include "rumdeed_oracle_collisions.h"

#include <complex.h>
#include <float.h>
#include <math.h>
#include <stddef.h>

/* src/mod_global.F90:26-75 (same literals as rumdeed_oracle.c) */
#define C_PI 3.141592653589793238462643383279502884197169399375105820974944592307816406286
#define C_H 6.62607015e-34
#define C_C 299792458.0
#define C_MU0 1.25663706212e-6
#define C_M0 9.1093837015e-31
#define C_Q0 1.602176634e-19
#define C_LEN 1.0e-9

static double c_eps0(void) { return 1.0 / (C_MU0 * (C_C * C_C)); }

/* src/mod_global.F90:50-68 */
void orc_coll_get_constants(orc_coll_constants *c)
{
    const double e0 = c_eps0();
    c->R_inf = C_M0 * (C_Q0 * C_Q0 * C_Q0 * C_Q0) / (8.0 * (e0 * e0) * (C_H * C_H * C_H) * C_C);
    c->Ryd = C_H * C_C * c->R_inf / C_Q0;
    c->N_n = 2.0;
    c->N_bind = 15.581;
    c->Z_eff = sqrt(c->N_bind * (c->N_n * c->N_n) / c->Ryd);
    c->Z_eff2 = c->Z_eff * c->Z_eff;
}

/* src/mod_collisions.F90:1872-1876 */
double orc_normal_dist(double mu, double sigma, double x)
{
    return 1.0 / (sqrt(2.0 * C_PI) * sigma) * exp(-((x - mu) * (x - mu)) / (2.0 * (sigma * sigma)));
}

/* src/mod_collisions.F90:1880-1886 */
double orc_folded_normal_dist(double mu, double sigma, double x)
{
    const double sigma2 = sigma * sigma;
    return sqrt(2.0 / (C_PI * sigma2)) * exp(-1.0 * (mu * mu + x * x) / (2.0 * sigma2)) * cosh(mu * x / sigma2);
}

/* src/mod_collisions.F90:1892-1904: maximum over a 0.1 degree grid on [0, 180] */
double orc_folded_normal_max(double mu, double sigma)
{
    const int n_grid = 1801;
    double best = 0.0;
    for (int k = 0; k <= n_grid - 1; ++k) {
        const double x = 180.0 * k / (n_grid - 1);
        const double f = orc_folded_normal_dist(mu, sigma, x);
        if (f > best) best = f;
    }
    return best;
}

/* src/mod_collisions.F90:1443-1449 */
double orc_kramers_cross_section(double energy)
{
    orc_coll_constants k;
    orc_coll_get_constants(&k);
    const double Z4 = (k.Z_eff * k.Z_eff) * (k.Z_eff * k.Z_eff);
    return 2.105e-26 * (k.Ryd * k.Ryd) * Z4 / (k.N_n * energy * ((k.N_n * k.N_n) * energy + k.Ryd * (k.Z_eff * k.Z_eff)));
}

/* BinarySearch, src/mod_global.F90:649-693, with 0-based indices */
int orc_binary_search(const double *list, int n, double value, int *i1, int *i2)
{
    int first = 0, last = n - 1;
    if ((last - first) < 1) {
        if (i1) *i1 = first;
        if (i2) *i2 = last;
        return first;
    }
    for (;;) {
        if ((last - first) == 1) break;
        /* Fortran: mid = (first + last)/2 on 1-based indices; identical bracketing needs the same midpoint */
        const int mid = ((first + 1) + (last + 1)) / 2 - 1;
        if (list[mid] > value) last = mid;
        else first = mid;
    }
    if (i1) *i1 = first;
    if (i2) *i2 = last;
    return (fabs(list[first] - value) < fabs(list[last] - value)) ? first : last;
}

static double cross_interp(const double *en, const double *dat, int n, double energy)
{
    int i1, i2;
    double e = energy;
    if (e < en[0]) e = en[0];          /* min(max(energy, first), last) */
    if (e > en[n - 1]) e = en[n - 1];
    orc_binary_search(en, n, e, &i1, &i2);
    const double y1 = dat[i1], y2 = dat[i2], x1 = en[i1], x2 = en[i2];
    const double h = (y1 - y2) / (x1 - x2);
    const double q = (y2 * x1 - y1 * x2) / (x1 - x2);
    return (h * e + q) * 1.0e-20;
}

/* src/mod_collisions.F90:1985-2012 */
double orc_find_cross_tot_data(const orc_cross_tables *T, double energy)
{
    const double a = 7.98, b = -0.005845, c = 4.628, d = -0.0007864;
    if ((energy > 70.0) && (energy <= 3000.0)) return (a * exp(b * energy) + c * exp(d * energy)) * 1.0e-20;
    return cross_interp(T->tot_energy, T->tot_data, T->n_tot, energy);
}

/* src/mod_collisions.F90:2014-2044 */
double orc_find_cross_ion_data(const orc_cross_tables *T, double energy)
{
    const double a = 2.251, b = -0.00311, c = 1.04, d = -0.0003378;
    if ((energy > 180.0) && (energy <= 3000.0)) return (a * exp(b * energy) + c * exp(d * energy)) * 1.0e-20;
    return cross_interp(T->ion_energy, T->ion_data, T->n_ion, energy);
}

/* Update_Collision_Data, src/mod_collisions.F90:2103-2134 */
void orc_update_collision_data(const orc_cross_tables *T, const double vel[3], double out[5])
{
    const double elec_max_speed2 = (2.0 * C_Q0 * 5000.0 / C_M0);
    const double nrm = sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]); /* norm2 */
    const double elec_cur_speed2 = nrm * nrm;
    double elec_energy;
    if (elec_cur_speed2 > elec_max_speed2) elec_energy = 0.5 * C_M0 * elec_max_speed2 / C_Q0;
    else elec_energy = 0.5 * C_M0 * elec_cur_speed2 / C_Q0;
    const double ion_cross_sec = orc_find_cross_ion_data(T, elec_energy);
    const double ion_cross_rad = sqrt(ion_cross_sec / C_PI);
    const double tot_cross_sec = orc_find_cross_tot_data(T, elec_energy);
    elec_energy = 0.5 * C_M0 * elec_cur_speed2 / C_Q0;
    const double recom_cross_rad = sqrt(orc_kramers_cross_section(elec_energy) / C_PI);
    out[0] = elec_energy;
    out[1] = ion_cross_sec;
    out[2] = ion_cross_rad;
    out[3] = recom_cross_rad;
    out[4] = tot_cross_sec;
}

/* ---- src/mod_polynomialroots.F90 ------------------------------------------------------------------------------ */
static int g_outputCode = 0; /* INTEGER,PRIVATE:: outputCode (module variable, :33) */
void orc_poly_reset_code(void) { g_outputCode = 0; }

#define EPS DBL_EPSILON
static double f_sign(double a, double b) { return copysign(fabs(a), b); } /* SIGN(a, b) */
static void swapd(double *a, double *b) { const double t = *b; *b = *a; *a = t; }

/* CubeRoot, :61-77 */
static double cube_root(double x)
{
    if (x < 0.0) return -exp(log(-x) / 3.0);
    if (x > 0.0) return exp(log(x) / 3.0);
    return 0.0;
}

/* QuadraticRoots, :125-175; a[0] + a[1] z + a[2] z^2 */
static void quadratic_roots(const double *a, double complex *z)
{
    if (a[0] == 0.0) {
        z[0] = 0.0;
        z[1] = -a[1] / a[2];
        g_outputCode = 21;
        return;
    }
    const double d = a[1] * a[1] - 4.0 * a[0] * a[2];
    if (fabs(d) <= 2.0 * EPS * a[1] * a[1]) {
        z[0] = -0.5 * a[1] / a[2];
        z[1] = z[0];
        g_outputCode = 22;
        return;
    }
    const double r = sqrt(fabs(d));
    if (d < 0.0) {
        const double x = -0.5 * a[1] / a[2];
        const double y = fabs(0.5 * r / a[2]);
        z[0] = x + y * I;
        z[1] = x - y * I;
        g_outputCode = 23;
        return;
    }
    if (a[1] != 0.0) {
        const double w = -(a[1] + f_sign(r, a[1]));
        z[0] = 2.0 * a[0] / w;
        z[1] = 0.5 * w / a[2];
        g_outputCode = 22;
        return;
    }
    const double x = fabs(0.5 * r / a[2]);
    z[0] = x;
    z[1] = -x;
    g_outputCode = 22;
}

/* CubicRoots, :178-333; a[0] + a[1] z + a[2] z^2 + a[3] z^3 */
static void cubic_roots(const double *a, double complex *z)
{
    const double RT3 = 1.7320508075689;
    double aq[3], arg, c, cf, d, p, p1, q, q1, r, ra, rb, rq, rt, r1, s, sf, sq, sum, t, tol, t1, w, w1, w2;
    double x, x1, x2, x3, y, y1, y2, y3;

    if (a[0] == 0.0) {
        z[0] = 0.0;
        quadratic_roots(a + 1, z + 1);
        return;
    }
    p = a[2] / (3.0 * a[3]);
    q = a[1] / a[3];
    r = a[0] / a[3];
    tol = 4.0 * EPS;

    c = 0.0;
    t = a[1] - p * a[2];
    if (fabs(t) > tol * fabs(a[1])) c = t / a[3];

    t = 2.0 * p * p - q;
    if (fabs(t) <= tol * fabs(q)) t = 0.0;
    d = r + p * t;
    if (fabs(d) <= tol * fabs(r)) goto L110;

    s = fmax(fmax(fabs(a[0]), fabs(a[1])), fabs(a[2]));
    p1 = a[2] / (3.0 * s);
    q1 = a[1] / s;
    r1 = a[0] / s;

    t1 = q - 2.25 * p * p;
    if (fabs(t1) <= tol * fabs(q)) t1 = 0.0;
    w = 0.25 * r1 * r1;
    w1 = 0.5 * p1 * r1 * t;
    w2 = q1 * q1 * t1 / 27.0;

    if (w1 >= 0.0) { w = w + w1; sq = w + w2; }
    else if (w2 < 0.0) { sq = w + (w1 + w2); }
    else { w = w + w2; sq = w + w1; }

    if (fabs(sq) <= tol * w) sq = 0.0;
    rq = fabs(s / a[3]) * sqrt(fabs(sq));
    if (sq >= 0.0) goto L40;

    /* all roots are real */
    arg = atan2(rq, -0.5 * d);
    cf = cos(arg / 3.0);
    sf = sin(arg / 3.0);
    rt = sqrt(-c / 3.0);
    y1 = 2.0 * rt * cf;
    y2 = -rt * (cf + RT3 * sf);
    y3 = -(d / y1) / y2;

    x1 = y1 - p;
    x2 = y2 - p;
    x3 = y3 - p;

    if (fabs(x1) > fabs(x2)) swapd(&x1, &x2);
    if (fabs(x2) > fabs(x3)) swapd(&x2, &x3);
    if (fabs(x1) > fabs(x2)) swapd(&x1, &x2);

    w = x3;

    if (fabs(x2) < 0.1 * fabs(x3)) goto L70;
    if (fabs(x1) < 0.1 * fabs(x2)) x1 = -(r / x3) / x2;
    z[0] = x1;
    z[1] = x2;
    z[2] = x3;
    return;

L40: /* real and complex roots */
    ra = cube_root(-0.5 * d - f_sign(rq, d));
    rb = -c / (3.0 * ra);
    t = ra + rb;
    w = -p;
    x = -p;
    if (fabs(t) <= tol * fabs(ra)) goto L41;
    w = t - p;
    x = -0.5 * t - p;
    if (fabs(x) <= tol * fabs(p)) x = 0.0;
L41:
    t = fabs(ra - rb);
    y = 0.5 * RT3 * t;

    if (t <= tol * fabs(ra)) goto L60;
    if (fabs(x) < fabs(y)) goto L50;
    s = fabs(x);
    t = y / x;
    goto L51;
L50:
    s = fabs(y);
    t = x / y;
L51:
    if (s < 0.1 * fabs(w)) goto L70;
    w1 = w / s;
    sum = 1.0 + t * t;
    if (w1 * w1 < 0.01 * sum) w = -((r / sum) / s) / s;
    z[0] = w;
    z[1] = x + y * I;
    z[2] = x - y * I;
    return;

L60: /* at least two roots are equal */
    if (fabs(x) < fabs(w)) goto L61;
    if (fabs(w) < 0.1 * fabs(x)) w = -(r / x) / x;
    z[0] = w;
    z[1] = x;
    z[2] = z[1];
    return;
L61:
    if (fabs(x) < 0.1 * fabs(w)) goto L70;
    z[0] = x;
    z[1] = z[0];
    z[2] = w;
    return;

L70: /* w is much larger in magnitude than the other roots */
    aq[0] = a[0];
    aq[1] = a[1] + a[0] / w;
    aq[2] = -a[3] * w;
    quadratic_roots(aq, z);
    z[2] = w;
    if (cimag(z[0]) == 0.0) return;
    z[2] = z[1];
    z[1] = z[0];
    z[0] = w;
    return;

L110: /* case when d = 0 */
    z[0] = -p;
    w = sqrt(fabs(c));
    if (c < 0.0) goto L120;
    z[1] = -p + w * I;
    z[2] = -p - w * I;
    return;
L120:
    if (p != 0.0) goto L130;
    z[1] = w;
    z[2] = -w;
    return;
L130:
    x = -(p + f_sign(w, p));
    z[2] = x;
    t = 3.0 * a[0] / (a[2] * x);
    if (fabs(p) > fabs(t)) goto L131;
    z[1] = t;
    return;
L131:
    z[1] = z[0];
    z[0] = t;
}

/* SelectSort, :512-526 (MINLOC returns the first minimum) */
static void select_sort4(double *a)
{
    for (int j = 0; j < 3; ++j) {
        int k = j;
        for (int m = j + 1; m < 4; ++m) if (a[m] < a[k]) k = m;
        if (j != k) swapd(&a[k], &a[j]);
    }
}

/* QuarticRoots, :336-510.  Its dummy argument outputCode is associated with the module variable by the only
 * caller (SolvePolynomial, :545), so assignments here and in the cubic / quadratic helpers hit the same variable. */
static void quartic_roots(const double *a, double complex *z)
{
    double complex w;
    double b, b2, c, d, e, h, p, q, r, t, temp[4], u, v, v1, v2, x, x1, x2, x3, y;

    if (a[0] == 0.0) {
        z[0] = 0.0;
        cubic_roots(a + 1, z + 1);
        return;
    }
    b = a[3] / (4.0 * a[4]);
    c = a[2] / a[4];
    d = a[1] / a[4];
    e = a[0] / a[4];
    b2 = b * b;

    p = 0.5 * (c - 6.0 * b2);
    q = d - 2.0 * b * (c - 4.0 * b2);
    r = b2 * (c - 3.0 * b2) - b * d + e;

    temp[0] = -q * q / 64.0;
    temp[1] = 0.25 * (p * p - r);
    temp[2] = p;
    temp[3] = 1.0;
    cubic_roots(temp, z);
    if (cimag(z[1]) != 0.0) goto L60;

    /* the resolvent cubic has only real roots: reorder them in increasing order */
    x1 = creal(z[0]);
    x2 = creal(z[1]);
    x3 = creal(z[2]);
    if (x1 > x2) swapd(&x1, &x2);
    if (x2 > x3) swapd(&x2, &x3);
    if (x1 > x2) swapd(&x1, &x2);

    u = 0.0;
    if (x3 > 0.0) u = sqrt(x3);
    if (x2 <= 0.0) goto L41;
    if (x1 >= 0.0) goto L30;
    if (fabs(x1) > x2) goto L40;
    x1 = 0.0;

L30:
    x1 = sqrt(x1);
    x2 = sqrt(x2);
    if (q > 0.0) x1 = -x1;
    temp[0] = ((x1 + x2) + u) - b;
    temp[1] = ((-x1 - x2) + u) - b;
    temp[2] = ((x1 - x2) - u) - b;
    temp[3] = ((-x1 + x2) - u) - b;
    select_sort4(temp);
    if (fabs(temp[0]) >= 0.1 * fabs(temp[3])) goto L31;
    t = temp[1] * temp[2] * temp[3];
    if (t != 0.0) temp[0] = e / t;
L31:
    z[0] = temp[0];
    z[1] = temp[1];
    z[2] = temp[2];
    z[3] = temp[3];
    g_outputCode = 31;
    return;

L40:
    v1 = sqrt(fabs(x1));
    v2 = 0.0;
    goto L50;
L41:
    v1 = sqrt(fabs(x1));
    v2 = sqrt(fabs(x2));
    if (q < 0.0) u = -u;

L50:
    x = -u - b;
    y = v1 - v2;
    z[0] = x + y * I;
    z[1] = x - y * I;
    x = u - b;
    y = v1 + v2;
    z[2] = x + y * I;
    z[3] = x - y * I;
    g_outputCode = 44;
    return;

L60: /* the resolvent cubic has complex roots */
    t = creal(z[0]);
    x = 0.0;
    if (t < 0.0) goto L61;
    else if (t == 0.0) goto L70;
    else goto L62;
L61:
    h = fabs(creal(z[1])) + fabs(cimag(z[1]));
    if (fabs(t) <= h) goto L70;
    goto L80;
L62:
    x = sqrt(t);
    if (q > 0.0) x = -x;

L70:
    w = csqrt(z[1]);
    u = 2.0 * creal(w);
    v = 2.0 * fabs(cimag(w));
    t = x - b;
    x1 = t + u;
    x2 = t - u;
    if (fabs(x1) <= fabs(x2)) goto L71;
    t = x1;
    x1 = x2;
    x2 = t;
L71:
    u = -x - b;
    h = u * u + v * v;
    if (x1 * x1 < 0.01 * fmin(x2 * x2, h)) x1 = e / (x2 * h);
    z[0] = x1;
    z[1] = x2;
    z[2] = u + v * I;
    z[3] = u - v * I;
    g_outputCode = 42;
    return;

L80:
    v = sqrt(fabs(t));
    z[0] = -b + v * I;
    z[1] = -b - v * I;
    z[2] = z[0];
    z[3] = z[1];
    g_outputCode = 23;
}

/* SolvePolynomial, :528-563 */
void orc_solve_polynomial(double quartic, double cubic, double quadratic, double linear, double constant,
                          int *code, double roots[8])
{
    double a[5];
    double complex z[5];
    int k;
    a[0] = constant; a[1] = linear; a[2] = quadratic; a[3] = cubic; a[4] = quartic;
    for (k = 0; k < 5; ++k) z[k] = NAN + NAN * I;
    for (k = 0; k < 8; ++k) roots[k] = NAN; /* unassigned intent(out) roots */

    if (quartic != 0.0) quartic_roots(a, z);
    else if (cubic != 0.0) cubic_roots(a, z);
    else if (quadratic != 0.0) quadratic_roots(a, z);
    else if (linear != 0.0) { z[0] = -constant / linear; g_outputCode = 1; }
    else g_outputCode = 0;

    *code = g_outputCode;
    if (g_outputCode > 0) { roots[0] = creal(z[0]); roots[1] = cimag(z[0]); }
    if (g_outputCode > 1) { roots[2] = creal(z[1]); roots[3] = cimag(z[1]); }
    if (g_outputCode > 23) { roots[4] = creal(z[2]); roots[5] = cimag(z[2]); }
    if (g_outputCode > 99) { roots[6] = creal(z[3]); roots[7] = cimag(z[3]); } /* never: root4 stays unassigned */
}

/* ---- Do_Discrete_Recombination_ots, src/mod_collisions.F90:86-245 ---------------------------------------------- */
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double norm2_3(const double *a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

/* The body of the electron loop for one pair, :128-196 */
int orc_recombination_pair(const double ion_pos[3], const double elec_pos[3], const double elec_vel[3],
                           const double elec_acc[3], double recom_rad, double time_step, double *t_out,
                           double *dist_out)
{
    double rel_pos[3], next[3], t = 0.0;
    int coll_happens = 0, c;
    for (c = 0; c < 3; ++c) rel_pos[c] = elec_pos[c] - ion_pos[c];
    const double cur_dist2 = dot3(rel_pos, rel_pos);
    const double recom_rad2 = recom_rad * recom_rad;

    if (cur_dist2 <= recom_rad2) {
        coll_happens = 1;
        t = 0.0;
    } else {
        const double a = 0.25 * dot3(elec_acc, elec_acc);
        const double b = dot3(elec_vel, elec_acc);
        const double cc = dot3(elec_vel, elec_vel) + dot3(rel_pos, elec_acc);
        const double dd = 2.0 * dot3(rel_pos, elec_vel);
        const double e = cur_dist2 - recom_rad2;
        int code;
        double z[8];
        orc_solve_polynomial(a, b, cc, dd, e, &code, z);
        if (code != 44 && code != 23) {
            if (code == 31) {
                int k;
                for (k = 0; k < 4; ++k) /* t1 .. t4 in order; t4 is the unassigned root4 (NaN: never true) */
                    if (z[2 * k + 1] == 0.0 && z[2 * k] > 0.0 && z[2 * k] <= time_step) { coll_happens = 1; t = z[2 * k]; break; }
            } else if (code == 42) {
                int k;
                for (k = 0; k < 2; ++k)
                    if (z[2 * k + 1] == 0.0 && z[2 * k] > 0.0 && z[2 * k] <= time_step) { coll_happens = 1; t = z[2 * k]; break; }
            }
        }
    }
    if (!coll_happens) return 0;
    for (c = 0; c < 3; ++c) next[c] = elec_pos[c] + elec_vel[c] * t + 0.5 * elec_acc[c] * (t * t);
    for (c = 0; c < 3; ++c) next[c] -= ion_pos[c];
    *t_out = t;
    *dist_out = norm2_3(next);
    return 1;
}

int orc_discrete_recombination_ots(int n, const double *pos, const double *vel, const double *acc,
                                   const int *species, int *mask, const int *life, const int *step_born,
                                   const int *emitter, const double *recom_rad, int step, double time_step,
                                   orc_recomb_event *events, int max_events, int *reason_out, int *n_expired)
{
    int nrRecombinations = 0, expired = 0, i, j, c;
    for (i = 0; i < n; ++i) {
        if ((species[i] != ORC_SPECIES_ION) || !mask[i]) continue;
        if (step >= life[i]) { /* end of life, :121-124 */
            mask[i] = 0;
            if (reason_out) reason_out[i] = ORC_REMOVE_TOP;
            ++expired;
            continue;
        }
        for (j = 0; j < n; ++j) {
            double t, dist;
            if ((species[j] != ORC_SPECIES_ELEC) || !mask[j]) continue;
            if (!orc_recombination_pair(&pos[3 * i], &pos[3 * j], &vel[3 * j], &acc[3 * j], recom_rad[j] * 1.0, time_step, &t, &dist))
                continue;
            /* serial execution: both masks are still true here (:211) */
            mask[i] = 0;
            mask[j] = 0;
            if (reason_out) { reason_out[i] = ORC_REMOVE_RECOM; reason_out[j] = ORC_REMOVE_RECOM; }
            if (nrRecombinations < max_events) {
                orc_recomb_event *ev = &events[nrRecombinations];
                ev->step = step;
                for (c = 0; c < 3; ++c) ev->ion_pos[c] = pos[3 * i + c];
                ev->elec_speed = norm2_3(&vel[3 * j]);
                ev->dist = dist;
                ev->recom_rad = recom_rad[j] * 1.0;
                ev->elec_slot = j;
                ev->ion_slot = i;
                ev->elec_emit = emitter[j];
                ev->ion_life = step - step_born[i];
                ev->t = t;
            }
            ++nrRecombinations;
            break; /* cycle ion */
        }
    }
    if (n_expired) *n_expired = expired;
    return nrRecombinations;
}

/* ---- scattering directions, src/mod_collisions.F90:1452-1577 ----------------------------------------------------- */
static void accept_reject_direction(orc_rng *r, double mu, double sigma, const double par_vel[3], double out[3])
{
    const double m_factor = orc_folded_normal_max(mu, sigma);
    const double len_vel = sqrt(par_vel[0] * par_vel[0] + par_vel[1] * par_vel[1] + par_vel[2] * par_vel[2]);
    double par_vec[3], len_vec, alpha;
    int n_tries = 0;
    for (;;) {
        par_vec[0] = orc_rng_uniform(r) - 0.5;
        par_vec[1] = orc_rng_uniform(r) - 0.5;
        par_vec[2] = orc_rng_uniform(r) - 0.5;
        len_vec = sqrt(par_vec[0] * par_vec[0] + par_vec[1] * par_vec[1] + par_vec[2] * par_vec[2]);
        if ((len_vel > 0.0) && (len_vec > 0.0)) {
            const double dot_p = par_vel[0] * par_vec[0] + par_vel[1] * par_vec[1] + par_vel[2] * par_vec[2];
            const double angle = acos(dot_p / (len_vec * len_vel)) * 180.0 / C_PI;
            alpha = orc_folded_normal_dist(mu, sigma, angle) / m_factor;
        } else {
            alpha = 1.0;
        }
        if (orc_rng_uniform(r) < alpha) break;
        if (++n_tries >= 1000000) break;
    }
    out[0] = par_vec[0] / len_vec;
    out[1] = par_vec[1] / len_vec;
    out[2] = par_vec[2] / len_vec;
}

void orc_get_injected_vec(orc_rng *r, double T, const double par_vel[3], double out[3])
{
    (void)T;
    accept_reject_direction(r, 5.0, 25.0, par_vel, out); /* mu = 5, sigma = 25, :1456 */
}

void orc_get_ejected_vec(orc_rng *r, double W, double T, const double par_vel[3], double out[3])
{
    const double a = -430.5, b = -0.5445, c = 89.32, sigma = 48.0;
    double angle_max;
    (void)W;
    if (T < 100.0) angle_max = a * pow(100.0, b) + c;
    else angle_max = a * pow(T, b) + c;
    accept_reject_direction(r, angle_max, sigma, par_vel, out);
}

/* ---- Do_Continuous_Ionization_ots, src/mod_collisions.F90:558-705 ------------------------------------------------ */
int orc_continuous_ionization_ots(orc_rng *r, const orc_cross_tables *T, int n, const double *pos,
                                  const double *prev_pos, double *vel, const int *species, const int *mask,
                                  int *emitter, double n_d, double cyl_radius, int step,
                                  orc_ionization_event *events, int max_events, int *nrCollisions)
{
    orc_coll_constants k;
    int nrIonizations = 0, ncoll = 0, i, c;
    orc_coll_get_constants(&k);
    for (i = 0; i < n; ++i) {
        double cd[5], d[3], direct_vec[3], nrm;
        if ((species[i] != ORC_SPECIES_ELEC) || !mask[i]) continue;
        const double *p = &pos[3 * i];
        if (!(sqrt(p[0] * p[0] + p[1] * p[1]) <= cyl_radius)) continue;
        orc_update_collision_data(T, &vel[3 * i], cd); /* what Update_Collision_Data_All_ots stored for slot i */
        const double elec_energy = cd[0];
        for (c = 0; c < 3; ++c) d[c] = p[c] - prev_pos[3 * i + c];
        const double elec_cur_path = norm2_3(d);
        const double elec_cur_speed = norm2_3(&vel[3 * i]);
        if (!(elec_energy > k.N_bind)) continue;
        const double cross_tot = cd[4];
        const double mean_path = 1.0 / (n_d * cross_tot);
        double alpha = elec_cur_path / mean_path;
        if (!(orc_rng_uniform(r) < alpha)) continue;
        {
            const double cross_ion = cd[1];
            alpha = cross_ion / cross_tot;
            if (orc_rng_uniform(r) < alpha) {
                const double E1 = elec_energy, E2 = E1 - k.N_bind;
                const double collE = E2 * orc_rng_uniform(r);
                const double ejecE = E2 - collE;
                double par_vel[3] = {vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]};
                orc_ionization_event ev;
                ev.step = step; ev.in_slot = i; ev.E1 = E1; ev.collE = collE; ev.ejecE = ejecE;
                ev.elec_emit = emitter[i];
                for (c = 0; c < 3; ++c) ev.pos[c] = p[c];
                /* colliding electron */
                orc_get_injected_vec(r, elec_energy, par_vel, direct_vec);
                nrm = norm2_3(direct_vec);
                for (c = 0; c < 3; ++c) direct_vec[c] = direct_vec[c] / nrm;
                for (c = 0; c < 3; ++c) { vel[3 * i + c] = direct_vec[c] * sqrt(2.0 * C_Q0 * collE / C_M0); ev.new_vel[c] = vel[3 * i + c]; }
                ev.in_speed = elec_cur_speed;
                ev.out_speed = norm2_3(&vel[3 * i]);
                /* ejected electron */
                for (c = 0; c < 3; ++c) ev.ejec_pos[c] = orc_rng_uniform(r);
                for (c = 0; c < 3; ++c) ev.ejec_pos[c] = p[c] + (2.0 * (ev.ejec_pos[c] - 0.5)) * C_LEN;
                orc_get_ejected_vec(r, elec_energy, elec_energy, par_vel, direct_vec);
                nrm = norm2_3(direct_vec);
                for (c = 0; c < 3; ++c) direct_vec[c] = direct_vec[c] / nrm;
                for (c = 0; c < 3; ++c) ev.ejec_vel[c] = direct_vec[c] * sqrt(2.0 * C_Q0 * ejecE / C_M0);
                ev.new_speed = norm2_3(ev.ejec_vel);
                /* created ion */
                for (c = 0; c < 3; ++c) ev.ion_pos[c] = orc_rng_uniform(r);
                for (c = 0; c < 3; ++c) ev.ion_pos[c] = p[c] + (2.0 * (ev.ion_pos[c] - 0.5)) * C_LEN;
                if (nrIonizations < max_events) events[nrIonizations] = ev;
                ++nrIonizations;
                emitter[i] = 2; /* ion_emitter, src/mod_global.F90:121 */
            }
        }
        ++ncoll;
    }
    if (nrCollisions) *nrCollisions = ncoll;
    return nrIonizations;
}

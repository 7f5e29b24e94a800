/*
 * boundary_check.c -- a plain C host that links librumdeed_b200.so the way the Fortran host would
 * (-lrumdeed_b200, include/rumdeed_b200.h) and walks the call sequence of INTEGRATION.md section 2 with
 * Fortran-layout arrays: (3,N) column-major double precision, default integers, 1-based slots on the
 * host side converted at the boundary.  The closest executable stand-in for fortran/mod_b200_bridge.F90,
 * which cannot be compiled in an image without a Fortran compiler.
 *
 *   Init_*            -> rb2_init                      (src/main.F90:146)
 *   Add_Particle      -> rb2_add_particles             (src/mod_pair.F90:29)
 *   Calc_Field_at     -> rb2_field_batch (M = 1)       (src/mod_verlet.F90:1466)
 *   Calc_Field_at_Batch                                 (src/mod_verlet.F90:1635)
 *   Update_Position   -> rb2_step + rb2_get_events     (src/main.F90:190)
 *   Mark_Particles_Remove / Remove_Particles           (src/mod_pair.F90:169, :352)
 *   Clean_up          -> rb2_finalize                  (src/main.F90:266)
 *
 * Exit codes: 0 all checks passed; 3 no usable GPU (rb2_init refused: there is no CPU fallback); 1 a check failed.
 * The known answers are the reference's own unit-test vectors (src/mod_tests.F90:431-437, :500-515, :1403-1452).
 */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "rumdeed_b200.h"

#define CHECK(call)                                                                          \
    do {                                                                                     \
        int rc__ = (call);                                                                   \
        if (rc__ != RB2_OK) {                                                                \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc__, rb2_last_error_string());         \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)
#define EXPECT(cond)                                                     \
    do {                                                                 \
        if (!(cond)) {                                                   \
            fprintf(stderr, "line %d: check failed: %s\n", __LINE__, #cond); \
            return 1;                                                    \
        }                                                                \
    } while (0)

static double rel(double a, double b) { return fabs(a - b) / fabs(b); }

int main(void)
{
    const double nm = 1.0e-9, q_0 = 1.602176634e-19, m_0 = 9.1093837015e-31;
    rb2_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    /* Test_Acceleration_Without_Image_Charge, src/mod_tests.F90:405-515: d = 100 nm, V = 2 V */
    cfg.geometry = RB2_GEOM_PLANAR;
    cfg.image_charge = 0;
    cfg.N_ic_max = 0;
    cfg.V_s = 2.0;
    cfg.d = 100.0 * nm;
    cfg.E_z = -1.0 * cfg.V_s / cfg.d;
    cfg.box_dim[0] = cfg.box_dim[1] = cfg.box_dim[2] = 100.0 * nm;
    cfg.time_step = 0.25e-15;
    cfg.capacity = 64;
    cfg.device = -1;
    if (!rb2_device_available() || rb2_init(&cfg) != RB2_OK) {
        /* no CPU fallback: every compute entry point must refuse, none may compute */
        double f[3] = {0, 0, 0}, p[3] = {0, 0, 0};
        if (rb2_field_batch(1, p, f) == RB2_OK || rb2_accel_only() == RB2_OK) return 1;
        printf("no usable sm_100 device: %s\n", rb2_last_error_string());
        return 3;
    }
    /* particles_cur_pos(3, N) as the Fortran host holds it: xyz of particle 1, xyz of particle 2, ... */
    const double pos[9] = {3.0 * nm, -10.0 * nm, 2.0 * nm, -9.0 * nm, 26.0 * nm, 80.0 * nm, 6.0 * nm, -24.0 * nm, 56.53 * nm};
    const double vel[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const int species[3] = {RB2_SPECIES_ELEC, RB2_SPECIES_ELEC, RB2_SPECIES_ION};
    const int emit[3] = {1, 1, 1}, sec[3] = {1, 1, 1}, life[3] = {-1, -1, -1};
    CHECK(rb2_add_particles(3, pos, vel, species, 0, emit, sec, life));
    rb2_counts k;
    CHECK(rb2_get_counts(&k));
    EXPECT(k.nrPart == 3 && k.nrElec == 2 && k.nrIon == 1 && k.nrID == 3);
    /* the reference builds the probe point from single-precision literals (:500) */
    const double pt[3] = {(double)-4.55f * nm, (double)-2.34f * nm, (double)96.44f * nm};
    double field[3];
    CHECK(rb2_field_window_open());
    CHECK(rb2_field_batch(1, pt, field));
    CHECK(rb2_field_window_close());
    EXPECT(rel(field[0], -314559.29097098) < 0.02 && rel(field[1], 1423979.07058996) < 0.02 && rel(field[2], -20246038.87978313) < 0.02);
    /* batch == point by point (Test_Planar_Batch_Field, :1554-1603) */
    double pts[12], fb[12];
    for (int m = 0; m < 4; ++m) { pts[3 * m] = (10.0 * m - 20.0) * nm; pts[3 * m + 1] = 5.0 * m * nm; pts[3 * m + 2] = (m == 0) ? 0.0 : 20.0 * m * nm; }
    CHECK(rb2_field_batch(4, pts, fb));
    for (int m = 0; m < 4; ++m) {
        double f1[3];
        CHECK(rb2_field_batch(1, &pts[3 * m], f1));
        for (int c = 0; c < 3; ++c) EXPECT(fabs(f1[c] - fb[3 * m + c]) <= 1e-12 * fabs(f1[c]) + 1e-6);
    }
    /* Calculate_Acceleration_Particles: closed-form Coulomb on particle 1 from 2 and 3 (+ vacuum field), :443-472 */
    CHECK(rb2_accel_only());
    double acc[9], chg[3], mass[3];
    CHECK(rb2_download_particles(NULL, NULL, NULL, acc, NULL, NULL, chg, mass, NULL, NULL, NULL, NULL, NULL, NULL, NULL));
    {
        const double mu_0 = 1.25663706212e-6, c = 299792458.0, eps0 = 1.0 / (mu_0 * c * c), pi = 3.14159265358979323846;
        double a[3] = {0, 0, 0};
        for (int j = 1; j < 3; ++j) {
            double d[3], r2 = 0.0;
            for (int cc = 0; cc < 3; ++cc) { d[cc] = pos[cc] - pos[3 * j + cc]; r2 += d[cc] * d[cc]; }
            const double r = sqrt(r2), pre = chg[0] * chg[j] / (4.0 * pi * eps0) / (r * r * r);
            for (int cc = 0; cc < 3; ++cc) a[cc] += pre * d[cc];
        }
        a[2] += chg[0] * cfg.E_z;
        for (int cc = 0; cc < 3; ++cc) EXPECT(fabs(acc[cc] - a[cc] / mass[0]) <= 1e-9 * fabs(a[cc] / mass[0]));
        EXPECT(chg[0] == -q_0 && chg[2] == q_0 && mass[0] == m_0);
    }
    /* Mark_Particles_Remove(2, remove_top) (Fortran slot 2 = C slot 1), Remove_Particles: survivors keep order and ids */
    const int idx = 2 - 1, why = RB2_REMOVE_TOP;
    CHECK(rb2_mark_remove(1, &idx, &why));
    CHECK(rb2_remove_marked(1, &k));
    EXPECT(k.nrPart == 2 && k.nrElec == 1 && k.nrIon == 1);
    int id[2];
    double p2[6];
    CHECK(rb2_download_particles(p2, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, id, NULL));
    EXPECT(id[0] == 0 && id[1] == 2 && p2[3] == pos[6] && p2[5] == pos[8]);
    CHECK(rb2_finalize());

    /* Test_Beeman_Kinematics, src/mod_tests.F90:1403-1452: one electron, three Update_Position calls */
    cfg.V_s = 2.0e3; cfg.d = 1000.0 * nm; cfg.E_z = -cfg.V_s / cfg.d; cfg.box_dim[2] = 1000.0 * nm;
    cfg.planes_N = 1; cfg.planes_z[0] = 500.0000001 * nm;
    CHECK(rb2_init(&cfg));
    const double p0[3] = {0.0, 0.0, 500.0 * nm}, v0[3] = {1.0e3, 0.0, 0.0};
    const int one = RB2_SPECIES_ELEC, e1 = 1, l1 = -1;
    CHECK(rb2_add_particles(1, p0, v0, &one, 1, &e1, &e1, &l1));
    rb2_step_result r;
    int crossings = 0;
    for (int step = 1; step <= 3; ++step) {
        CHECK(rb2_step(step, &r));
        if (r.n_events > 0) {
            rb2_event ev[4];
            int n = 0;
            CHECK(rb2_get_events(4, ev, &n));
            EXPECT(n == r.n_events && ev[0].kind == 3 && ev[0].plane == 0 && ev[0].index == 0 && ev[0].id == 0);
            crossings += n;
        }
        CHECK(rb2_remove_marked(step, &k));
    }
    EXPECT(crossings == 1);
    double pf[3], vf[3];
    CHECK(rb2_download_particles(pf, NULL, vf, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL));
    const double a_z = q_0 * cfg.V_s / (m_0 * cfg.d), dt = cfg.time_step;
    EXPECT(rel(pf[2] - 500.0 * nm, 4.5 * a_z * dt * dt) < 1e-9);
    EXPECT(rel(vf[2], 3.0 * a_z * dt) < 1e-12 && rel(pf[0], 3.0 * 1.0e3 * dt) < 1e-12);
    EXPECT(rel(r.ramo_current[RB2_SPECIES_ELEC], q_0 * 3.0 * a_z * dt / cfg.d) < 1e-12);
    CHECK(rb2_finalize());
    printf("boundary_check ok\n");
    return 0;
}

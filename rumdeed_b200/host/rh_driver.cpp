// rh_driver.cpp -- Init / main-loop step / Clean_up of the host mirror and the output writers.
// Mirrors src/main.F90:100-266 (mode dispatch, time loop) and the writers of src/mod_pair.F90:774-867,
// src/mod_verlet.F90:2056 (volt.dt), :360 (planes), src/mod_field_emission_v2.F90:203 (emitted.dt).
#include <math.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <chrono>

#include "rh_host.hpp"

namespace rh {

static FILE *open_out(Sim &s, const char *name, const char *mode)
{
    const std::string p = s.out_dir + "/" + name;
    return fopen(p.c_str(), mode);
}

static void fill_config(Sim &s, int geometry)
{
    const Globals &g = s.g;
    rb2_config &c = s.cfg;
    memset(&c, 0, sizeof(c));
    c.geometry = geometry;
    c.image_charge = g.image_charge ? 1 : 0;
    c.N_ic_max = g.N_ic_max;
    c.planes_N = g.planes_N;
    c.V_s = g.V_s;
    c.d = g.d;
    c.E_z = -1.0 * g.V_d / g.d;  // Set_Voltage, src/mod_verlet.F90:2050-2052
    for (int k = 0; k < 3; ++k) c.box_dim[k] = g.box_dim[k];
    c.time_step = g.time_step;
    for (int k = 0; k < RB2_PLANES_MAX; ++k) c.planes_z[k] = g.planes_z[k];
    c.a_foci = s.a_foci; c.eta_1 = s.eta_1; c.shift_z = s.shift_z;
    c.pre_fac_E_tip = s.pre_fac_E_tip; c.pre_fac_E_tip_unit_voltage = s.pre_fac_E_tip_unit_voltage;
    c.h_tip = s.h_tip; c.r_tip = s.r_tip; c.max_xi = s.max_xi;
    c.capacity = g.max_particles;
    c.device = -1;
}

// src/main.F90:100-151 + Init (:416-741)
int Init(Sim &s)
{
    Globals &g = s.g;
    uint64_t seed = g.seed;
    if (seed == 0) {  // like the reference: seed from /dev/urandom (src/main.F90:747-793)
        FILE *f = fopen("/dev/urandom", "rb");
        if (!f || fread(&seed, sizeof(seed), 1, f) != 1) seed = 0x1234567ULL;
        if (f) fclose(f);
    }
    s.rng.seed(seed);
    s.rng_photo.seed(seed ^ 0x9E3779B97F4A7C15ULL);
    s.photo_cand.clear();
    s.photo_head = 0;
    int geometry = RB2_GEOM_PLANAR, rc = 0;
    switch (g.emission_mode) {  // src/main.F90:106-144
        case EMISSION_PHOTO: rc = Init_Photo_Emission(s); break;
        case EMISSION_TIP: rc = Init_Emission_Tip(s); geometry = RB2_GEOM_TIP; break;
        case EMISSION_FIELD_THERMO: rc = Init_Field_Thermo_Emission(s); break;
        case EMISSION_FIELD_V2: rc = Init_Field_Emission_v2(s); break;
        default:
            return s.fail("RUMDEED: ERROR UNKNOWN EMISSION MODEL (modes 1, 3, 9, 10 are on the device path): " + std::to_string(g.emission_mode));
    }
    if (rc) return rc;
    if (!s.ptr.ptr_Do_Emission) return s.fail("RUMDEED: ERROR ptr_Do_Emission is not associated");  // Check_Pointers, :820-847
    fill_config(s, geometry);
    if (s.check(rb2_init(&s.cfg), "rb2_init")) return -1;
    memset(&s.counts, 0, sizeof(s.counts));
    s.nrEmitted_total = s.nrAbsorbed_top = s.nrAbsorbed_bot = 0;
    s.ramo_integral = 0.0;
    if (s.write_files) {
        mkdir(s.out_dir.c_str(), 0755);
        s.ud_ramo = open_out(s, "ramo_current.dt", "w");
        s.ud_emit = open_out(s, "emitted.dt", "w");
        s.ud_absorb = open_out(s, "absorbed.dt", "w");
        s.ud_absorb_top = open_out(s, "absorbed_top.dt", "w");
        s.ud_absorb_bot = open_out(s, "absorbed_bot.dt", "w");
        s.ud_field = open_out(s, "field.dt", "w");
        s.ud_integrand = open_out(s, "integration.dt", "w");
        s.ud_volt = open_out(s, "volt.dt", "w");
        s.ud_density_emit = open_out(s, "density_emit.bin", "wb");
        s.ud_density_emit_elec = open_out(s, "density_emit_elec.bin", "wb");  // src/main.F90 Init: one file per species
        s.ud_density_emit_ion = open_out(s, "density_emit_ion.bin", "wb");
        s.ud_density_emit_atom = open_out(s, "density_emit_atom.bin", "wb");
        if (g.write_ramo_sec) s.ud_ramo_sec = open_out(s, "ramo_current.bin", "wb");  // src/main.F90:596
        if (g.write_position_file) s.ud_pos = open_out(s, "position.bin", "wb");  // src/main.F90:548
        s.ud_density_absorb_top = open_out(s, "density_absorb_top.bin", "wb");
        s.ud_density_absorb_bot = open_out(s, "density_absorb_bot.bin", "wb");
        for (int k = 0; k < g.planes_N; ++k) {
            char nm[64];
            snprintf(nm, sizeof(nm), "planes-%d.bin", k + 1);
            s.planes_ud[k] = open_out(s, nm, "wb");
        }
        // init.dt: the run parameters (src/main.F90:852-938, text part)
        if (FILE *f = open_out(s, "init.dt", "w")) {
            fprintf(f, "V_s = %.16E\nd = %.16E\ntime_step = %.16E\nsteps = %d\nemission_mode = %d\nimage_charge = %d\nN_ic_max = %d\n",
                    g.V_s, g.d, g.time_step, g.steps, g.emission_mode, (int)g.image_charge, g.N_ic_max);
            fclose(f);
        }
        // init.bin, src/main.F90:928-938: epsilon_r, m_eeff, m_ieff, length_scale, time_scale, vel_scale, cur_scale (f64),
        // MAX_PARTICLES, MAX_EMITTERS, MAX_SECTIONS, MAX_LIFE_TIME (i32) -- 72 bytes
        if (FILE *f = open_out(s, "init.bin", "wb")) {
            const double d7[7] = {1.0, 1.0, 1.0, length_scale, time_scale, length_scale / time_scale, cur_scale};
            const int i4[4] = {MAX_PARTICLES, MAX_EMITTERS, MAX_SECTIONS, RB2_MAX_LIFE_TIME};
            fwrite(d7, sizeof(double), 7, f);
            fwrite(i4, sizeof(int), 4, f);
            fclose(f);
        }
    }
    // ramo_current_emit(sec, emit): the device keeps a table of the sections the work function defines
    if (g.write_ramo_sec || s.ramo_sections > 0) {
        if (s.ramo_sections < 1) s.ramo_sections = std::max(1, std::min(MAX_SECTIONS, s.work.y_num * s.work.x_num));
        if (s.check(rb2_set_option("ramo_sections", (double)s.ramo_sections), "rb2_set_option(ramo_sections)")) return -1;
        s.ramo_current_emit.assign((size_t)MAX_SECTIONS * MAX_EMITTERS, 0.0);
    }
    if (Init_Collisions(s)) return -1;  // src/main.F90:397 (Read_Cross_Section_Data), :164-166
    return 0;
}

static void write_event_files(Sim &s)
{
    const int n = s.last.n_events;
    if (n <= 0) return;
    std::vector<rb2_event> ev((size_t)n);
    int got = 0;
    if (rb2_get_events(n, ev.data(), &got) != RB2_OK) return;
    for (int k = 0; k < got; ++k) {
        const rb2_event &e = ev[k];
        FILE *f = nullptr;
        if (e.kind == 1) f = s.ud_density_absorb_top;
        else if (e.kind == 2) f = s.ud_density_absorb_bot;
        else if (e.kind == 3 && e.plane >= 0 && e.plane < RB2_PLANES_MAX) f = s.planes_ud[e.plane];
        if (!f) continue;
        if (e.kind == 2) {  // src/mod_pair.F90:255-256: x, y, emit, sec, id
            const double xy[2] = {e.x, e.y};
            const int t[3] = {e.emit, e.sec, e.id};
            fwrite(xy, sizeof(double), 2, f);
            fwrite(t, sizeof(int), 3, f);
        } else {  // src/mod_pair.F90:243-245, src/mod_verlet.F90:360-362: x, y, vx, vy, vz, emit, sec, id
            const double v[5] = {e.x, e.y, e.vx, e.vy, e.vz};
            const int t[3] = {e.emit, e.sec, e.id};
            fwrite(v, sizeof(double), 5, f);
            fwrite(t, sizeof(int), 3, f);
        }
    }
}

// Write_Position, src/mod_pair.F90:841-867: stream of (steps, N_steps) once, then per step (step, nrPart) and per
// particle x, y, z [m], emitter, section, id.
static int write_position(Sim &s, int step)
{
    if (!s.ud_pos) return 0;
    const int N_steps = 1;
    if (step == 1) { const int h[2] = {s.g.steps, N_steps}; fwrite(h, sizeof(int), 2, s.ud_pos); }
    const int n = s.counts.nrPart;
    const int h[2] = {step, n};
    fwrite(h, sizeof(int), 2, s.ud_pos);
    if (n < 1) return 0;
    std::vector<double> pos((size_t)3 * n);
    std::vector<int> emit((size_t)n), sec((size_t)n), id((size_t)n);
    if (s.check(rb2_download_particles(pos.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                       emit.data(), sec.data(), nullptr, id.data(), nullptr), "rb2_download_particles")) return -1;
    for (int i = 0; i < n; ++i) {
        fwrite(&pos[(size_t)3 * i], sizeof(double), 3, s.ud_pos);
        const int t[3] = {emit[i], sec[i], id[i]};
        fwrite(t, sizeof(int), 3, s.ud_pos);
    }
    return 0;
}

// Sample_Elec_Position, src/mod_pair.F90:975-1037: every sample_elec_rate steps the nearest-neighbour sweep (on the
// device, rb2_nearest_electron) and out/elec-<step>.bin with x, y, z, nearest distance [m] of every electron.
static int sample_elec_position(Sim &s, int step)
{
    if (!s.g.sample_elec_file || !s.write_files) return 0;
    if (step % s.g.sample_elec_rate != 0) return 0;
    const int n = s.counts.nrPart;
    std::vector<double> pos((size_t)3 * (n > 0 ? n : 1)), dist((size_t)(n > 0 ? n : 1));
    std::vector<int> species((size_t)(n > 0 ? n : 1));
    if (n > 0) {
        if (s.check(rb2_nearest_electron(dist.data(), nullptr), "rb2_nearest_electron")) return -1;
        if (s.check(rb2_download_particles(pos.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, species.data(),
                                           nullptr, nullptr, nullptr, nullptr, nullptr, nullptr), "rb2_download_particles")) return -1;
    }
    char nm[64];
    snprintf(nm, sizeof(nm), "elec-%d.bin", step);
    FILE *f = open_out(s, nm, "wb");
    if (!f) { fprintf(stderr, "RUMDEED: Failed to open the electron position file.\n"); return 0; }
    for (int i = 0; i < n; ++i) {
        if (species[i] != species_elec) continue;
        const double rec[4] = {pos[(size_t)3 * i], pos[(size_t)3 * i + 1], pos[(size_t)3 * i + 2], dist[i]};
        fwrite(rec, sizeof(double), 4, f);
    }
    fclose(f);
    return 0;
}

// One iteration of the main loop, src/main.F90:175-219
int Step(Sim &s, int step)
{
    Globals &g = s.g;
    s.cur_step = step;
    using clk = std::chrono::steady_clock;
    auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const auto t0 = clk::now();
    // ptr_Do_Emission(i)
    if (s.ptr.ptr_Do_Emission(s, step)) return -1;
    const auto t1 = clk::now();
    s.t_emission += secs(t0, t1);
    s.nrEmitted_total += s.slog.nrElecEmit;
    s.cur_time = g.time_step * step / time_scale;
    if (s.ud_emit)  // src/mod_field_emission_v2.F90:203: "(E14.6, *(tr8, i6))"
        fprintf(s.ud_emit, "%14.6E        %6d        %6d        %6d        %6d\n", s.cur_time, step, s.slog.nrElecEmit, s.counts.nrElec, s.slog.nrElecEmit);
    // Update_Position(i): Set_Voltage + Beeman step on the device
    g.V_d = g.V_s;
    if (s.ud_volt) fprintf(s.ud_volt, "%12.4E  %8d  %18.8E  %18.8E\n", s.cur_time, step, g.V_d, 0.0);
    const auto t2 = clk::now();
    if (s.check(rb2_step(step, &s.last), "rb2_step")) return -1;
    const auto t3 = clk::now();
    s.t_md_step += secs(t2, t3);
    s.t_dev_step += 1e-3 * s.last.step_ms;
    s.t_dev_accel += 1e-3 * s.last.accel_ms;
    s.counts = s.last.counts;
    double ramo_cur = 0.0;
    for (int k = 1; k <= 3; ++k) ramo_cur += s.last.ramo_current[k];
    ramo_cur /= cur_scale;
    s.ramo_integral += ramo_cur * g.time_step;
    if (s.ud_ramo) {  // Write_Ramo_Current, src/mod_pair.F90:812-837
        auto nrm = [](const double v[3]) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
        const double E_z = -1.0 * g.V_d / g.d;
        const double avg_mob = (E_z != 0.0) ? nrm(s.last.avg_elec_vel) / (-1.0 * E_z) : 0.0;  // Average_Velocities, src/mod_verlet.F90:441-445
        fprintf(s.ud_ramo, "%12.4E  %8d  %12.4E  %12.4E  %6d  %6d  %6d  %12.4E  %12.4E  %12.4E  %12.4E  %12.4E  %12.4E  %12.4E\n", s.cur_time, step,
                ramo_cur, g.V_d, s.counts.nrPart, s.counts.nrElec, s.counts.nrIon, avg_mob, nrm(s.last.avg_part_vel),
                nrm(s.last.avg_elec_vel), nrm(s.last.avg_ion_vel), s.last.ramo_current[1], s.last.ramo_current[2], s.last.ramo_current[3]);
    }
    if (s.ramo_sections > 0) {  // Write_Ramo_Current, src/mod_pair.F90:822-826: the whole (MAX_SECTIONS, MAX_EMITTERS) array
        if (s.check(rb2_get_ramo_sections(MAX_SECTIONS, MAX_EMITTERS, s.ramo_current_emit.data()), "rb2_get_ramo_sections")) return -1;
        if (s.ud_ramo_sec) fwrite(s.ramo_current_emit.data(), sizeof(double), s.ramo_current_emit.size(), s.ud_ramo_sec);
    }
    if (s.write_files) write_event_files(s);
    if (write_position(s, step)) return -1;        // src/main.F90:191
    if (sample_elec_position(s, step)) return -1;  // src/main.F90:193
    // Remove_Particles(i): Write_Absorbed (src/mod_pair.F90:790-805) + compaction
    s.nrAbsorbed_top += s.counts.nrElec_remove_top;
    s.nrAbsorbed_bot += s.counts.nrElec_remove_bot;
    if (s.ud_absorb) fprintf(s.ud_absorb, "%12.4E  %8d  %8d  %8d  %8d\n", s.cur_time, step, s.counts.nrPart_remove, s.counts.nrElec_remove, s.counts.nrIon_remove);
    if (s.ud_absorb_top) fprintf(s.ud_absorb_top, "%12.4E  %8d  %8d  %8d  %8d  %8d\n", s.cur_time, step, s.counts.nrPart_remove_top, s.counts.nrElec_remove_top, s.counts.nrIon_remove_top, s.counts.nrElec_remove_top);
    if (s.ud_absorb_bot) fprintf(s.ud_absorb_bot, "%12.4E  %8d  %8d  %8d  %8d\n", s.cur_time, step, s.counts.nrPart_remove_bot, s.counts.nrElec_remove_bot, s.counts.nrIon_remove_bot);
    if (s.ud_absorb_recom) fprintf(s.ud_absorb_recom, "%12.4E  %8d  %8d  %8d  %8d\n", s.cur_time, step, s.recom_counts[0], s.recom_counts[1], s.recom_counts[2]);
    s.recom_counts[0] = s.recom_counts[1] = s.recom_counts[2] = 0;
    rb2_counts k{};
    const auto t4 = clk::now();
    if (s.check(rb2_remove_marked(step, &k), "rb2_remove_marked")) return -1;
    const auto t5 = clk::now();
    s.t_remove += secs(t4, t5);
    s.t_io += secs(t1, t2) + secs(t3, t4);
    s.counts = k;
    // Do_Collisions(i), src/main.F90:202
    if (Do_Collisions(s, step)) return -1;
    s.t_collisions += secs(t5, clk::now());
    return 0;
}

// src/main.F90:233-266
int Clean_up(Sim &s)
{
    if (s.ptr.ptr_Clean_Up) s.ptr.ptr_Clean_Up(s);
    FILE **fs[] = {&s.ud_ramo, &s.ud_emit, &s.ud_absorb, &s.ud_absorb_top, &s.ud_absorb_bot, &s.ud_field, &s.ud_integrand, &s.ud_volt,
                   &s.ud_density_emit, &s.ud_density_emit_elec, &s.ud_density_emit_ion, &s.ud_density_emit_atom, &s.ud_ramo_sec,
                   &s.ud_density_absorb_top, &s.ud_density_absorb_bot, &s.ud_pos, &s.ud_coll, &s.ud_ionization_data,
                   &s.ud_recombination_data, &s.ud_density_absorb_recom, &s.ud_absorb_recom};
    for (FILE **f : fs) if (*f) { fclose(*f); *f = nullptr; }
    for (int k = 0; k < RB2_PLANES_MAX; ++k) if (s.planes_ud[k]) { fclose(s.planes_ud[k]); s.planes_ud[k] = nullptr; }
    if (s.write_files) {  // Write_Life_Time, src/mod_pair.F90:776-786
        std::vector<long long> lt((size_t)(RB2_MAX_LIFE_TIME + 1) * 4);
        if (rb2_get_life_time(lt.data()) == RB2_OK) {
            if (FILE *f = open_out(s, "lifetime.dt", "w")) {
                for (int i = 1; i <= RB2_MAX_LIFE_TIME; ++i)
                    fprintf(f, "%12.4E  %6d  %6lld  %6lld\n", i * s.g.time_step / time_scale, i, lt[(size_t)i * 4 + 1], lt[(size_t)i * 4 + 2]);
                fclose(f);
            }
        }
    }
    rb2_finalize();
    return 0;
}

}  // namespace rh

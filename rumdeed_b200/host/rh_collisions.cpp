// rh_collisions.cpp -- host side of the electron / N2 collision step (collision_mode 1 and 2): cross-section
// files, the call into the device path and the three output files.  Mirrors
//   Read_Cross_Section            src/mod_collisions.F90:1909-1983   (N2-tot-cross.txt, N2-ion-cross.txt in the run dir)
//   Do_Collisions                 src/mod_verlet.F90:164-170  ->  Do_Electron_Atom_Collisions, src/mod_collisions.F90:30-76
//   Write_Ionization_Data         src/mod_pair.F90:930-935    (out/ionization_data.bin, stream)
//   Write_Recombination_Data      src/mod_pair.F90:919-926    (out/recombination_data.bin, stream)
//   Mark_Particles_Remove(.., remove_recom)   src/mod_pair.F90:259-272, :302-315   (out/density_absorb_recom.bin)
// The collision arithmetic itself runs on the device (rb2_do_collisions).
#include <math.h>

#include <fstream>
#include <sstream>

#include "rh_host.hpp"

namespace rh {

static int read_two_columns(const std::string &path, std::vector<double> &a, std::vector<double> &b, std::string &err)
{
    std::ifstream f(path);
    if (!f) { err = "RUMDEED: ERROR UNABLE TO OPEN file " + path; return -1; }
    std::string line;
    while (std::getline(f, line)) {
        std::stringstream ss(line);
        double x, y;
        if (ss >> x >> y) { a.push_back(x); b.push_back(y); }
    }
    if (a.size() < 2) { err = "RUMDEED: cross-section file " + path + " holds fewer than two rows"; return -1; }
    return 0;
}

int Init_Collisions(Sim &s)
{
    Globals &g = s.g;
    if (g.collision_mode == 0) return 0;
    if (g.collision_mode != 1 && g.collision_mode != 2)
        return s.fail("RUMDEED: COLLISION_MODE " + std::to_string(g.collision_mode) +
                      " (discrete ionisation with N2 atoms as particles) is not on the device path; modes 1 and 2 are");
    std::vector<double> te, td, ie, id;
    if (read_two_columns(s.dir + "/N2-tot-cross.txt", te, td, s.err)) return -1;
    if (read_two_columns(s.dir + "/N2-ion-cross.txt", ie, id, s.err)) return -1;
    rb2_collision_config c{};
    c.collision_mode = g.collision_mode;
    c.ion_life_time = g.ion_life_time;
    c.n_d = g.P_abs / (k_b * g.T_temp);  // src/main.F90:382-383 (P_abs already scaled by P_ntp)
    c.cyl_radius = g.emitters_dim[0];    // emitters_dim(1,1), src/mod_collisions.F90:594
    c.n_tot = (int)te.size(); c.n_ion = (int)ie.size();
    c.tot_energy = te.data(); c.tot_data = td.data(); c.ion_energy = ie.data(); c.ion_data = id.data();
    if (s.check(rb2_collisions_init(&c), "rb2_collisions_init")) return -1;
    if (s.write_files) {
        auto open = [&](const char *nm, const char *mode) { return fopen((s.out_dir + "/" + nm).c_str(), mode); };
        s.ud_coll = open("collisions.dt", "w");
        s.ud_ionization_data = open("ionization_data.bin", "wb");
        s.ud_recombination_data = open("recombination_data.bin", "wb");
        s.ud_density_absorb_recom = open("density_absorb_recom.bin", "wb");
        s.ud_absorb_recom = open("absorbed_recom.dt", "w");
    }
    return 0;
}

int Do_Collisions(Sim &s, int step)
{
    Globals &g = s.g;
    if (g.collision_mode == 0) return 0;
    rb2_collision_result r{};
    if (step >= g.collision_delay) {
        if (s.check(rb2_do_collisions(step, s.rng.next(), &r), "rb2_do_collisions")) return -1;
        s.counts = r.counts;
        s.t_dev_collisions += 1e-3 * r.ms;
        s.nrIonizations_total += r.nrIonizations;
        s.nrRecombinations_total += r.nrRecombinations;
        s.recom_counts[0] = r.nrPart_remove_recom; s.recom_counts[1] = r.nrElec_remove_recom; s.recom_counts[2] = r.nrIon_remove_recom;
        if (s.write_files && r.nrIonizations > 0 && s.ud_ionization_data) {
            std::vector<rb2_ionization_record> ev((size_t)r.nrIonizations);
            int n = 0;
            if (s.check(rb2_get_ionization_records(r.nrIonizations, ev.data(), &n), "rb2_get_ionization_records")) return -1;
            for (int k = 0; k < n && k < r.nrIonizations; ++k) {
                const rb2_ionization_record &e = ev[(size_t)k];
                // the two Add_Particle calls of the event (ejected electron, then ion; emitter = ion_emitter = 2, default
                // section 1, src/mod_collisions.F90:668-685) write their density_emit*.bin records (src/mod_pair.F90:85-123)
                for (int which = 0; which < 2; ++which) {
                    const int pid = which == 0 ? e.new_id : e.ion_id;
                    if (pid < 0) continue;  // dropped at MAX_PARTICLES
                    const double *pp = which == 0 ? e.ejec_pos : e.ion_pos;
                    const double p3[3] = {pp[0] / length_scale, pp[1] / length_scale, pp[2] / length_scale};
                    const int sp = which == 0 ? species_elec : species_ion;
                    FILE *fs = which == 0 ? s.ud_density_emit_elec : s.ud_density_emit_ion;
                    if (fs) { const int t2[2] = {2, pid}; fwrite(p3, sizeof(double), 3, fs); fwrite(t2, sizeof(int), 2, fs); }
                    if (s.ud_density_emit) { const int t4[4] = {2, 1, pid, sp}; fwrite(p3, sizeof(double), 3, s.ud_density_emit); fwrite(t4, sizeof(int), 4, s.ud_density_emit); }
                }
                const double d[8] = {e.pos[0], e.pos[1], e.pos[2], e.in_speed, e.out_speed, e.new_speed, 0.0, 0.0};
                const int t[4] = {e.in_slot + 1, e.new_id, e.ion_id, e.elec_emit};  // inID is the Fortran slot
                fwrite(&e.step, sizeof(int), 1, s.ud_ionization_data);
                fwrite(d, sizeof(double), 8, s.ud_ionization_data);
                fwrite(t, sizeof(int), 4, s.ud_ionization_data);
            }
        }
        if (s.write_files && r.nrRecombinations > 0) {
            std::vector<rb2_recomb_record> ev((size_t)r.nrRecombinations);
            int n = 0;
            if (s.check(rb2_get_recombination_records(r.nrRecombinations, ev.data(), &n), "rb2_get_recombination_records")) return -1;
            const double t_now = s.cur_time / g.time_step * time_scale;  // cur_time/time_step*time_scale, :271
            for (int k = 0; k < n && k < r.nrRecombinations; ++k) {
                const rb2_recomb_record &e = ev[(size_t)k];
                if (s.ud_density_absorb_recom) {  // Mark(ion) then Mark(electron), src/mod_collisions.F90:213-216
                    const int ti[4] = {e.ion_emit, e.ion_sec, e.ion_id, species_ion};
                    fwrite(e.ion_pos, sizeof(double), 3, s.ud_density_absorb_recom);
                    fwrite(ti, sizeof(int), 4, s.ud_density_absorb_recom);
                    fwrite(&t_now, sizeof(double), 1, s.ud_density_absorb_recom);
                    const int te[4] = {e.elec_emit, e.elec_sec, e.elec_id, species_elec};
                    fwrite(e.elec_pos, sizeof(double), 3, s.ud_density_absorb_recom);
                    fwrite(te, sizeof(int), 4, s.ud_density_absorb_recom);
                    fwrite(&t_now, sizeof(double), 1, s.ud_density_absorb_recom);
                }
                if (s.ud_recombination_data) {
                    const double d[6] = {e.ion_pos[0], e.ion_pos[1], e.ion_pos[2], e.elec_speed, e.dist, e.recom_rad};
                    const int t[4] = {e.elec_slot + 1, e.ion_slot + 1, e.elec_emit, e.ion_life};
                    fwrite(&e.step, sizeof(int), 1, s.ud_recombination_data);
                    fwrite(d, sizeof(double), 6, s.ud_recombination_data);
                    fwrite(t, sizeof(int), 4, s.ud_recombination_data);
                }
            }
        }
    }
    if (s.ud_coll)  // '(i6,tr2,i6,tr2,i6,tr2,i6)', src/mod_collisions.F90:74-75
        fprintf(s.ud_coll, "%6d  %6d  %6d  %6d\n", step, r.nrCollisions, r.nrIonizations, r.nrRecombinations);
    return 0;
}

}  // namespace rh

// rh_input.cpp -- random numbers, the /input/ namelist reader, and the work / laser files.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <map>
#include <sstream>

#include "rh_host.hpp"

namespace rh {

// ---- xoshiro256++ ---------------------------------------------------------------------------------
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
void Rng::seed(uint64_t v)
{
    for (int i = 0; i < 4; ++i) {  // splitmix64
        uint64_t z = (v += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        s[i] = z ^ (z >> 31);
    }
}
uint64_t Rng::next()
{
    const uint64_t result = rotl(s[0] + s[3], 23) + s[0];
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
// Marsaglia polar method, src/mod_global.F90:578-595
void Rng::box_muller(const double mean[2], const double std[2], double out[2])
{
    double x0, x1, w;
    do {
        x0 = 2.0 * uniform() - 1.0;
        x1 = 2.0 * uniform() - 1.0;
        w = x0 * x0 + x1 * x1;
    } while (!((w < 1.0) && (w > 0.0)));
    const double f = sqrt((-2.0 * log(w)) / w);
    out[0] = x0 * f * std[0] + mean[0];
    out[1] = x1 * f * std[1] + mean[1];
}
// src/mod_global.F90:600-643
int Rng::poisson(double lambda)
{
    const double Poisson_Step = 500.0;
    double lambda_left = lambda, p = 1.0;
    int k = 0;
    do {
        k = k + 1;
        p = p * uniform();
        while ((p < 1.0) && (lambda_left > 0.0)) {
            if (lambda_left > Poisson_Step) { p = p * exp(Poisson_Step); lambda_left -= Poisson_Step; }
            else { p = p * exp(lambda_left); lambda_left = 0.0; }
        }
    } while (p >= 1.0);
    return k - 1;
}

// ---- namelist ------------------------------------------------------------------------------------------
static std::string upper(std::string s)
{
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)toupper(c); });
    return s;
}
static std::string trim(const std::string &s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
static bool parse_double(std::string t, double &v)
{
    t = trim(t);
    if (t.empty()) return false;
    for (char &c : t) if (c == 'd' || c == 'D') c = 'e';  // Fortran double exponent
    char *end = nullptr;
    v = strtod(t.c_str(), &end);
    return end && *end == '\0';
}
static bool parse_logical(std::string t, bool &v)
{
    t = upper(trim(t));
    if (t == ".TRUE." || t == "T" || t == ".T.") { v = true; return true; }
    if (t == ".FALSE." || t == "F" || t == ".F.") { v = false; return true; }
    return false;
}

// Reads the &INPUT ... / group (src/mod_global.F90:426-439) and applies the unit scaling of
// Read_Input_Variables (src/main.F90:348-383): nm -> m, ps -> s.
int Read_Input_Variables(const std::string &path, Globals &g, std::string &err)
{
    std::ifstream f(path);
    if (!f) { err = "RUMDEED: ERROR UNABLE TO OPEN file input (" + path + ")"; return -1; }
    std::map<std::string, std::vector<std::string>> kv;
    std::string line, key;
    bool in_group = false;
    while (std::getline(f, line)) {
        size_t c = line.find('!');
        if (c != std::string::npos) line = line.substr(0, c);
        line = trim(line);
        if (line.empty()) continue;
        if (!in_group) {
            if (upper(line).rfind("&INPUT", 0) == 0) { in_group = true; line = trim(line.substr(6)); if (line.empty()) continue; }
            else continue;
        }
        if (line == "/" || upper(line) == "&END") break;
        size_t eq = line.find('=');
        std::string vals = line;
        if (eq != std::string::npos) {
            key = upper(trim(line.substr(0, eq)));
            size_t par = key.find('(');
            if (par != std::string::npos) key = trim(key.substr(0, par));
            vals = line.substr(eq + 1);
            kv[key];
        }
        if (key.empty()) continue;
        std::stringstream ss(vals);
        std::string tok;
        while (std::getline(ss, tok, ',')) {
            tok = trim(tok);
            if (tok == "/") { in_group = false; break; }
            if (!tok.empty()) kv[key].push_back(tok);
        }
    }
    auto dbl = [&](const char *k, double &v) { auto it = kv.find(k); if (it != kv.end() && !it->second.empty()) { double t; if (parse_double(it->second[0], t)) v = t; } };
    auto integer = [&](const char *k, int &v) { auto it = kv.find(k); if (it != kv.end() && !it->second.empty()) { double t; if (parse_double(it->second[0], t)) v = (int)llround(t); } };
    auto logical = [&](const char *k, bool &v) { auto it = kv.find(k); if (it != kv.end() && !it->second.empty()) { bool t; if (parse_logical(it->second[0], t)) v = t; } };
    auto vec = [&](const char *k, double *v, int n) { auto it = kv.find(k); if (it != kv.end()) for (int i = 0; i < n && i < (int)it->second.size(); ++i) { double t; if (parse_double(it->second[i], t)) v[i] = t; } };
    dbl("V_S", g.V_s);
    vec("BOX_DIM", g.box_dim, 3);
    dbl("TIME_STEP", g.time_step);
    integer("STEPS", g.steps);
    integer("NREMIT", g.nrEmit);
    integer("EMISSION_MODE", g.emission_mode);
    logical("IMAGE_CHARGE", g.image_charge);
    integer("N_IC_MAX", g.N_ic_max);
    integer("COLLISION_MODE", g.collision_mode);
    integer("COLLISION_DELAY", g.collision_delay);
    integer("ION_LIFE_TIME", g.ion_life_time);
    dbl("T_TEMP", g.T_temp);
    dbl("P_ABS", g.P_abs);
    vec("EMITTERS_DIM", g.emitters_dim, 3);
    vec("EMITTERS_POS", g.emitters_pos, 3);
    integer("EMITTERS_TYPE", g.emitters_type);
    integer("EMITTERS_DELAY", g.emitters_delay);
    integer("PLANES_N", g.planes_N);
    vec("PLANES_Z", g.planes_z, RB2_PLANES_MAX);
    logical("MH_BATCH", g.mh_batch);
    logical("MH_DEVICE", g.mh_device);  // extension: lock-step chains resident on the GPU (rb2_mh_planar)
    if (g.mh_device) g.mh_batch = true;
    logical("MH_HOST", g.mh_host);      // extension: run the default serial chains as a host loop (one field call per jump)
    logical("WRITE_RAMO_SEC", g.write_ramo_sec);
    logical("WRITE_POSITION_FILE", g.write_position_file);
    logical("SAMPLE_ELEC_FILE", g.sample_elec_file);
    integer("SAMPLE_ELEC_RATE", g.sample_elec_rate);
    integer("CUBA_METHOD", g.cuba_method);
    dbl("CUBA_EPSABS", g.cuba_epsabs);
    dbl("CUBA_EPSREL", g.cuba_epsrel);
    integer("CUBA_MINEVAL", g.cuba_mineval);
    integer("CUBA_MAXEVAL", g.cuba_maxeval);
    // extensions of this host (not in the reference namelist)
    integer("MAX_PARTICLES", g.max_particles);
    { double sd = 0.0; dbl("SEED", sd); if (sd > 0.0) g.seed = (uint64_t)sd; }

    if (g.planes_N > RB2_PLANES_MAX) { err = "RUMDEED: ERROR planes_N exceeds planes_N_max"; return -1; }
    if (g.nrEmit > MAX_EMITTERS) { err = "RUMDEED: ERROR nrEmit exceeds MAX_EMITTERS"; return -1; }
    for (int k = 0; k < 3; ++k) g.box_dim[k] *= length_scale;
    g.d = g.box_dim[2];
    g.V_d = g.V_s;
    g.E_z = -1.0 * g.V_d / g.d;
    for (int k = 0; k < 3; ++k) { g.emitters_dim[k] *= length_scale; g.emitters_pos[k] *= length_scale; }
    for (int k = 0; k < RB2_PLANES_MAX; ++k) g.planes_z[k] *= length_scale;
    g.time_step *= time_scale;
    g.time_step2 = g.time_step * g.time_step;
    g.P_abs *= P_ntp;
    if (g.steps <= 0) { err = "ERROR: steps <= 0"; return -1; }
    if (g.sample_elec_rate < 1) g.sample_elec_rate = 1;
    return 0;
}

// ---- work function file, src/mod_work_function.F90:53-160 ------------------------------------------
int WorkFunction::read(const std::string &path, std::string &err)
{
    std::ifstream f(path);
    if (!f) { err = "RUMDEED: Failed to open file work. ABORTING (" + path + ")"; return -1; }
    if (!(f >> type)) { err = "work: cannot read WORK_TYPE"; return -1; }
    if (type != 1) { err = "RUMDEED: work function type " + std::to_string(type) + " is not on the device path (checkerboard only)"; return -1; }
    if (!(f >> y_num >> x_num) || y_num < 1 || x_num < 1) { err = "work: bad matrix size"; return -1; }
    w_theta_arr.assign((size_t)y_num * x_num, 0.0);
    for (int i = 0; i < y_num; ++i)
        for (int j = 0; j < x_num; ++j)
            if (!(f >> w_theta_arr[(size_t)i * x_num + j])) { err = "work: short matrix"; return -1; }
    return 0;
}

// w_theta_checkerboard, src/mod_work_function.F90:389-487 (nrEmit == 1 branch)
double WorkFunction::w_theta_xy(const Globals &g, const double pos[3], int *sec) const
{
    const double x_len = 1.0 / x_num, y_len = 1.0 / y_num;
    const double x = (pos[0] - g.emitters_pos[0]) / g.emitters_dim[0];
    const double y = (pos[1] - g.emitters_pos[1]) / g.emitters_dim[1];
    int x_i = (int)floor(x / x_len) + 1, y_i = (int)floor(y / y_len) + 1;
    if (x_i > x_num) x_i = x_num; else if (x_i < 1) x_i = 1;
    if (y_i > y_num) y_i = y_num; else if (y_i < 1) y_i = 1;
    if (sec) *sec = x_num * (y_i - 1) + x_i;
    y_i = y_num - y_i + 1;  // reverse the y-direction in the array
    return w_theta_arr[(size_t)(y_i - 1) * x_num + (x_i - 1)];
}

// ---- laser file, src/mod_photo_emission.f90:56-147 --------------------------------------------------
int Laser::read(const std::string &path, std::string &err)
{
    std::ifstream f(path);
    if (!f) { err = "RUMDEED: Failed to open file laser. ABORTING (" + path + ")"; return -1; }
    if (!(f >> gauss_mode >> laser_mode >> photon_mode)) { err = "laser: bad first line"; return -1; }
    if (gauss_mode != 1 && gauss_mode != 2) { err = "RUMDEED: ERROR UNKNOWN LASER TYPE"; return -1; }
    if (laser_mode != 1 && laser_mode != 2) { err = "RUMDEED: ERROR UNKNOWN LASER MODE"; return -1; }
    if (photon_mode != 1 && photon_mode != 2) { err = "RUMDEED: ERROR UNKNOWN LASER MODE"; return -1; }
    if (!(f >> laser_energy >> laser_variation)) { err = "laser: bad energy line"; return -1; }
    if (gauss_mode == 1 && !(f >> gauss_center >> gauss_width >> gauss_amplitude)) { err = "laser: bad pulse line"; return -1; }
    return 0;
}

}  // namespace rh

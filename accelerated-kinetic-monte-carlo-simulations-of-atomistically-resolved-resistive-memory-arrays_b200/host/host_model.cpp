// host_model.cpp -- host-side model pieces that feed the hot path and must read the reference's input files
// unchanged (no GPU involved).  Exposed through the C ABI (include/kmc_b200.h, "Host model").
//
//  * parameters.txt grammar: reference src/input_parser.cpp:3-249 (key = substring match on "<key> ", comment
//    lines start with "//", trailing "//" comments stripped) and the read_* helpers :261-373
//    (read_bool: line contains '1'; read_int: token after '='; read_double: LAST numeric token, 0 rejected;
//    read_string: last token; read_vec_double: every token that parses as a double prefix;
//    read_vec_string: tokens after '='), derived values :391-398.
//  * xyz files: reference src/utils.cpp:72-98 (line 1 = N, line 2 skipped, then "El x y z ..."),
//    element names src/utils.cpp:7-29.
//  * Device::makeSubstoichiometric: reference src/Device.cpp:180-211 (std::mt19937(rnd_seed),
//    uniform_real_distribution<double>, atoms = sites that are not d/Od, :116-145).
//  * KMCProcess layers: reference src/structure_input.h:8-50, src/KMCProcess.cpp:34-50.
//  * KMC_comm row partition: reference src/KMC_comm.h:245-263.
// Validated against the reference's own compiled parser/reader (oracle/_ref) in tests/test_host_model.py.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/kmc_b200.h"

void kmc_set_error(const char *fmt, ...);

namespace {

int element_from_name(const std::string &s) {
    if (s == "d") return KMCB200_DEFECT;
    if (s == "Od") return KMCB200_OXYGEN_DEFECT;
    if (s == "V") return KMCB200_VACANCY;
    if (s == "O") return KMCB200_O;
    if (s == "Hf") return KMCB200_Hf;
    if (s == "N") return KMCB200_N;
    if (s == "Ti") return KMCB200_Ti;
    if (s == "Pt") return KMCB200_Pt;
    return -1;  // the reference exits on unknown names (utils.cpp:25-29); note "Ni" is not accepted there either
}

std::string strip_comment(const std::string &line) {
    size_t pos = line.find("//");
    return pos == std::string::npos ? line : line.substr(0, pos);
}
bool has_key(const std::string &line, const char *key) { return line.find(key) != std::string::npos; }

bool parse_bool(const std::string &line) {
    if (line.find("1") != std::string::npos) return true;
    if (line.find("0") != std::string::npos) return false;
    throw std::invalid_argument("Invalid input to read_bool: " + line);
}
int parse_int(const std::string &line) {
    std::istringstream stream(line);
    std::string token;
    while (stream >> token) {
        if (token == "=") {
            if (stream >> token) {
                int value;
                if (std::istringstream(token) >> value) return value;
                throw std::invalid_argument("Invalid integer after equal sign: " + token);
            }
        }
    }
    throw std::invalid_argument("Equal sign and integer not found in input: " + line);
}
double parse_double(const std::string &line) {
    std::istringstream stream(line);
    double value = 0.0;
    std::string token;
    while (stream >> token) {
        double tmp;
        if (std::istringstream(token) >> tmp) {
            value = tmp;
        } else if (token.find_first_not_of("0123456789.eE-") == std::string::npos) {
            value = std::stod(token);
        }
    }
    if (value != 0.0) return value;
    throw std::invalid_argument("No double value found in input: " + line);
}
std::string parse_string(const std::string &line) {
    std::istringstream stream(line);
    std::string last, word;
    while (stream >> word) last = word;
    return last;
}
std::vector<double> parse_vec_double(const std::string &line) {
    std::vector<double> out;
    std::istringstream stream(line);
    std::string token;
    while (stream >> token) {
        double v;
        if (std::istringstream(token) >> v) out.push_back(v);
    }
    return out;
}
std::vector<std::string> parse_vec_string(const std::string &line) {
    std::vector<std::string> out;
    std::istringstream stream(line);
    std::string token;
    bool eq = false;
    while (stream >> token) {
        if (eq) out.push_back(token);
        if (token == "=") eq = true;
    }
    return out;
}

struct Parsed {
    kmcb200_params p;
    std::vector<double> V_switch, t_switch;
    double G_coeff = 1;
};

void parse_file(const char *path, Parsed &P) {
    std::ifstream in(path);
    if (!in.is_open()) throw std::runtime_error(std::string("cannot open ") + path);
    kmcb200_params &p = P.p;
    std::memset(&p, 0, sizeof(p));
    std::string raw;
    auto set_str = [](char *dst, const std::string &s) { std::strncpy(dst, s.c_str(), 511); dst[511] = 0; };
    while (std::getline(in, raw)) {
        if (raw.substr(0, 2) == "//") continue;
        std::string line = strip_comment(raw);
        if (has_key(line, "rnd_seed ")) p.rnd_seed = (unsigned)parse_int(line);
        if (has_key(line, "restart ")) p.restart = parse_bool(line);
        if (has_key(line, "restart_xyz_file ")) set_str(p.restart_xyz_file, parse_string(line));
        if (has_key(line, "log_freq ")) p.log_freq = parse_int(line);
        if (has_key(line, "output_freq ")) p.output_freq = parse_int(line);
        if (has_key(line, "atom_xyz_file ")) set_str(p.atom_xyz_file, parse_string(line));
        if (has_key(line, "interstitial_xyz_file ")) set_str(p.interstitial_xyz_file, parse_string(line));
        if (has_key(line, "pristine ")) p.pristine = parse_bool(line);
        if (has_key(line, "shift ")) p.shift = parse_bool(line);
        if (has_key(line, "pbc ")) p.pbc = parse_bool(line);
        if (has_key(line, "num_atoms_first_layer ")) p.num_atoms_first_layer = parse_int(line);
        if (has_key(line, "num_layers_contact ")) p.num_layers_contact = parse_int(line);
        if (has_key(line, "num_atoms_contact ")) p.num_atoms_contact = parse_int(line);
        if (has_key(line, "num_atoms_reservoir ")) p.num_atoms_reservoir = parse_int(line);
        if (has_key(line, "initial_vacancy_concentration ")) p.initial_vacancy_concentration = parse_double(line);
        if (has_key(line, "nn_dist ")) p.nn_dist = parse_double(line);
        if (has_key(line, "attempt_frequency ")) p.freq = parse_double(line);
        if (has_key(line, "shifts ")) {
            auto v = parse_vec_double(line);
            p.n_shifts = (int)v.size();
            for (size_t i = 0; i < v.size() && i < 3; ++i) p.shifts[i] = v[i];
        }
        if (has_key(line, "lattice ")) {
            auto v = parse_vec_double(line);
            p.n_lattice = (int)v.size();
            for (size_t i = 0; i < v.size() && i < 3; ++i) p.lattice[i] = v[i];
        }
        if (has_key(line, "metals ")) {
            for (auto &e : parse_vec_string(line)) {
                int id = element_from_name(e);
                if (id < 0) throw std::invalid_argument("Unknown element type: " + e);
                if (p.num_metals < 8) p.metals[p.num_metals++] = id;
            }
        }
        if (has_key(line, "solve_potential ")) p.solve_potential = parse_bool(line);
        if (has_key(line, "solve_current ")) p.solve_current = parse_bool(line);
        if (has_key(line, "solve_heating_global ")) p.solve_heating_global = parse_bool(line);
        if (has_key(line, "solve_heating_local ")) p.solve_heating_local = parse_bool(line);
        if (has_key(line, "perturb_structure ")) p.perturb_structure = parse_bool(line);
        if (has_key(line, "V_switch ")) P.V_switch = parse_vec_double(line);
        if (has_key(line, "t_switch ")) P.t_switch = parse_vec_double(line);
        if (has_key(line, "Icc ")) p.Icc = parse_double(line);
        if (has_key(line, "Rs ")) p.Rs = parse_double(line);
        if (has_key(line, "sigma ")) p.sigma = parse_double(line);
        if (has_key(line, "epsilon ")) p.epsilon = parse_double(line);
        if (has_key(line, "m_r ")) p.m_r = parse_double(line);
        if (has_key(line, "V0 ")) p.V0 = parse_double(line);
        if (has_key(line, "background_temp ")) p.background_temp = parse_double(line);
        if (has_key(line, "t_ox ")) p.t_ox = parse_double(line);
        if (has_key(line, "A ")) {
            p.A = 1;
            for (double d : parse_vec_double(line)) p.A *= d;
        }
    }
    p.n_V_switch = (int)P.V_switch.size();
    p.n_t_switch = (int)P.t_switch.size();
    p.V_switch0 = P.V_switch.empty() ? 0 : P.V_switch[0];
    p.t_switch0 = P.t_switch.empty() ? 0 : P.t_switch[0];
    // set_expression_parameters (input_parser.cpp:391-398)
    p.high_G = P.G_coeff * 1;
    p.low_G = P.G_coeff * 1e-8;
    p.k = 8.987552e9 / p.epsilon;
}

// reference src/structure_input.h:10-50
struct LayerDef { double E_gen, E_rec, E_Vdiff, E_Odiff, start_x, end_x; };
const LayerDef kLayers[5] = {
    {0.0, 0.0, 0.0, 0.76, -22.0, 0.0},            // contact
    {3.93, 0.0, 1.09, 0.76, 0.0, 3.0},            // interface
    {3.93, 0.0, 1.09, 0.76, 3.0, 48.1431},        // oxide
    {1.66, 0.0, 1.09, 0.76, 48.1431, 52.643100},  // interface
    {1.73, 0.0, 0.0, 2.8, 52.643100, 90.0},       // contact
};

}  // namespace

extern "C" int kmcb200_parse_parameters(const char *path, kmcb200_params *out) {
    if (!path || !out) { kmc_set_error("null argument"); return KMCB200_E_ARG; }
    try {
        Parsed P;
        parse_file(path, P);
        *out = P.p;
    } catch (const std::exception &e) {
        kmc_set_error("parameter file %s: %s", path, e.what());
        return KMCB200_E_IO;
    }
    return 0;
}

extern "C" int kmcb200_parse_parameter_vector(const char *path, int which, int cap, double *out) {
    try {
        Parsed P;
        parse_file(path, P);
        const std::vector<double> &v = which == 0 ? P.V_switch : P.t_switch;
        for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
        return (int)v.size();
    } catch (const std::exception &e) {
        kmc_set_error("parameter file %s: %s", path, e.what());
        return KMCB200_E_IO;
    }
}

extern "C" int kmcb200_xyz_count(const char *path) {
    std::ifstream xyz(path);
    if (!xyz.is_open()) { kmc_set_error("cannot open %s", path); return KMCB200_E_IO; }
    std::string line;
    std::getline(xyz, line);
    std::istringstream iss(line);
    int N = 0;
    iss >> N;
    return N;
}

extern "C" int kmcb200_read_xyz(const char *path, int cap, int *element, double *x, double *y, double *z) {
    std::ifstream xyz(path);
    if (!xyz.is_open()) { kmc_set_error("cannot open %s", path); return KMCB200_E_IO; }
    std::string line;
    std::getline(xyz, line);
    int N = 0;
    { std::istringstream iss(line); iss >> N; }
    std::getline(xyz, line);
    double x_ = 0, y_ = 0, z_ = 0;  // like the reference, a short line keeps the previous values
    std::string el;
    for (int i = 0; i < N; ++i) {
        std::getline(xyz, line);
        std::istringstream iss(line);
        iss >> el >> x_ >> y_ >> z_;
        int id = element_from_name(el);
        if (id < 0) { kmc_set_error("%s: unknown element '%s' at site %d", path, el.c_str(), i); return KMCB200_E_IO; }
        if (i < cap) { element[i] = id; x[i] = x_; y[i] = y_; z[i] = z_; }
    }
    return N;
}

extern "C" int kmcb200_make_substoichiometric(int N, int *element, double vacancy_concentration, unsigned rnd_seed) {
    if (!element || N <= 0) { kmc_set_error("invalid argument"); return KMCB200_E_ARG; }
    std::mt19937 rng(0);
    rng.seed(rnd_seed);
    std::uniform_real_distribution<double> dist(0.0, 1.0);
    std::vector<int> atom_ind;  // Device::updateAtomLists
    int num_O = 0;
    for (int i = 0; i < N; ++i) {
        if (element[i] != KMCB200_DEFECT && element[i] != KMCB200_OXYGEN_DEFECT) atom_ind.push_back(i);
        if (element[i] == KMCB200_O) num_O++;
    }
    int N_atom = (int)atom_ind.size();
    int num_V_add = (int)(vacancy_concentration * num_O);
    int converted = 0;
    if (num_V_add > num_O) num_V_add = num_O;  // the reference would loop forever
    while (num_V_add > 0) {
        double r = dist(rng);
        int loc = (int)(r * N_atom);
        if (element[atom_ind[loc]] == KMCB200_O) {
            element[atom_ind[loc]] = KMCB200_VACANCY;
            num_V_add--;
            converted++;
        }
    }
    return converted;
}

extern "C" int kmcb200_num_layers(void) { return 5; }

extern "C" int kmcb200_layer_table(double *E_gen, double *E_rec, double *E_Vdiff, double *E_Odiff, double *start_x,
                                   double *end_x) {
    for (int l = 0; l < 5; ++l) {
        if (E_gen) E_gen[l] = kLayers[l].E_gen;
        if (E_rec) E_rec[l] = kLayers[l].E_rec;
        if (E_Vdiff) E_Vdiff[l] = kLayers[l].E_Vdiff;
        if (E_Odiff) E_Odiff[l] = kLayers[l].E_Odiff;
        if (start_x) start_x[l] = kLayers[l].start_x;
        if (end_x) end_x[l] = kLayers[l].end_x;
    }
    return 5;
}

extern "C" int kmcb200_assign_layers(int N, const double *x, int *site_layer) {
    for (int i = 0; i < N; ++i) {
        int id = -1;
        for (int j = 0; j < 5; ++j)
            if (kLayers[j].start_x <= x[i] && x[i] <= kLayers[j].end_x) id = j;  // last match wins
        if (id < 0) {
            kmc_set_error("Site #%d is not inside the device!", i);
            return KMCB200_E_ARG;
        }
        site_layer[i] = id;
    }
    return 0;
}

extern "C" void kmcb200_partition(int nrows, int nranks, int *counts, int *displs) {
    int per = nrows / nranks;
    for (int i = 0; i < nranks; ++i) counts[i] = per + (i < nrows % nranks ? 1 : 0);
    displs[0] = 0;
    for (int i = 1; i < nranks; ++i) displs[i] = displs[i - 1] + counts[i - 1];
}

extern "C" void kmcb200_partition_aligned(int nrows, int nranks, int *counts, int *displs) {
    const int nchunks = (nrows + KMCB200_CHUNK - 1) / KMCB200_CHUNK;
    // granule: one dot chunk, or one dot GROUP of chunks for systems whose dots are combined in two levels
    const int gran = (nchunks > 256 ? KMCB200_DOT_GROUP : 1) * KMCB200_CHUNK;
    const int ngran = (nrows + gran - 1) / gran;
    int per = ngran / nranks;
    int acc = 0;
    for (int i = 0; i < nranks; ++i) {
        int c = per + (i < ngran % nranks ? 1 : 0);
        int rows = c * gran;
        if (acc + rows > nrows) rows = nrows - acc;
        if (rows < 0) rows = 0;
        displs[i] = acc;
        counts[i] = rows;
        acc += rows;
    }
}

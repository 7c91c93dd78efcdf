// ref_host_shim.cpp -- C-callable window onto the REFERENCE's own host-side code
// (src/input_parser.cpp, src/utils.cpp, src/random_num.h), compiled from /root/reference in place.
// TEST INFRASTRUCTURE: used to validate this repo's parameter parser, xyz reader, RNG stream,
// site_dist and v_solve restatements against the real thing.  Output goes to oracle/_ref/ only.
#include "input_parser.h"   // reference
#include "random_num.h"     // reference
#include "utils.h"          // reference
#include <cstdlib>
#include <cstring>

extern "C" {

// LAPACK/BLAS are absent in this image.  utils.cpp references them only from its dense gesv()/gemm() wrappers
// (src/utils.cpp:421-428,532-538), which nothing in this shim calls; abort loudly if that ever changes.
void dgesv_(int *, int *, double *, int *, int *, double *, int *, int *) { std::abort(); }
void dgemv_(char *, int *, int *, double *, double *, int *, double *, int *, double *, double *, int *) { std::abort(); }

struct ref_params {
    unsigned rnd_seed;
    int restart, pristine, shift, pbc;
    int solve_potential, solve_current, solve_heating_global, solve_heating_local, perturb_structure;
    int log_freq, output_freq;
    int num_atoms_first_layer, num_layers_contact, num_atoms_contact, num_atoms_reservoir;
    int num_metals, metals[8];
    int n_V_switch, n_t_switch, n_lattice, n_shifts;
    double V_switch0, t_switch0;
    double lattice[3], shifts[3];
    double initial_vacancy_concentration, freq, nn_dist, sigma, epsilon, k, high_G, low_G;
    double background_temp, m_r, V0, Icc, Rs, t_ox, A;
    char restart_xyz_file[512], atom_xyz_file[512], interstitial_xyz_file[512];
};

int ref_parse_params(const char *path, ref_params *o) {
    KMCParameters p{std::string(path)};
    std::memset(o, 0, sizeof(*o));
    o->rnd_seed = p.rnd_seed; o->restart = p.restart; o->pristine = p.pristine; o->shift = p.shift; o->pbc = p.pbc;
    o->solve_potential = p.solve_potential; o->solve_current = p.solve_current;
    o->solve_heating_global = p.solve_heating_global; o->solve_heating_local = p.solve_heating_local;
    o->perturb_structure = p.perturb_structure; o->log_freq = p.log_freq; o->output_freq = p.output_freq;
    o->num_atoms_first_layer = p.num_atoms_first_layer; o->num_layers_contact = p.num_layers_contact;
    o->num_atoms_contact = p.num_atoms_contact; o->num_atoms_reservoir = p.num_atoms_reservoir;
    o->num_metals = (int)p.metals.size();
    for (int i = 0; i < o->num_metals && i < 8; ++i) o->metals[i] = (int)p.metals[i];
    o->n_V_switch = (int)p.V_switch.size(); o->n_t_switch = (int)p.t_switch.size();
    o->V_switch0 = p.V_switch.empty() ? 0 : p.V_switch[0];
    o->t_switch0 = p.t_switch.empty() ? 0 : p.t_switch[0];
    o->n_lattice = (int)p.lattice.size(); o->n_shifts = (int)p.shifts.size();
    for (int i = 0; i < 3 && i < o->n_lattice; ++i) o->lattice[i] = p.lattice[i];
    for (int i = 0; i < 3 && i < o->n_shifts; ++i) o->shifts[i] = p.shifts[i];
    o->initial_vacancy_concentration = p.initial_vacancy_concentration; o->freq = p.freq; o->nn_dist = p.nn_dist;
    o->sigma = p.sigma; o->epsilon = p.epsilon; o->k = p.k; o->high_G = p.high_G; o->low_G = p.low_G;
    o->background_temp = p.background_temp; o->m_r = p.m_r; o->V0 = p.V0; o->Icc = p.Icc; o->Rs = p.Rs;
    o->t_ox = p.t_ox; o->A = p.A;
    std::strncpy(o->restart_xyz_file, p.restart_xyz_file.c_str(), 511);
    std::strncpy(o->atom_xyz_file, p.atom_xyz_file.c_str(), 511);
    std::strncpy(o->interstitial_xyz_file, p.interstitial_xyz_file.c_str(), 511);
    return 0;
}

int ref_read_xyz(const char *path, int cap, int *element, double *x, double *y, double *z) {
    std::vector<ELEMENT> e; std::vector<double> vx, vy, vz;
    int N = read_xyz(std::string(path), e, vx, vy, vz);
    for (int i = 0; i < N && i < cap; ++i) { element[i] = (int)e[i]; x[i] = vx[i]; y[i] = vy[i]; z[i] = vz[i]; }
    return N;
}

void *ref_rng_create(unsigned seed) { auto *r = new RandomNumberGenerator(); r->setSeed(seed); return r; }
double ref_rng_next(void *r) { return ((RandomNumberGenerator *)r)->getRandomNumber(); }

double ref_site_dist(double x1, double y1, double z1, double x2, double y2, double z2, const double *lattice, int pbc) {
    std::vector<double> l(lattice, lattice + 3);
    return site_dist(x1, y1, z1, x2, y2, z2, l, pbc != 0);
}
double ref_v_solve(double r, int charge, double sigma, double k, double q) { return v_solve(r, charge, sigma, k, q); }

// The reference's surviving CPU implementation of the long-range charge sum, Device::poisson_gridless
// (src/potential_solver.cpp:74-94): OpenMP over sites i, ALL sites j with a non-zero charge, PBC-aware site_dist, no
// 20 A cutoff -- NOT numerically equivalent to the live GPU kernel, a timing baseline only.  Device itself
// cannot be compiled here (its headers pull in ROCm/MPI), so the 10-line loop is driven from this shim over the
// reference's own compiled site_dist() and v_solve() (src/utils.cpp).  Rows [row_begin, row_end) only: bounded sample.
double ref_poisson_gridless_rows(int N, const double *x, const double *y, const double *z, const int *charge,
                                 const double *lattice, int pbc, double sigma, double k, int row_begin, int row_end,
                                 double *out) {
    std::vector<double> l(lattice, lattice + 3);
    double q = 1.60217663e-19;  // Device::q, src/Device.h:119 (v_solve takes non-const references)
    double checksum = 0.0;
#pragma omp parallel for reduction(+ : checksum) firstprivate(q, sigma, k)
    for (int i = row_begin; i < row_end; i++) {
        double V_temp = 0;
        for (int j = 0; j < N; j++) {
            if (i != j && charge[j] != 0) {
                double r_dist = (1e-10) * site_dist(x[i], y[i], z[i], x[j], y[j], z[j], l, pbc != 0);
                int qj = charge[j];
                V_temp += v_solve(r_dist, qj, sigma, k, q);
            }
        }
        out[i - row_begin] = V_temp;
        checksum += V_temp;
    }
    return checksum;
}

}  // extern "C"

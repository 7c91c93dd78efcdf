// gpu_solvers_b200.hpp -- the reference's entry-point NAMES and ARGUMENT LISTS (src/gpu_solvers.h:36-263) on top of the
// C ABI in kmc_b200.h.  Header-only.  A DeviceKMC host (src/kmc_main.cpp) that includes this header instead of
// gpu_solvers.h and links libkmc_b200.so keeps its call sites for the field-solve + event-selection path:
//
//   compute_neighbor_list, compute_cutoff_list, initialize_sparsity_K, update_charge_gpu,
//   background_potential_gpu_sparse, poisson_gridless_gpu, sum_and_gather_potential, execute_kmc_step_mpi,
//   copytoConstMemory                         (src/kmc_main.cpp:199,205,215,239,342,364,405,479,491)
//
// Type mapping (reference -> here): hipblasHandle_t / hipsolverDnHandle_t -> opaque void* (unused: the solvers are
// hand-written); MPI_Comm -> kmcb200_comm_t (one process per GPU, rank/size carried explicitly); ELEMENT stays an
// int-sized enum with the reference's values; RandomNumberGenerator keeps its interface (std::mt19937 +
// uniform_real_distribution<double>, src/random_num.h) and is kept in step with the device generator.
// Errors: the reference's gpuErrchk prints and exit(1)s (src/utils.h:145-154); KMCB200_CHECK does the same.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "kmc_b200.h"

#define KMCB200_CHECK(call)                                                                      \
    do {                                                                                         \
        int rc_ = (call);                                                                        \
        if (rc_ != 0) {                                                                          \
            std::fprintf(stderr, "kmc_b200: %s (%s:%d)\n", kmcb200_last_error(), __FILE__, __LINE__); \
            std::exit(1);                                                                        \
        }                                                                                        \
    } while (0)

// ---- reference enums (src/utils.h:37-60) -------------------------------------------------------------------------
enum ELEMENT { DEFECT, OXYGEN_DEFECT, VACANCY, O_EL, Hf_EL, Ni_EL, Ti_EL, Pt_EL, N_EL, NULL_ELEMENT };
enum EVENTTYPE { VACANCY_GENERATION, VACANCY_RECOMBINATION, VACANCY_DIFFUSION, ION_DIFFUSION, NULL_EVENT };
static_assert(sizeof(ELEMENT) == sizeof(int), "ELEMENT must be int sized (device arrays are int32)");

typedef void *hipblasHandle_t;      // unused by this implementation
typedef void *hipsolverDnHandle_t;  // unused by this implementation

struct kmcb200_comm_t {  // stands in for MPI_Comm: one process per GPU
    int rank = 0, size = 1;
};
typedef kmcb200_comm_t MPI_Comm_b200;

// src/random_num.h:4-26
class RandomNumberGenerator {
public:
    RandomNumberGenerator() : rng(0) {}
    void setSeed(unsigned int seed) { rng.seed(seed); }
    double getRandomNumber() { return distribution(rng); }
    // raw engine state (624 words + position) for the device generator
    void getState(unsigned *mt624, int *pos) {
        std::stringstream ss;
        ss << rng;
        for (int i = 0; i < 624; ++i) { unsigned long v; ss >> v; mt624[i] = (unsigned)v; }
        unsigned long p; ss >> p; *pos = (int)p;
    }
    void discardDoubles(unsigned long long n) { rng.discard(2ULL * n); }  // one double = two 32-bit draws
private:
    std::mt19937 rng;
    std::uniform_real_distribution<double> distribution{0.0, 1.0};
};

// src/KMC_comm.h: only the row partitions the path uses (counts/displs per logical communicator)
struct KMC_comm {
    kmcb200_comm_t comm_K, comm_pairwise, comm_events;
    int rank_K = 0, size_K = 1, rank_pairwise = 0, size_pairwise = 1, rank_events = 0, size_events = 1;
    std::vector<int> counts_K, displs_K, counts_pairwise, displs_pairwise, counts_events, displs_events;
    KMC_comm(kmcb200_comm_t world, int nrows_K, int /*nrows_T*/, int nrows_pairwise, int nrows_events) {
        comm_K = comm_pairwise = comm_events = world;
        rank_K = rank_pairwise = rank_events = world.rank;
        size_K = size_pairwise = size_events = world.size;
        auto part = [&](int n, std::vector<int> &c, std::vector<int> &d) {
            c.resize(world.size); d.resize(world.size);
            kmcb200_partition(n, world.size, c.data(), d.data());  // src/KMC_comm.h:249-263
        };
        part(nrows_K, counts_K, displs_K);
        part(nrows_pairwise, counts_pairwise, displs_pairwise);
        part(nrows_events, counts_events, displs_events);
    }
};

// src/gpu_buffers.h: the device SoA the path touches (+ the library handles that replace K_distributed etc.)
struct GPUBuffers {
    kmcb200_ctx *ctx = nullptr;
    int N_ = 0, nn_ = 0, N_cutoff_ = 0, num_metal_types_ = 0;
    ELEMENT *site_element = nullptr;
    int *site_charge = nullptr, *site_layer = nullptr, *neigh_idx = nullptr, *cutoff_idx = nullptr;
    double *site_x = nullptr, *site_y = nullptr, *site_z = nullptr;
    double *site_potential_boundary = nullptr, *site_potential_charge = nullptr;
    std::vector<int> metals_h;
    double lattice_h[3] = {0, 0, 0}, sigma_h = 0, k_h = 0, freq_h = 0, T_bg_h = 0;
    kmcb200_kmat *K_distributed = nullptr;  // replaces Distributed_matrix* K_distributed + K_p_distributed + contact CSR
    kmcb200_events *events = nullptr;       // event list + device RNG (allocated by compute_neighbor_list)
    int last_cg_iterations = 0, last_n_events = 0;

    template <class T>
    T *dmalloc(size_t n) { void *p = nullptr; KMCB200_CHECK(kmcb200_malloc(ctx, &p, n * sizeof(T))); return (T *)p; }
    template <class T>
    void h2d(T *dst, const T *src, size_t n) { KMCB200_CHECK(kmcb200_memcpy_h2d(ctx, dst, src, n * sizeof(T))); }
    template <class T>
    void d2h(T *dst, const T *src, size_t n) { KMCB200_CHECK(kmcb200_memcpy_d2h(ctx, dst, src, n * sizeof(T))); }

    // GPUBuffers ctor of the reference (src/gpu_buffers.h:93-160)
    GPUBuffers(kmcb200_ctx *ctx_, const std::vector<int> &site_layer_in, double freq_in, int N,
               const std::vector<int> &site_element_in, const std::vector<double> &x, const std::vector<double> &y,
               const std::vector<double> &z, int nn, double sigma_in, double k_in, const double *lattice_in,
               const std::vector<int> &metals, double T_bg_in)
        : ctx(ctx_), N_(N), nn_(nn), num_metal_types_((int)metals.size()), metals_h(metals) {
        site_element = (ELEMENT *)dmalloc<int>(N);
        site_charge = dmalloc<int>(N); site_layer = dmalloc<int>(N);
        site_x = dmalloc<double>(N); site_y = dmalloc<double>(N); site_z = dmalloc<double>(N);
        site_potential_boundary = dmalloc<double>(N); site_potential_charge = dmalloc<double>(N);
        h2d((int *)site_element, site_element_in.data(), N); h2d(site_layer, site_layer_in.data(), N);
        h2d(site_x, x.data(), N); h2d(site_y, y.data(), N); h2d(site_z, z.data(), N);
        KMCB200_CHECK(kmcb200_memset(ctx, site_charge, 0, N * sizeof(int)));
        KMCB200_CHECK(kmcb200_memset(ctx, site_potential_boundary, 0, N * sizeof(double)));
        KMCB200_CHECK(kmcb200_memset(ctx, site_potential_charge, 0, N * sizeof(double)));
        for (int i = 0; i < 3; ++i) lattice_h[i] = lattice_in[i];
        sigma_h = sigma_in; k_h = k_in; freq_h = freq_in; T_bg_h = T_bg_in;
        KMCB200_CHECK(kmcb200_synchronize(ctx));
    }
};

// ---- src/gpu_solvers.h:43 -----------------------------------------------------------------------------------------
// (nn_dist = 3.5 and max_num_neighbors = 52 are hard-coded in the reference, src/neighbor_lists_gpu.cu:265-266)
inline void compute_neighbor_list(kmcb200_comm_t &event_comm, int *counts, int *displ, GPUBuffers &gpubuf) {
    const int rank = event_comm.rank;
    gpubuf.neigh_idx = gpubuf.dmalloc<int>((size_t)counts[rank] * 52);
    KMCB200_CHECK(kmcb200_compute_neighbor_list(gpubuf.ctx, gpubuf.N_, gpubuf.site_x, gpubuf.site_y, gpubuf.site_z, 3.5,
                                                52, displ[rank], counts[rank], gpubuf.neigh_idx));
}
// src/gpu_solvers.h:46.  Only N_cutoff_ is produced: the list itself is never needed by poisson_gridless_gpu here.
inline void compute_cutoff_list(kmcb200_comm_t &pairwise_comm, int *counts, int *displ, GPUBuffers &gpubuf) {
    const int rank = pairwise_comm.rank;
    KMCB200_CHECK(kmcb200_cutoff_size(gpubuf.ctx, gpubuf.N_, (const int *)gpubuf.site_element, gpubuf.site_x,
                                      gpubuf.site_y, gpubuf.site_z, 20.0, displ[rank], counts[rank], nullptr,
                                      &gpubuf.N_cutoff_));
}
// src/gpu_solvers.h:53
inline void initialize_sparsity_K(GPUBuffers &gpubuf, int pbc, const double nn_dist, int num_atoms_contact,
                                  KMC_comm &kmc_comm) {
    const int r = kmc_comm.rank_K;
    KMCB200_CHECK(kmcb200_initialize_sparsity_K(gpubuf.ctx, gpubuf.N_, gpubuf.site_x, gpubuf.site_y, gpubuf.site_z,
                                                gpubuf.lattice_h, pbc, nn_dist, num_atoms_contact, num_atoms_contact,
                                                kmc_comm.displs_K[r], kmc_comm.counts_K[r], &gpubuf.K_distributed));
}
// src/gpu_solvers.h:149-153
inline void update_charge_gpu(ELEMENT *d_site_element, int *d_site_charge, int *d_neigh_idx, int N, int nn,
                              const std::vector<int> &metals, const int *count, const int *displ, kmcb200_comm_t &comm,
                              kmcb200_ctx *ctx) {
    KMCB200_CHECK(kmcb200_update_charge(ctx, (const int *)d_site_element, d_site_charge, d_neigh_idx, N, nn,
                                        metals.data(), (int)metals.size(), displ[comm.rank], count[comm.rank]));
}
// src/gpu_solvers.h:162-164
inline void background_potential_gpu_sparse(hipblasHandle_t, hipsolverDnHandle_t, GPUBuffers &gpubuf, const int N,
                                            const int N_left_tot, const int N_right_tot, const double d_Vd,
                                            const int /*pbc*/, const double d_high_G, const double d_low_G,
                                            const double /*nn_dist*/, const int /*num_metals*/,
                                            int /*kmc_step_count*/) {
    KMCB200_CHECK(kmcb200_background_potential(gpubuf.ctx, gpubuf.K_distributed, N, N_left_tot, N_right_tot,
                                               (const int *)gpubuf.site_element, gpubuf.site_charge,
                                               gpubuf.metals_h.data(), gpubuf.num_metal_types_, d_Vd, d_high_G, d_low_G,
                                               gpubuf.site_potential_boundary, &gpubuf.last_cg_iterations));
}
// src/gpu_solvers.h:173-178 (sigma / k are host scalars here; cutoff_window / cutoff_idx are not needed)
inline void poisson_gridless_gpu(kmcb200_ctx *ctx, const int /*num_atoms_contact*/, const int /*pbc*/, const int N,
                                 const double sigma, const double k, const double *posx, const double *posy,
                                 const double *posz, const ELEMENT *site_element, const int *site_charge,
                                 double *site_potential_charge, const int rank, const int /*size*/, const int *count,
                                 const int *displ) {
    KMCB200_CHECK(kmcb200_poisson_gridless(ctx, N, posx, posy, posz, (const int *)site_element, site_charge, sigma, k,
                                           20.0, displ[rank], count[rank], site_potential_charge));
}
// src/gpu_solvers.h:181
inline void sum_and_gather_potential(GPUBuffers &gpubuf, int /*num_atoms_first_layer*/, KMC_comm &) {
    KMCB200_CHECK(kmcb200_sum_potential(gpubuf.ctx, gpubuf.N_, gpubuf.site_potential_charge,
                                        gpubuf.site_potential_boundary));
}
// src/gpu_solvers.h:262
inline void copytoConstMemory(GPUBuffers &gpubuf, std::vector<double> E_gen, std::vector<double> E_rec,
                              std::vector<double> E_Vdiff, std::vector<double> E_Odiff) {
    if (!gpubuf.events)
        KMCB200_CHECK(kmcb200_events_create(gpubuf.ctx, gpubuf.N_, gpubuf.nn_, gpubuf.neigh_idx, &gpubuf.events));
    KMCB200_CHECK(kmcb200_set_activation_energies(gpubuf.events, (int)E_gen.size(), E_gen.data(), E_rec.data(),
                                                  E_Vdiff.data(), E_Odiff.data()));
}
// src/gpu_solvers.h:250-260.  The host RandomNumberGenerator stays the source of truth: its state is uploaded before
// the device loop and advanced by the 2 doubles per event the loop consumed (src/kmc_events.cu:469,515).
inline double execute_kmc_step_mpi(kmcb200_comm_t, GPUBuffers &gpubuf, const int N, const int * /*count*/,
                                   const int * /*displs*/, const int nn, const int *neigh_idx, const int *site_layer,
                                   const int /*pbc*/, const double T_bg, const double freq, const double sigma,
                                   const double k, const double *posx, const double *posy, const double *posz,
                                   const double *site_potential_charge, ELEMENT *site_element, int *site_charge,
                                   RandomNumberGenerator &rng) {
    unsigned mt[624];
    int pos = 0;
    rng.getState(mt, &pos);
    KMCB200_CHECK(kmcb200_rng_set_state(gpubuf.events, mt, pos));
    double event_time = 0.0;
    int n_events = 0;
    KMCB200_CHECK(kmcb200_execute_kmc_step(gpubuf.ctx, gpubuf.events, N, nn, neigh_idx, site_layer, T_bg, freq, sigma, k,
                                           posx, posy, posz, site_potential_charge, (int *)site_element, site_charge, 0,
                                           &event_time, &n_events));
    rng.discardDoubles(2ULL * (unsigned long long)n_events);
    gpubuf.last_n_events = n_events;
    return event_time;
}

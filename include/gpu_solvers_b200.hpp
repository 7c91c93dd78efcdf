// gpu_solvers_b200.hpp -- the reference's solver-layer interface on top of the C ABI of libkmc_b200.so (kmc_b200.h).
//
// A DeviceKMC host that includes this header instead of src/gpu_solvers.h, src/gpu_buffers.h, src/KMC_comm.h and the
// dist_iterative headers, and links libkmc_b200.so, keeps its call sites UNCHANGED for the field-solve + event-selection
// path: every function below has the name and the exact parameter list of the reference declaration it cites.
// tests/test_shim_compiles_reference.py compiles the reference's own setup and superstep code
// (src/kmc_main.cpp:184-240 and :328-545, taken verbatim from the reference checkout at test time) against it.
//
//   src/gpu_solvers.h:43   compute_neighbor_list        :46  compute_cutoff_list        :53  initialize_sparsity_K
//   :54  initialize_sparsity_CB    :57  initialize_sparsity_T      :143-146 update_CB_edge_gpu_sparse
//   :149-153 update_charge_gpu     :162-164 background_potential_gpu_sparse    :173-178 poisson_gridless_gpu
//   :181 sum_and_gather_potential  :212-218 update_power_gpu_sparse_dist       :250-260 execute_kmc_step_mpi
//   :262 copytoConstMemory         :229-233 update_temperatureglobal_gpu
//   dist_iterative/dist_conjugate_gradient.h:33-47  iterative_solver::conjugate_gradient_jacobi<spmv>
//   dist_iterative/dist_spmv.h:22-27                dspmv::gpu_packing_cam
//   src/gpu_buffers.h:12-162 GPUBuffers   src/KMC_comm.h:4-391 KMC_comm   src/random_num.h:4-26 RandomNumberGenerator
//
// Type mapping: hipblas / hipsolver / rocsparse handles and hipStream_t are opaque pointers (the solvers are hand-written
// kernels; nothing is forwarded to a vendor library); rocsparse_dnvec_descr is a (pointer, length) pair; MPI names come from
// kmcb200_mpi_compat.h unless KMCB200_HAVE_MPI is defined.  Device and KMCParameters are only forward declared: the entry
// points that receive them (compute_neighbor_list / compute_cutoff_list) read everything they need from GPUBuffers, like
// the reference (src/neighbor_lists_gpu.cu:257-373 reads gpubuf and hard-coded constants only).
// What the reference keeps in __constant__ memory, rocSPARSE descriptors and Distributed_matrix internals lives in
// library handles (kmcb200_kmat / kmcb200_events / kmcb200_tmat) owned by a per-process runtime record.
// Errors: the reference's gpuErrchk prints and exit(1)s (src/utils.h:145-154); KMCB200_CHECK does the same.
// Multi-rank (one process per GPU, like the reference's one MPI rank per GCD): the K solve (row blocks) and the Coulomb sum
// (site rows) are sharded, charges / rate list / event selection are replicated (every rank draws the same events from the
// same generator).  The bootstrap (CUDA-IPC handles, halo need maps, barriers) runs over a shared directory named by the
// environment variable KMCB200_RENDEZVOUS (kmcb200_rdv_*); the per-step exchanges run over NVLink peer memory
// (kmcb200_comm_allgather, and the halo / dot exchanges inside the PCG kernels).  Rank and size come from the launcher's
// environment (kmcb200_mpi_compat.h).  Rank boundaries of the K rows are placed on the library's dot granule
// (kmcb200_partition_aligned) instead of the reference's row-count split, which keeps results bit-identical for every
// rank count.  The Kirchhoff chain is single-rank.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "kmc_b200.h"
#ifdef KMCB200_HAVE_MPI
#include <mpi.h>
#else
#include "kmcb200_mpi_compat.h"
#endif

#define KMCB200_CHECK(call)                                                                      \
    do {                                                                                         \
        int rc_ = (call);                                                                        \
        if (rc_ != 0) {                                                                          \
            std::fprintf(stderr, "kmc_b200: %s (%s:%d)\n", kmcb200_last_error(), __FILE__, __LINE__); \
            std::exit(1);                                                                        \
        }                                                                                        \
    } while (0)

// ---- reference enums and small types (src/utils.h:37-72) -----------------------------------------------------------
enum ELEMENT { DEFECT, OXYGEN_DEFECT, VACANCY, O_EL, Hf_EL, Ni_EL, Ti_EL, Pt_EL, N_EL, NULL_ELEMENT };
enum EVENTTYPE { VACANCY_GENERATION, VACANCY_RECOMBINATION, VACANCY_DIFFUSION, ION_DIFFUSION, NULL_EVENT };
static_assert(sizeof(ELEMENT) == sizeof(int), "ELEMENT must be int sized (device arrays are int32)");
struct Layer {
    std::string type;
    double E_gen_0 = 0, E_rec_1 = 0, E_diff_2 = 0, E_diff_3 = 0;
    double start_x = 0, end_x = 0;
    double init_vac_percentage = 0;
    Layer() {}
    void init_layer(std::string type_, double E_gen_0_, double E_rec_1_, double E_diff_2_, double E_diff_3_, double start_x_,
                    double end_x_) {
        type = type_; E_gen_0 = E_gen_0_; E_rec_1 = E_rec_1_; E_diff_2 = E_diff_2_; E_diff_3 = E_diff_3_;
        start_x = start_x_; end_x = end_x_;
    }
};

typedef void *hipblasHandle_t;
typedef void *hipsolverHandle_t;
typedef void *hipsolverDnHandle_t;
typedef void *hipsparseHandle_t;
typedef void *hipStream_t;
typedef void *rocsparse_handle;
struct rocsparse_dnvec_descr {  // dense vector descriptor: what the solver layer needs of it
    double *values = nullptr;
    long long size = 0;
};
inline int hipblasCreate(hipblasHandle_t *h) { *h = nullptr; return 0; }
inline int hipsolverCreate(hipsolverHandle_t *h) { *h = nullptr; return 0; }
inline int hipblasDestroy(hipblasHandle_t) { return 0; }
inline int hipsolverDestroy(hipsolverHandle_t) { return 0; }

class Device;         // src/Device.h       (host side; stays the caller's)
class KMCParameters;  // src/input_parser.h (host side; stays the caller's)
class GPUBuffers;

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

// ---- per-process runtime: the library handles behind the reference's globals ---------------------------------------
namespace kmcb200 {
struct Runtime {
    kmcb200_ctx *ctx = nullptr;
    GPUBuffers *gpubuf = nullptr;      // the (single) GPUBuffers of the process: poisson_gridless_gpu needs site_element
    kmcb200_events *events = nullptr;  // event list + device MT19937 (src/kmc_events.cu:356-361 + __constant__ energies)
    kmcb200_tmat *tmat = nullptr;      // T_distributed + tunnel block
    kmcb200_comm *comm = nullptr;      // row-sharded exchange plan of the K solve (size > 1)
    kmcb200_rdv *rdv = nullptr;        // bootstrap rendezvous (size > 1)
    std::vector<int> metals_h;
    std::vector<double> E_gen, E_rec, E_Vdiff, E_Odiff;
    int last_cg_iterations = 0, last_n_events = 0, last_T_iterations = 0;
};
inline Runtime &rt() {
    static Runtime r;
    return r;
}
inline kmcb200_ctx *ctx() {
    Runtime &r = rt();
    if (!r.ctx) {
        const char *d = std::getenv("KMCB200_DEVICE") ? std::getenv("KMCB200_DEVICE") : std::getenv("LOCAL_RANK");
        KMCB200_CHECK(kmcb200_create(&r.ctx, d ? std::atoi(d) : 0, nullptr));
    }
    return r.ctx;
}
template <class T>
inline T *dmalloc(size_t n) {
    void *p = nullptr;
    KMCB200_CHECK(kmcb200_malloc(ctx(), &p, (n ? n : 1) * sizeof(T)));
    return (T *)p;
}
template <class T>
inline void h2d(T *dst, const T *src, size_t n) { KMCB200_CHECK(kmcb200_memcpy_h2d(ctx(), dst, src, n * sizeof(T))); }
template <class T>
inline void d2h(T *dst, const T *src, size_t n) { KMCB200_CHECK(kmcb200_memcpy_d2h(ctx(), dst, src, n * sizeof(T))); }
inline void sync() { KMCB200_CHECK(kmcb200_synchronize(ctx())); }
inline void single_rank_only(int size, const char *who) {
    if (size > 1) {
        std::fprintf(stderr, "kmc_b200: %s is a single-rank call in this build\n", who);
        std::exit(1);
    }
}
// bootstrap rendezvous of the node's ranks (opened on first use)
inline kmcb200_rdv *rdv(int rank, int size) {
    Runtime &r = rt();
    if (!r.rdv) {
        const char *dir = std::getenv("KMCB200_RENDEZVOUS");
        if (!dir) {
            std::fprintf(stderr, "kmc_b200: %d ranks need KMCB200_RENDEZVOUS=<directory shared by the ranks> for the bootstrap\n", size);
            std::exit(1);
        }
        KMCB200_CHECK(kmcb200_rdv_open(dir, rank, size, &r.rdv));
    }
    return r.rdv;
}
}  // namespace kmcb200

inline int hipDeviceSynchronize() { kmcb200::sync(); return 0; }
#ifndef KMCB200_HAVE_MPI
inline int kmcb200_host_barrier(MPI_Comm c) {
    kmcb200::sync();
    KMCB200_CHECK(kmcb200_rdv_barrier(kmcb200::rdv(c->rank, c->size)));
    return 0;
}
#endif

// ---- dist_iterative/dist_objects.h: the two objects the callers hold ---------------------------------------------------
class Distributed_matrix {  // dist_objects.h:66-232; the CSR blocks / halo plan live in the kmcb200_kmat
public:
    int matrix_size = 0, rows_this_rank = 0;
    long long nnz = 0;
    int size = 1, rank = 0;
    int *counts = nullptr, *displacements = nullptr;
    MPI_Comm comm = MPI_COMM_NULL;
    kmcb200_kmat *kmat = nullptr;
    Distributed_matrix(kmcb200_kmat *K, int matrix_size_, int *counts_, int *displacements_, MPI_Comm comm_)
        : matrix_size(matrix_size_), counts(counts_), displacements(displacements_), comm(comm_), kmat(K) {
        MPI_Comm_rank(comm_, &rank);
        MPI_Comm_size(comm_, &size);
        KMCB200_CHECK(kmcb200_kmat_info(K, &rows_this_rank, &nnz, nullptr, nullptr));
    }
    ~Distributed_matrix() { kmcb200_kmat_destroy(kmat); }
};
class Distributed_vector {  // dist_objects.h:10-49; vec_d[0] is this rank's piece (halo pieces live in the exchange arena)
public:
    int matrix_size = 0, rows_this_rank = 0, size = 1, rank = 0;
    int *counts = nullptr, *displacements = nullptr;
    int number_of_neighbours = 1;
    int *neighbours = nullptr;
    MPI_Comm comm = MPI_COMM_NULL;
    double **vec_d = nullptr;
    Distributed_vector(int matrix_size_, int *counts_, int *displacements_, int number_of_neighbours_, int *neighbours_,
                       MPI_Comm comm_)
        : matrix_size(matrix_size_), counts(counts_), displacements(displacements_),
          number_of_neighbours(number_of_neighbours_), neighbours(neighbours_), comm(comm_) {
        MPI_Comm_rank(comm_, &rank);
        MPI_Comm_size(comm_, &size);
        rows_this_rank = counts_[rank];
        vec_d = new double *[1];
        vec_d[0] = kmcb200::dmalloc<double>(rows_this_rank);
    }
    ~Distributed_vector() { if (vec_d) { kmcb200_free(kmcb200::ctx(), vec_d[0]); delete[] vec_d; } }
};

// ---- src/KMC_comm.h:4-391 (split == false path of the shipped main: every logical communicator is the world) ----------
class KMC_comm {
public:
    int rank_global = 0, rank_K = 0, rank_T = 0, rank_pairwise = 0, rank_events = 0;
    int size_global = 1, size_K = 1, size_T = 1, size_pairwise = 1, size_events = 1;
    int root_K = 0, root_T = 0, root_pairwise = 0;
    MPI_Group group_global = 0, group_K = 0, group_T = 0, group_pairwise = 0;
    MPI_Comm comm_K = MPI_COMM_NULL, comm_T = MPI_COMM_NULL, comm_pairwise = MPI_COMM_NULL, comm_events = MPI_COMM_NULL;
    int *counts_K = nullptr, *counts_T = nullptr, *counts_pairwise = nullptr, *counts_events = nullptr;
    int *displs_K = nullptr, *displs_T = nullptr, *displs_pairwise = nullptr, *displs_events = nullptr;

    KMC_comm(MPI_Comm comm_global, int nrows_K, int nrows_T, int nrows_pairwise, int nrows_events, bool split, int *ratio) {
        (void)ratio;
        MPI_Comm_rank(comm_global, &rank_global);
        MPI_Comm_size(comm_global, &size_global);
        if (split) {
            std::fprintf(stderr, "kmc_b200: KMC_comm(split = true) is not supported (the shipped main uses split = false, "
                                 "src/kmc_main.cpp:161)\n");
            std::exit(1);
        }
        comm_K = comm_T = comm_pairwise = comm_events = comm_global;
        rank_K = rank_T = rank_pairwise = rank_events = rank_global;
        size_K = size_T = size_pairwise = size_events = size_global;
        auto part = [&](int n, int *&c, int *&d) {  // src/KMC_comm.h:249-263
            c = new int[size_global];
            d = new int[size_global];
            kmcb200_partition(n, size_global, c, d);
        };
        part(nrows_K, counts_K, displs_K);
        if (size_global > 1) kmcb200_partition_aligned(nrows_K, size_global, counts_K, displs_K);  // dot-granule boundaries
        part(nrows_T, counts_T, displs_T);
        part(nrows_pairwise, counts_pairwise, displs_pairwise);
        part(nrows_events, counts_events, displs_events);
        // the reference forces comm_T = MPI_COMM_NULL here (src/KMC_comm.h:243): the current solver never runs in its
        // shipped main.  Set KMCB200_ENABLE_CURRENT=1 to keep comm_T alive and run the Kirchhoff chain.
        if (!std::getenv("KMCB200_ENABLE_CURRENT")) comm_T = MPI_COMM_NULL;
    }
};

// ---- src/gpu_buffers.h:12-162 -----------------------------------------------------------------------------------------
// Member variables are pointers to GPU memory unless passed by value, like the reference's.
class GPUBuffers {
public:
    int *site_charge = nullptr;
    double *site_power = nullptr, *site_potential_boundary = nullptr, *site_potential_charge = nullptr,
           *site_temperature = nullptr;
    double *site_CB_edge = nullptr;
    double *T_bg = nullptr;
    double *atom_power = nullptr, *atom_CB_edge = nullptr, *atom_virtual_potentials = nullptr;
    int *atom_charge = nullptr;
    ELEMENT *site_element = nullptr, *atom_element = nullptr;
    double *site_x = nullptr, *site_y = nullptr, *site_z = nullptr;
    double *atom_x = nullptr, *atom_y = nullptr, *atom_z = nullptr;
    ELEMENT *metal_types = nullptr;
    double *sigma = nullptr, *k = nullptr, *lattice = nullptr, *freq = nullptr;
    int *neigh_idx = nullptr, *cutoff_window = nullptr, *cutoff_idx = nullptr, *site_layer = nullptr;
    int num_metal_types_ = 0, N_ = 0, nn_ = 0, N_atom_ = 0, N_sub_ = 0, N_cutoff_ = 0;
    std::vector<double> E_gen_host, E_rec_host, E_Vdiff_host, E_Odiff_host;
    Distributed_matrix *K_distributed = nullptr;
    Distributed_vector *K_p_distributed = nullptr;
    Distributed_matrix *T_distributed = nullptr;  // (the T system lives in the runtime's kmcb200_tmat)
    Distributed_vector *T_p_distributed = nullptr;
    double lattice_host[3] = {0, 0, 0};

    void sync_HostToGPU(Device &device);  // src/gpu_buffers.cpp:10-34: defined by the host next to its Device class
    void sync_GPUToHost(Device &device);  // src/gpu_buffers.cpp:36-55   (helpers: kmcb200_sync_host_to_gpu / _gpu_to_host)
    void copy_Tbg_toGPU(double new_T_bg) { kmcb200::h2d(T_bg, &new_T_bg, 1); kmcb200::sync(); }

    GPUBuffers() {}
    GPUBuffers(std::vector<Layer> layers, std::vector<int> site_layer_in, double freq_in, int N, int N_atom,
               std::vector<ELEMENT> site_element_in, std::vector<double> site_x_in, std::vector<double> site_y_in,
               std::vector<double> site_z_in, int nn, double sigma_in, double k_in, std::vector<double> lattice_in,
               std::vector<ELEMENT> metals, int num_metals_types, MPI_Comm comm, int N_contact) {
        using namespace kmcb200;
        (void)comm; (void)N_contact;
        N_ = N; N_atom_ = N_atom; N_sub_ = N_atom + 1; nn_ = nn; num_metal_types_ = num_metals_types;
        for (auto l : layers) {
            E_gen_host.push_back(l.E_gen_0); E_rec_host.push_back(l.E_rec_1);
            E_Vdiff_host.push_back(l.E_diff_2); E_Odiff_host.push_back(l.E_diff_3);
        }
        site_layer = dmalloc<int>(N); site_element = (ELEMENT *)dmalloc<int>(N);
        metal_types = (ELEMENT *)dmalloc<int>(num_metals_types);
        site_x = dmalloc<double>(N); site_y = dmalloc<double>(N); site_z = dmalloc<double>(N);
        site_power = dmalloc<double>(N); site_CB_edge = dmalloc<double>(N);
        site_potential_boundary = dmalloc<double>(N); site_potential_charge = dmalloc<double>(N);
        site_temperature = dmalloc<double>(N); site_charge = dmalloc<int>(N);
        T_bg = dmalloc<double>(1); sigma = dmalloc<double>(1); k = dmalloc<double>(1); lattice = dmalloc<double>(3);
        freq = dmalloc<double>(1);
        atom_CB_edge = dmalloc<double>(N_atom + 2); atom_virtual_potentials = dmalloc<double>(N_atom + 2);
        kmcb200_ctx *c = ctx();
        KMCB200_CHECK(kmcb200_memset(c, atom_virtual_potentials, 0, (size_t)(N_atom + 2) * sizeof(double)));
        KMCB200_CHECK(kmcb200_memset(c, site_charge, 0, (size_t)N * sizeof(int)));
        KMCB200_CHECK(kmcb200_memset(c, site_power, 0, (size_t)N * sizeof(double)));
        KMCB200_CHECK(kmcb200_memset(c, site_CB_edge, 0, (size_t)N * sizeof(double)));
        KMCB200_CHECK(kmcb200_memset(c, site_potential_boundary, 0, (size_t)N * sizeof(double)));
        KMCB200_CHECK(kmcb200_memset(c, site_potential_charge, 0, (size_t)N * sizeof(double)));
        h2d(site_layer, site_layer_in.data(), N);
        h2d(site_x, site_x_in.data(), N); h2d(site_y, site_y_in.data(), N); h2d(site_z, site_z_in.data(), N);
        h2d((int *)site_element, (const int *)site_element_in.data(), N);
        h2d((int *)metal_types, (const int *)metals.data(), num_metals_types);
        h2d(sigma, &sigma_in, 1); h2d(k, &k_in, 1); h2d(freq, &freq_in, 1); h2d(lattice, lattice_in.data(), 3);
        for (int i = 0; i < 3; ++i) lattice_host[i] = lattice_in[i];
        sync();
        Runtime &r = rt();
        r.gpubuf = this;
        r.metals_h.assign(num_metals_types, 0);
        for (int i = 0; i < num_metals_types; ++i) r.metals_h[i] = (int)metals[i];
    }
    void freeGPUmemory() {}
};

// helpers for the host's GPUBuffers::sync_* definitions (any Device-like class with the reference's member names)
template <class DeviceT>
inline void kmcb200_sync_host_to_gpu(GPUBuffers &g, DeviceT &device) {  // src/gpu_buffers.cpp:10-34
    using namespace kmcb200;
    h2d((int *)g.site_element, (const int *)device.site_element.data(), g.N_);
    h2d(g.site_charge, device.site_charge.data(), g.N_);
    h2d(g.site_power, device.site_power.data(), g.N_);
    h2d(g.site_CB_edge, device.site_CB_edge.data(), g.N_);
    h2d(g.site_potential_boundary, device.site_potential_boundary.data(), g.N_);
    h2d(g.site_potential_charge, device.site_potential_charge.data(), g.N_);
    h2d(g.site_temperature, device.site_temperature.data(), g.N_);
    h2d(g.T_bg, &device.T_bg, 1);
    sync();
}
template <class DeviceT>
inline void kmcb200_sync_gpu_to_host(GPUBuffers &g, DeviceT &device) {  // src/gpu_buffers.cpp:36-55
    using namespace kmcb200;
    d2h((int *)device.site_element.data(), (const int *)g.site_element, g.N_);
    d2h(device.site_charge.data(), g.site_charge, g.N_);
    d2h(device.site_power.data(), g.site_power, g.N_);
    d2h(device.site_CB_edge.data(), g.site_CB_edge, g.N_);
    d2h(device.site_potential_boundary.data(), g.site_potential_boundary, g.N_);
    d2h(device.site_potential_charge.data(), g.site_potential_charge, g.N_);
    d2h(device.site_temperature.data(), g.site_temperature, g.N_);
    d2h(&device.T_bg, g.T_bg, 1);
    sync();
}

// =====================================================================================================================
// Entry points (reference names and parameter lists)
// =====================================================================================================================

// src/gpu_solvers.h:43.  nn_dist = 3.5 and max_num_neighbors = 52 are hard-coded by the reference
// (src/neighbor_lists_gpu.cu:265-266).  The table is built for ALL sites on every rank: charge update, rate list and event
// selection are replicated in this build (north_star: event selection stays single-GPU).
inline void compute_neighbor_list(MPI_Comm &event_comm, int *counts, int *displ, Device &device, GPUBuffers &gpubuf,
                                  KMCParameters &p) {
    (void)event_comm; (void)counts; (void)displ; (void)device; (void)p;
    gpubuf.neigh_idx = kmcb200::dmalloc<int>((size_t)gpubuf.N_ * 52);
    KMCB200_CHECK(kmcb200_compute_neighbor_list(kmcb200::ctx(), gpubuf.N_, gpubuf.site_x, gpubuf.site_y, gpubuf.site_z, 3.5,
                                                52, 0, gpubuf.N_, gpubuf.neigh_idx));
}
// src/gpu_solvers.h:46.  Produces gpubuf.N_cutoff_ (src/neighbor_lists_gpu.cu:340-342); the 20 A list itself
// (N x N_cutoff int32) is never needed by poisson_gridless_gpu here, so cutoff_idx / cutoff_window stay null.
inline void compute_cutoff_list(MPI_Comm &pairwise_comm, int *counts, int *displ, Device &device, GPUBuffers &gpubuf,
                                KMCParameters &p) {
    (void)device; (void)p;
    int rank = 0;
    MPI_Comm_rank(pairwise_comm, &rank);
    KMCB200_CHECK(kmcb200_cutoff_size(kmcb200::ctx(), gpubuf.N_, (const int *)gpubuf.site_element, gpubuf.site_x,
                                      gpubuf.site_y, gpubuf.site_z, 20.0, displ[rank], counts[rank], nullptr,
                                      &gpubuf.N_cutoff_));
}
// src/gpu_solvers.h:53
inline void initialize_sparsity_K(GPUBuffers &gpubuf, int pbc, const double nn_dist, int num_atoms_contact,
                                  KMC_comm &kmc_comm) {
    using namespace kmcb200;
    const int r = kmc_comm.rank_K, size = kmc_comm.size_K;
    const int n = gpubuf.N_ - 2 * num_atoms_contact;
    Runtime &R = rt();
    if (size > 1) {
        // exchange plan over NVLink peer memory: arena + CUDA-IPC handles exchanged through the rendezvous directory
        long long cap = 0;
        for (int q = 0; q < size; ++q) {
            if (kmc_comm.counts_K[q] > cap) cap = kmc_comm.counts_K[q];
            if (kmc_comm.counts_pairwise[q] > cap) cap = kmc_comm.counts_pairwise[q];
        }
        KMCB200_CHECK(kmcb200_comm_create_ex(ctx(), r, size, n, kmc_comm.counts_K, kmc_comm.displs_K, cap, &R.comm));
        std::vector<unsigned char> mine(64), all((size_t)64 * size);
        KMCB200_CHECK(kmcb200_comm_ipc_handle(R.comm, mine.data()));
        KMCB200_CHECK(kmcb200_rdv_allgather(rdv(r, size), mine.data(), 64, all.data()));
        KMCB200_CHECK(kmcb200_comm_open_peers(R.comm, all.data()));
    }
    kmcb200_kmat *K = nullptr;
    KMCB200_CHECK(kmcb200_initialize_sparsity_K(ctx(), gpubuf.N_, gpubuf.site_x, gpubuf.site_y, gpubuf.site_z,
                                                gpubuf.lattice_host, pbc, nn_dist, num_atoms_contact, num_atoms_contact,
                                                kmc_comm.displs_K[r], kmc_comm.counts_K[r], &K));
    if (size > 1) {
        KMCB200_CHECK(kmcb200_kmat_attach_comm(K, R.comm));
        // halo need maps: which rows of the other ranks does my block reference (all-gathered out of band, once)
        unsigned char *need_d = dmalloc<unsigned char>((size_t)n), *all_d = dmalloc<unsigned char>((size_t)n * size);
        std::vector<unsigned char> need_h((size_t)n), all_h((size_t)n * size);
        KMCB200_CHECK(kmcb200_kmat_need_map(K, need_d));
        d2h(need_h.data(), need_d, (size_t)n);
        sync();
        KMCB200_CHECK(kmcb200_rdv_allgather(rdv(r, size), need_h.data(), (size_t)n, all_h.data()));
        h2d(all_d, all_h.data(), (size_t)n * size);
        KMCB200_CHECK(kmcb200_comm_set_send_masks(R.comm, all_d));
        sync();
        kmcb200_free(ctx(), need_d);
        kmcb200_free(ctx(), all_d);
        KMCB200_CHECK(kmcb200_rdv_barrier(rdv(r, size)));  // every rank's masks are in place before the first exchange
    }
    gpubuf.K_distributed = new Distributed_matrix(K, n, kmc_comm.counts_K, kmc_comm.displs_K, kmc_comm.comm_K);
    int self = r;
    gpubuf.K_p_distributed = new Distributed_vector(n, kmc_comm.counts_K, kmc_comm.displs_K, 1, &self, kmc_comm.comm_K);
}
// src/gpu_solvers.h:54: the CB-edge solve reuses the K sparsity here
inline void initialize_sparsity_CB(GPUBuffers &, int, const double, int) {}
// src/gpu_solvers.h:149-153
inline void update_charge_gpu(ELEMENT *d_site_element, int *d_site_charge, int *d_neigh_idx, int N, int nn,
                              const ELEMENT *d_metals, const int num_metals, const int *count, const int *displ,
                              MPI_Comm &comm) {
    (void)count; (void)displ; (void)comm;
    kmcb200::Runtime &r = kmcb200::rt();
    std::vector<int> metals(num_metals);
    if (r.gpubuf && d_metals == r.gpubuf->metal_types && (int)r.metals_h.size() == num_metals) metals = r.metals_h;
    else { kmcb200::d2h(metals.data(), (const int *)d_metals, num_metals); kmcb200::sync(); }
    KMCB200_CHECK(kmcb200_update_charge(kmcb200::ctx(), (const int *)d_site_element, d_site_charge, d_neigh_idx, N, nn,
                                        metals.data(), num_metals, 0, N));  // all rows: charges are replicated
}
// src/gpu_solvers.h:162-164
inline void background_potential_gpu_sparse(hipblasHandle_t handle_cublas, hipsolverDnHandle_t handle, GPUBuffers &gpubuf,
                                            const int N, const int N_left_tot, const int N_right_tot, const double d_Vd,
                                            const int pbc, const double d_high_G, const double d_low_G,
                                            const double nn_dist, const int num_metals, int kmc_step_count) {
    (void)handle_cublas; (void)handle; (void)pbc; (void)nn_dist; (void)kmc_step_count;
    kmcb200::Runtime &r = kmcb200::rt();
    KMCB200_CHECK(kmcb200_background_potential(kmcb200::ctx(), gpubuf.K_distributed->kmat, N, N_left_tot, N_right_tot,
                                               (const int *)gpubuf.site_element, gpubuf.site_charge, r.metals_h.data(),
                                               num_metals, d_Vd, d_high_G, d_low_G, gpubuf.site_potential_boundary,
                                               &r.last_cg_iterations));
    // every rank needs the whole boundary potential (rate list and events are replicated): NVLink all-gather of the row
    // blocks instead of the reference's MPI_Gatherv to the root (src/kmc_main.cpp:367-384) + later MPI_Bcast
    if (r.comm)
        KMCB200_CHECK(kmcb200_comm_allgather(r.comm, gpubuf.site_potential_boundary + N_left_tot,
                                             gpubuf.K_distributed->counts, gpubuf.K_distributed->displacements));
}
// src/gpu_solvers.h:173-178.  sigma / k / lattice are DEVICE pointers (src/gpu_buffers.h:130-134).  The membership test of
// the reference's cutoff list (element in {d, Od, V, O}, src/neighbor_lists_gpu.cu:96,123) is evaluated inline from the
// process's site_element array, so cutoff_window / cutoff_idx / N_cutoff are not read.
inline void poisson_gridless_gpu(const int num_atoms_contact, const int pbc, const int N, const double *lattice,
                                 const double *sigma, const double *k, const double *posx, const double *posy,
                                 const double *posz, const int *site_charge, double *site_potential_charge, const int rank,
                                 const int size, const int *count, const int *displ, const int *cutoff_window,
                                 const int *cutoff_idx, const int N_cutoff) {
    (void)num_atoms_contact; (void)pbc; (void)lattice; (void)size; (void)cutoff_window; (void)cutoff_idx; (void)N_cutoff;
    kmcb200::Runtime &r = kmcb200::rt();
    if (!r.gpubuf) { std::fprintf(stderr, "kmc_b200: poisson_gridless_gpu before GPUBuffers was constructed\n"); std::exit(1); }
    double sk[2];
    kmcb200::d2h(&sk[0], sigma, 1); kmcb200::d2h(&sk[1], k, 1); kmcb200::sync();
    KMCB200_CHECK(kmcb200_poisson_gridless(kmcb200::ctx(), N, posx, posy, posz, (const int *)r.gpubuf->site_element,
                                           site_charge, sk[0], sk[1], 20.0, displ[rank], count[rank], site_potential_charge));
}
// src/gpu_solvers.h:181
inline void sum_and_gather_potential(GPUBuffers &gpubuf, int num_atoms_first_layer, KMC_comm &kmc_comm) {
    (void)num_atoms_first_layer;
    kmcb200::Runtime &r = kmcb200::rt();
    if (r.comm)  // the Coulomb potentials of the other ranks' site rows (src/kmc_main.cpp:411-427 + potential_solver_gpu.cu:1133-1142)
        KMCB200_CHECK(kmcb200_comm_allgather(r.comm, gpubuf.site_potential_charge, kmc_comm.counts_pairwise,
                                             kmc_comm.displs_pairwise));
    KMCB200_CHECK(kmcb200_sum_potential(kmcb200::ctx(), gpubuf.N_, gpubuf.site_potential_charge,
                                        gpubuf.site_potential_boundary));
}
// src/gpu_solvers.h:262 (the reference copies into __constant__ memory, src/kmc_events.cu:566-572)
inline void copytoConstMemory(std::vector<double> E_gen, std::vector<double> E_rec, std::vector<double> E_Vdiff,
                              std::vector<double> E_Odiff) {
    kmcb200::Runtime &r = kmcb200::rt();
    r.E_gen = E_gen; r.E_rec = E_rec; r.E_Vdiff = E_Vdiff; r.E_Odiff = E_Odiff;
    if (r.events)
        KMCB200_CHECK(kmcb200_set_activation_energies(r.events, (int)E_gen.size(), E_gen.data(), E_rec.data(), E_Vdiff.data(),
                                                      E_Odiff.data()));
}
// src/gpu_solvers.h:250-260.  T_bg / freq / sigma / k / lattice are DEVICE pointers.  The host RandomNumberGenerator stays
// the source of truth: its state is uploaded before the device-resident loop and advanced by the two doubles per event the
// loop consumed (src/kmc_events.cu:469,515).
inline double execute_kmc_step_mpi(MPI_Comm comm, const int N, const int *count, const int *displs, const int nn,
                                   const int *neigh_idx, const int *site_layer, const double *lattice, const int pbc,
                                   const double *T_bg, const double *freq, const double *sigma, const double *k,
                                   const double *posx, const double *posy, const double *posz,
                                   const double *site_potential_charge, const double *site_temperature,
                                   ELEMENT *site_element, int *site_charge, RandomNumberGenerator &rng) {
    (void)comm; (void)count; (void)displs; (void)lattice; (void)pbc; (void)site_temperature;
    kmcb200::Runtime &r = kmcb200::rt();
    if (!r.events) {
        KMCB200_CHECK(kmcb200_events_create(kmcb200::ctx(), N, nn, neigh_idx, &r.events));
        if (r.E_gen.empty()) { std::fprintf(stderr, "kmc_b200: execute_kmc_step_mpi before copytoConstMemory\n"); std::exit(1); }
        KMCB200_CHECK(kmcb200_set_activation_energies(r.events, (int)r.E_gen.size(), r.E_gen.data(), r.E_rec.data(),
                                                      r.E_Vdiff.data(), r.E_Odiff.data()));
    }
    double s4[4];
    kmcb200::d2h(&s4[0], T_bg, 1); kmcb200::d2h(&s4[1], freq, 1); kmcb200::d2h(&s4[2], sigma, 1); kmcb200::d2h(&s4[3], k, 1);
    kmcb200::sync();
    unsigned mt[624];
    int pos = 0;
    rng.getState(mt, &pos);
    KMCB200_CHECK(kmcb200_rng_set_state(r.events, mt, pos));
    double event_time = 0.0;
    int n_events = 0;
    KMCB200_CHECK(kmcb200_execute_kmc_step(kmcb200::ctx(), r.events, N, nn, neigh_idx, site_layer, s4[0], s4[1], s4[2], s4[3],
                                           posx, posy, posz, site_potential_charge, (int *)site_element, site_charge, 0,
                                           &event_time, &n_events));
    rng.discardDoubles((unsigned long long)n_events * 2ULL);
    r.last_n_events = n_events;
    return event_time;
}

// ---- Kirchhoff / current chain ---------------------------------------------------------------------------------------
// src/gpu_solvers.h:143-146
inline void update_CB_edge_gpu_sparse(hipblasHandle_t handle_cublas, hipsolverDnHandle_t handle, GPUBuffers &gpubuf,
                                      const int N, const int N_left_tot, const int N_right_tot, const double d_Vd,
                                      const int pbc, const double d_high_G, const double d_low_G, const double nn_dist,
                                      const int num_metals) {
    (void)handle_cublas; (void)handle; (void)pbc; (void)nn_dist;
    kmcb200::Runtime &r = kmcb200::rt();
    int it = 0;
    KMCB200_CHECK(kmcb200_update_CB_edge(kmcb200::ctx(), gpubuf.K_distributed->kmat, N, N_left_tot, N_right_tot,
                                         (const int *)gpubuf.site_element, r.metals_h.data(), num_metals, d_Vd, d_high_G,
                                         d_low_G, gpubuf.site_CB_edge, 50000, &it));
}
// src/gpu_solvers.h:57
inline void initialize_sparsity_T(GPUBuffers &gpubuf, int pbc, const double nn_dist, int num_source_inj, int num_ground_ext,
                                  int num_layers_contact, KMC_comm &kmc_comm) {
    (void)pbc;
    kmcb200::single_rank_only(kmc_comm.size_T, "initialize_sparsity_T");
    kmcb200::Runtime &r = kmcb200::rt();
    if (r.tmat) kmcb200_tmat_destroy(r.tmat);
    KMCB200_CHECK(kmcb200_initialize_sparsity_T(kmcb200::ctx(), gpubuf.N_, (const int *)gpubuf.site_element, gpubuf.site_x,
                                                gpubuf.site_y, gpubuf.site_z, nn_dist, num_source_inj, num_ground_ext,
                                                num_layers_contact, &r.tmat));
    KMCB200_CHECK(kmcb200_tmat_info(r.tmat, &gpubuf.N_atom_, nullptr, nullptr, nullptr));
    gpubuf.N_sub_ = gpubuf.N_atom_ + 1;
}
// src/gpu_solvers.h:212-218: one assembly + one split-sparse solve + I_macro per call (the reference's body is a timing
// harness around exactly these steps, src/current_solver_gpu.cu:1494-1801)
inline void update_power_gpu_sparse_dist(hipblasHandle_t handle, hipsolverDnHandle_t handle_cusolver, GPUBuffers &gpubuf,
                                         const int num_source_inj, const int num_ground_ext, const int num_layers_contact,
                                         const double Vd, const double high_G, const double low_G, const double loop_G,
                                         const double G0, const double tol, const double nn_dist, const double m_e,
                                         const double V0, int num_metals, double *imacro, const bool solve_heating_local,
                                         const bool solve_heating_global, const double alpha_disp) {
    (void)handle; (void)handle_cusolver; (void)num_source_inj; (void)num_ground_ext; (void)num_layers_contact; (void)tol;
    (void)nn_dist; (void)solve_heating_local; (void)solve_heating_global; (void)alpha_disp;
    kmcb200::Runtime &r = kmcb200::rt();
    if (!r.tmat) { std::fprintf(stderr, "kmc_b200: update_power_gpu_sparse_dist before initialize_sparsity_T\n"); std::exit(1); }
    KMCB200_CHECK(kmcb200_update_power_sparse(kmcb200::ctx(), r.tmat, (const int *)gpubuf.site_element, gpubuf.site_charge,
                                              gpubuf.site_CB_edge, r.metals_h.data(), num_metals, Vd, high_G, low_G, loop_G,
                                              G0, m_e, V0, gpubuf.atom_virtual_potentials, imacro, &r.last_T_iterations));
}

// src/gpu_solvers.h:229-233 (heat_solver_gpu.cu:43-69)
inline void update_temperatureglobal_gpu(const double *site_power, double *T_bg, const int N, const double a_coeff,
                                         const double b_coeff, const double number_steps, const double C_thermal,
                                         const double small_step) {
    KMCB200_CHECK(kmcb200_update_temperature_global(kmcb200::ctx(), site_power, T_bg, N, a_coeff, b_coeff, number_steps,
                                                    C_thermal, small_step));
}

// ---- dist_iterative solver layer ------------------------------------------------------------------------------------
namespace dspmv {
// dist_iterative/dist_spmv.h:22-27: vecAp_local = A * p_distributed.vec_d[0] (halo exchange inside the library)
inline void gpu_packing_cam(Distributed_matrix &A_distributed, Distributed_vector &p_distributed,
                            rocsparse_dnvec_descr &vecAp_local, hipStream_t &default_stream,
                            rocsparse_handle &default_rocsparseHandle) {
    (void)default_stream; (void)default_rocsparseHandle;
    KMCB200_CHECK(kmcb200_spmv(kmcb200::ctx(), A_distributed.kmat, p_distributed.vec_d[0], vecAp_local.values));
}
inline void gpu_packing(Distributed_matrix &A, Distributed_vector &p, rocsparse_dnvec_descr &v, hipStream_t &s,
                        rocsparse_handle &h) {
    gpu_packing_cam(A, p, v, s, h);
}
}  // namespace dspmv
namespace iterative_solver {
// dist_iterative/dist_conjugate_gradient.h:33-47.  The SpMV template argument selects the reference's exchange algorithm;
// here the SpMV, its halo exchange and the dot products are fused kernels inside kmcb200_pcg_jacobi.
template <void (*distributed_spmv)(Distributed_matrix &, Distributed_vector &, rocsparse_dnvec_descr &, hipStream_t &,
                                   rocsparse_handle &)>
inline void conjugate_gradient_jacobi(Distributed_matrix &A_distributed, Distributed_vector &p_distributed, double *r_local_d,
                                      double *x_local_d, double *diag_inv_local_d, double relative_tolerance,
                                      int max_iterations, MPI_Comm comm) {
    (void)p_distributed; (void)comm;
    int it = 0;
    KMCB200_CHECK(kmcb200_pcg_jacobi(kmcb200::ctx(), A_distributed.kmat, r_local_d, x_local_d, diag_inv_local_d,
                                     relative_tolerance, max_iterations, &it));
    kmcb200::rt().last_cg_iterations = it;
    if (A_distributed.rank == 0)  // dist_conjugate_gradient.cpp:272-274
        std::printf("iteration K = %d\n", it + 1);
}
}  // namespace iterative_solver

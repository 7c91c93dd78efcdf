/*
 * kmc_b200.h -- C ABI of libkmc_b200.so: the B200-native (sm_100a) field-solve + event-selection
 * hot path of DeviceKMC.  Plain C: opaque handles, raw DEVICE pointers (unless a parameter is
 * marked "host"), POD scalars.  Every function returns 0 on success or a negative KMCB200_E_* code;
 * kmcb200_last_error() returns a human readable message for the calling thread's last failure.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference
 * repository root, i.e. src/gpu_solvers.h is the reference's extern "C" header).
 * The reference passes C++ references (GPUBuffers&, KMC_comm&, MPI_Comm&, RandomNumberGenerator&)
 * through its extern "C" block; this ABI flattens those objects into pointers + sizes.  The header
 * include/gpu_solvers_b200.hpp restores the reference's names and argument lists on top of it.
 *
 * All work is enqueued on the context's stream; calls return without a host synchronisation unless
 * they hand a host scalar back (documented per function).
 *
 * Process model: one process per GPU, as the reference runs one MPI rank per GCD.  A context is bound to
 * the device given to kmcb200_create(), which also makes that device current; calls on a context must be
 * made with its device current and from one host thread at a time.
 */
#ifndef KMC_B200_H
#define KMC_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KMCB200_VERSION 100

/* error codes */
#define KMCB200_OK 0
#define KMCB200_E_CUDA (-1)      /* a CUDA runtime call failed            */
#define KMCB200_E_ARG (-2)       /* invalid argument                       */
#define KMCB200_E_NOGPU (-3)     /* no usable CUDA device                  */
#define KMCB200_E_CAPACITY (-4)  /* a fixed capacity was exceeded          */
#define KMCB200_E_IO (-5)        /* file could not be read / parsed        */
#define KMCB200_E_COMM (-6)      /* multi-GPU exchange not initialised     */

/* ELEMENT / EVENTTYPE values: reference src/utils.h:37-60 */
enum { KMCB200_DEFECT = 0, KMCB200_OXYGEN_DEFECT = 1, KMCB200_VACANCY = 2, KMCB200_O = 3, KMCB200_Hf = 4,
       KMCB200_Ni = 5, KMCB200_Ti = 6, KMCB200_Pt = 7, KMCB200_N = 8, KMCB200_NULL_ELEMENT = 9 };
enum { KMCB200_VACANCY_GENERATION = 0, KMCB200_VACANCY_RECOMBINATION = 1, KMCB200_VACANCY_DIFFUSION = 2,
       KMCB200_ION_DIFFUSION = 3, KMCB200_NULL_EVENT = 4 };

#define KMCB200_MAX_LAYERS 5   /* reference src/kmc_events.cu:8  (MAX_NUM_LAYERS) */
#define KMCB200_MAX_METALS 4
#define KMCB200_CHUNK 256      /* rows per deterministic dot-product chunk  */
#define KMCB200_SPMV_LANES 8   /* lanes per CSR row in the SpMV row reduction */
#define KMCB200_DOT_GROUP 64   /* chunks per group of the two-level dot combine (systems with more than 256 chunks) */

typedef struct kmcb200_ctx kmcb200_ctx;        /* device + stream + scratch                         */
typedef struct kmcb200_kmat kmcb200_kmat;      /* K matrix: reference Distributed_matrix + contact CSR */
typedef struct kmcb200_events kmcb200_events;  /* event list workspace + KMC RNG                     */
typedef struct kmcb200_comm kmcb200_comm;      /* row-sharded solver exchange plan (peer memory)     */

const char *kmcb200_last_error(void);
int kmcb200_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py reports the delta) */
long long kmcb200_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Context.  Replaces the per-rank hipSetDevice + handle creation in reference src/kmc_main.cpp:72-101,
 * 241-245.  stream: a cudaStream_t created by the caller (e.g. torch's current stream); NULL selects the
 * CUDA legacy default stream. */
int kmcb200_create(kmcb200_ctx **ctx_out, int device_ordinal, void *stream);
int kmcb200_destroy(kmcb200_ctx *ctx);
int kmcb200_set_stream(kmcb200_ctx *ctx, void *stream);
int kmcb200_synchronize(kmcb200_ctx *ctx);
int kmcb200_device_info(kmcb200_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem);

/* measurement helper: FP64 FMA peak (TFLOP/s, 2 flops per FMA) of the context's device, best of 3 timed launches:
 * the roofline denominator of the Coulomb sum (bench.py) */
int kmcb200_fp64_peak(kmcb200_ctx *ctx, double *tflops_host);

/* Device memory helpers for hosts that do not bring their own allocator
 * (reference: hipMalloc/hipMemcpy in src/gpu_buffers.h:93-160, src/gpu_buffers.cpp:10-55). */
int kmcb200_malloc(kmcb200_ctx *ctx, void **dptr_out, size_t bytes);
int kmcb200_free(kmcb200_ctx *ctx, void *dptr);
int kmcb200_memcpy_h2d(kmcb200_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes); /* async on stream */
int kmcb200_memcpy_d2h(kmcb200_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes); /* async on stream */
int kmcb200_memcpy_d2d(kmcb200_ctx *ctx, void *dst_dev, const void *src_dev, size_t bytes); /* async on stream */
int kmcb200_memset(kmcb200_ctx *ctx, void *dst_dev, int value, size_t bytes);
int kmcb200_host_alloc_pinned(void **hptr_out, size_t bytes);
int kmcb200_host_free_pinned(void *hptr);

/* ------------------------------------------------------------------------------------------------
 * a1.  Neighbour table.  Replaces compute_neighbor_list (src/gpu_solvers.h:43,
 * src/neighbor_lists_gpu.cu:257-290, kernel :55-78).  O(N) cell list instead of the O(N^2) kernel;
 * identical output: for rows [row_start, row_start+row_count) the first `nn` sites j (ascending j) with
 * non-PBC sqrt(dx^2+dy^2+dz^2) < nn_dist, j != i; unused slots = -1.  neigh_out: row_count*nn int32. */
int kmcb200_compute_neighbor_list(kmcb200_ctx *ctx, int N, const double *x, const double *y, const double *z,
                                  double nn_dist, int nn, int row_start, int row_count, int *neigh_out);

/* a2.  Cutoff list.  Replaces compute_cutoff_list (src/gpu_solvers.h:46, src/neighbor_lists_gpu.cu:293-373).
 * The 20 A list is NOT needed by kmcb200_poisson_gridless (it evaluates the membership test inline); these two
 * calls exist so a host that wants gpubuf.N_cutoff_ / gpubuf.cutoff_idx gets bit-identical data.
 * kmcb200_cutoff_size returns (host) the max row count == N_cutoff_ (:340-342) and optionally per-row counts. */
int kmcb200_cutoff_size(kmcb200_ctx *ctx, int N, const int *element, const double *x, const double *y,
                        const double *z, double cutoff_radius, int row_start, int row_count,
                        int *counts_out /* device, row_count, may be NULL */, int *max_count_host);
int kmcb200_cutoff_list(kmcb200_ctx *ctx, int N, const int *element, const double *x, const double *y,
                        const double *z, double cutoff_radius, int max_num_cutoff, int row_start,
                        int row_count, int *cutoff_idx_out /* device, row_count*max_num_cutoff, -1 padded */);

/* ------------------------------------------------------------------------------------------------
 * a3/a4.  K sparsity.  Replaces initialize_sparsity_K (src/gpu_solvers.h:53,
 * src/iterative_solvers_gpu.cu:262-488) and the Distributed_matrix / Distributed_vector objects it
 * creates (dist_iterative/dist_objects.h:66-232).  Rows = interior sites N_left+row_start ..
 * +row_count (this rank's block of the N-N_left-N_right interior rows); columns = ALL interior sites
 * with site_dist(pbc) < nn_dist, diagonal included, ascending, stored as one CSR with GLOBAL interior
 * column ids (the reference's per-neighbour-rank sub-blocks are views of it, see
 * kmcb200_kmat_block_view).  Also builds the left/right contact CSR blocks (:449-474).
 * lattice: host pointer to 3 doubles. */
int kmcb200_initialize_sparsity_K(kmcb200_ctx *ctx, int N, const double *x, const double *y, const double *z,
                                  const double *lattice_host, int pbc, double nn_dist, int N_left,
                                  int N_right, int row_start, int row_count, kmcb200_kmat **kmat_out);
/* Wrap a caller-provided CSR (dist_iterative_test shape: exported A_row_ptr/A_col_indices/A_data) */
int kmcb200_kmat_from_csr(kmcb200_ctx *ctx, int rows, int cols_global, int row_start, const int *row_ptr_dev,
                          const int *col_dev, const double *val_dev, kmcb200_kmat **kmat_out);
int kmcb200_kmat_destroy(kmcb200_kmat *kmat);
/* sizes (host): rows, nnz of interior block, nnz of left/right contact blocks */
int kmcb200_kmat_info(kmcb200_kmat *kmat, int *rows, long long *nnz, long long *left_nnz, long long *right_nnz);
/* device pointers owned by kmat (valid until destroy) */
int kmcb200_kmat_pointers(kmcb200_kmat *kmat, int **row_ptr, int **col, double **val, int **left_row_ptr,
                          int **left_col, int **right_row_ptr, int **right_col, double **inv_diag, double **rhs);
/* Reference layout view: CSR of the sub-block whose columns are interior rows [col_start, col_start+col_count)
 * with block-local column ids -- what Distributed_matrix::row_ptr_d[k]/col_indices_d[k] hold for the neighbour
 * owning that range (src/iterative_solvers_gpu.cu:392-412).  row_ptr_out: rows+1, col_out: caller sized via the
 * returned nnz (call first with col_out = NULL). */
int kmcb200_kmat_block_view(kmcb200_kmat *kmat, int col_start, int col_count, int *row_ptr_out, int *col_out,
                            long long *nnz_host);

/* ------------------------------------------------------------------------------------------------
 * a5.  Replaces update_charge_gpu (src/gpu_solvers.h:149-153, src/potential_solver_gpu.cu:12-85).
 * neigh is the table of rows [row_start, row_start+row_count) (row-local, like the reference's per-rank
 * neigh_idx).  metals: host array.  No collective: charges are replicated. */
int kmcb200_update_charge(kmcb200_ctx *ctx, const int *element, int *charge, const int *neigh, int N, int nn,
                          const int *metals_host, int num_metals, int row_start, int row_count);

/* ------------------------------------------------------------------------------------------------
 * a6.  K values + diagonal + Jacobi inverse diagonal + rhs in ONE fused pass.  Replaces the 7 kernels and
 * 5 memsets of background_potential_gpu_sparse (src/potential_solver_gpu.cu:893-1029). */
int kmcb200_assemble_K(kmcb200_ctx *ctx, kmcb200_kmat *kmat, int N, int N_left, int N_right, const int *element,
                       const int *charge, const int *metals_host, int num_metals, double Vd, double high_G,
                       double low_G);

/* a7.  Jacobi-PCG.  Replaces iterative_solver::conjugate_gradient_jacobi<dspmv::gpu_packing_cam>
 * (dist_iterative/dist_conjugate_gradient.h:33-47, .cpp:149-276) and dspmv::gpu_packing_cam
 * (dist_iterative/dist_spmv.h:22-27).  r_local: in rhs / out residual; x_local: in warm start / out solution;
 * both this rank's rows.  iterations_host receives the iteration count (host sync). */
int kmcb200_pcg_jacobi(kmcb200_ctx *ctx, kmcb200_kmat *kmat, double *r_local, double *x_local,
                       const double *diag_inv_local, double relative_tolerance, int max_iterations,
                       int *iterations_host);
/* y = A x (x: this rank's rows; halo handled internally).  dspmv::gpu_packing_cam equivalent.  Row-sharded use
 * requires a STRUCTURALLY SYMMETRIC matrix (like the reference's Distributed_matrix, dist_objects.h:66 "assumes that the
 * matrix is symmetric"): the halo flags that order the reuse of the double-buffered p vector rely on "I send to q <=> I
 * receive from q".  kmcb200_initialize_sparsity_K matrices are symmetric by construction. */
int kmcb200_spmv(kmcb200_ctx *ctx, kmcb200_kmat *kmat, const double *x_local, double *y_local);
/* y = A x and x.(A x) in one pass (the fused SpMV + p.Ap kernel of the PCG iteration); single rank.  dot_host may be
 * NULL (no host sync; the scalar stays on the device). */
int kmcb200_spmv_dot(kmcb200_ctx *ctx, kmcb200_kmat *kmat, const double *x_local, double *y_local, double *dot_host);
/* deterministic dot product of the summation spec (hipblasDdot + MPI_Allreduce replacement) */
int kmcb200_dot(kmcb200_ctx *ctx, const double *u, const double *v, long long n, double *result_host);

/* a6+a7.  Replaces background_potential_gpu_sparse (src/gpu_solvers.h:162-164,
 * src/potential_solver_gpu.cu:846-1128): assemble, then PCG with tol = 1e-14*N_interface, max_it 10000,
 * warm start and result in site_potential_boundary[N_left + row_start ...]. */
int kmcb200_background_potential(kmcb200_ctx *ctx, kmcb200_kmat *kmat, int N, int N_left, int N_right,
                                 const int *element, const int *charge, const int *metals_host, int num_metals,
                                 double Vd, double high_G, double low_G, double *site_potential_boundary,
                                 int *iterations_host);

/* ------------------------------------------------------------------------------------------------
 * (e) Multi-GPU: one process per GPU, interior rows sharded in contiguous blocks whose boundaries are multiples of
 * 256 (kmcb200_partition_aligned).  Replaces the GPU-aware MPI traffic of the reference's distributed PCG
 * (dist_iterative/dist_spmv_gpu_packing.cpp:106-228 halo exchange, dist_conjugate_gradient.cpp:188,213,241,265
 * Allreduce) by remote stores over NVLink peer memory issued by the producing kernels.  Bootstrap:
 *   1. every rank: kmcb200_comm_create, kmcb200_comm_ipc_handle  -> all-gather the 64-byte handles out of band
 *   2. every rank: kmcb200_comm_open_peers(all handles)
 *   3. every rank: kmcb200_initialize_sparsity_K(row_start, row_count of this rank), kmcb200_kmat_attach_comm,
 *      kmcb200_kmat_need_map -> all-gather the n_global-byte maps out of band -> kmcb200_comm_set_send_masks
 * afterwards kmcb200_pcg_jacobi / kmcb200_spmv / kmcb200_background_potential work on the sharded matrix and return
 * bit-identical results on every rank and for every rank count. */
/* interior non-zeros per interior row, for all n rows (device int32[n]): input of an nnz-balanced partition */
int kmcb200_sparsity_K_row_counts(kmcb200_ctx *ctx, int N, const double *x, const double *y, const double *z,
                                  const double *lattice_host, int pbc, double nn_dist, int N_left, int N_right,
                                  int *row_nnz_dev);
int kmcb200_comm_create(kmcb200_ctx *ctx, int rank, int size, int n_global_rows, const int *counts_host,
                        const int *displs_host, kmcb200_comm **comm_out);
/* same, with a staging region of gather_capacity doubles for kmcb200_comm_allgather (>= the largest slice this rank
 * will contribute) */
int kmcb200_comm_create_ex(kmcb200_ctx *ctx, int rank, int size, int n_global_rows, const int *counts_host,
                           const int *displs_host, long long gather_capacity, kmcb200_comm **comm_out);
int kmcb200_comm_destroy(kmcb200_comm *comm);
/* In-place all-gather of row slices of a device vector over NVLink peer memory: on entry vec[displs[rank] ..
 * + counts[rank]) is this rank's slice, on return every rank holds all slices.  Replaces the MPI_Gatherv + MPI_Bcast of
 * the potentials (src/kmc_main.cpp:367-384,411-427, src/potential_solver_gpu.cu:1133-1142).  Host sync. */
int kmcb200_comm_allgather(kmcb200_comm *comm, double *vec_dev, const int *counts_host, const int *displs_host);
/* Out-of-band rendezvous of the ranks of ONE node through a shared directory (bootstrap: IPC handles, need maps,
 * barriers) for hosts without MPI / torch.distributed.  allgather: `all` receives size * bytes. */
typedef struct kmcb200_rdv kmcb200_rdv;
int kmcb200_rdv_open(const char *dir, int rank, int size, kmcb200_rdv **rdv_out);
int kmcb200_rdv_allgather(kmcb200_rdv *rdv, const void *mine_host, size_t bytes, void *all_host);
int kmcb200_rdv_barrier(kmcb200_rdv *rdv);
int kmcb200_rdv_close(kmcb200_rdv *rdv);
int kmcb200_comm_ipc_handle(kmcb200_comm *comm, void *handle64_host);
int kmcb200_comm_open_peers(kmcb200_comm *comm, const void *handles_host /* size * 64 bytes */);
int kmcb200_kmat_attach_comm(kmcb200_kmat *kmat, kmcb200_comm *comm);
int kmcb200_kmat_need_map(kmcb200_kmat *kmat, unsigned char *need_dev /* n_global bytes, device */);
int kmcb200_comm_set_send_masks(kmcb200_comm *comm, const unsigned char *all_need_dev /* size * n_global, device */);
int kmcb200_comm_info(kmcb200_comm *comm, int *rank, int *size, unsigned *recv_mask, long long *arena_bytes);

/* ------------------------------------------------------------------------------------------------
 * a8.  Replaces poisson_gridless_gpu (src/gpu_solvers.h:173-178, src/potential_solver_gpu.cu:1525-1564,
 * 1620-1655).  Overwrites site_potential_charge[row_start .. row_start+row_count). */
int kmcb200_poisson_gridless(kmcb200_ctx *ctx, int N, const double *x, const double *y, const double *z,
                             const int *element, const int *charge, double sigma, double k,
                             double cutoff_radius, int row_start, int row_count, double *site_potential_charge);
/* last poisson call (host sync): number of charged sources, (i,j) distance tests, pairs inside the cutoff (the ones
 * that evaluate erfc / division).  Any pointer may be NULL. */
int kmcb200_poisson_stats(kmcb200_ctx *ctx, long long *num_charged, long long *pair_tests, long long *pairs_in_range);
/* Host-only helper of the pair sum: the largest double d2max with sqrt(d2max) < cutoff (IEEE correctly rounded sqrt), so that
 * the kernel's squared-distance test  d2 <= d2max  selects exactly the pairs of the reference's  sqrt(d2) < cutoff
 * (src/potential_solver_gpu.cu:1549-1551).  Needs no device. */
double kmcb200_cutoff_d2max(double cutoff);

/* a9.  Replaces the kernel of sum_and_gather_potential (src/gpu_solvers.h:181,
 * src/potential_solver_gpu.cu:832-843,1130-1151): site_potential_charge += site_potential_boundary. */
int kmcb200_sum_potential(kmcb200_ctx *ctx, int N, double *site_potential_charge,
                          const double *site_potential_boundary);

/* ------------------------------------------------------------------------------------------------
 * a10/a11.  Events.  kmcb200_set_activation_energies replaces copytoConstMemory (src/gpu_solvers.h:262,
 * src/kmc_events.cu:566-572).  kmcb200_events_create owns the event list (event_type/event_prob of
 * src/kmc_events.cu:356-361), the hierarchical rate sums, the reverse neighbour index and the device-resident
 * MT19937 (bit-identical stream to the host std::mt19937 of src/random_num.h). */
int kmcb200_events_create(kmcb200_ctx *ctx, int N, int nn, const int *neigh, kmcb200_events **ev_out);
int kmcb200_events_destroy(kmcb200_events *ev);
int kmcb200_set_activation_energies(kmcb200_events *ev, int num_layers, const double *E_gen, const double *E_rec,
                                    const double *E_Vdiff, const double *E_Odiff);
int kmcb200_rng_seed(kmcb200_events *ev, unsigned seed);
int kmcb200_rng_set_state(kmcb200_events *ev, const unsigned *mt624_host, int pos);
int kmcb200_rng_get_state(kmcb200_events *ev, unsigned *mt624_host, int *pos_host);
/* draw n doubles from the device generator (advances it); test hook for the RNG stream parity */
int kmcb200_rng_draw(kmcb200_events *ev, int n, double *out_host);
/* Replaces execute_kmc_step_mpi (src/gpu_solvers.h:250-260, src/kmc_events.cu:333-563): builds the rate list
 * from site_potential_charge (the summed potential), then runs the residence-time loop
 * `while (event_time < 1/freq)` on the device; mutates site_element/site_charge.  max_events <= 0: unlimited.
 * Host sync: returns the last drawn event time and the number of events. */
int kmcb200_execute_kmc_step(kmcb200_ctx *ctx, kmcb200_events *ev, int N, int nn, const int *neigh,
                             const int *site_layer, double T_bg, double freq, double sigma, double k,
                             const double *x, const double *y, const double *z,
                             const double *site_potential_charge, int *site_element, int *site_charge,
                             int max_events, double *event_time_host, int *n_events_host);
/* only the rate list (build_event_list_split, src/kmc_events.cu:130-229); exposes device pointers */
int kmcb200_build_event_list(kmcb200_ctx *ctx, kmcb200_events *ev, int N, int nn, const int *neigh,
                             const int *site_layer, double T_bg, double freq, double sigma, double k,
                             const double *x, const double *y, const double *z,
                             const double *site_potential_charge, const int *site_element,
                             const int *site_charge);
int kmcb200_events_pointers(kmcb200_events *ev, double **event_prob, unsigned char **event_type);
/* event log of the last execute call: rows of (i, j, type, slot); returns up to max_rows rows */
int kmcb200_events_log(kmcb200_events *ev, int max_rows, int *log_host, double *psum_host, int *rows_host);

/* ------------------------------------------------------------------------------------------------
 * a12.  Kirchhoff / current chain (single rank).  The reference's driver of this chain is a timing harness that
 * exit(1)s (src/current_solver_gpu.cu:1449-1801) behind a guard that is never true in its shipped main
 * (src/KMC_comm.h:243); these calls are the chain itself.
 *
 * kmcb200_update_CB_edge replaces update_CB_edge_gpu_sparse (src/gpu_solvers.h:143-146,
 * src/potential_solver_gpu.cu:673-772): Laplace-type solve on the K sparsity for the conduction-band edge of every site
 * [J]; contacts fixed to +-Vd/2.  site_CB_edge: N doubles (interior: initial guess in).  K: the 1-rank K sparsity.
 * max_iterations bounds the CG loop of solve_sparse_CG_Jacobi (the reference only warns after 50000). */
typedef struct kmcb200_tmat kmcb200_tmat; /* T_distributed + tunnel sub-block + atom arrays (src/gpu_buffers.h) */
int kmcb200_update_CB_edge(kmcb200_ctx *ctx, kmcb200_kmat *kmat, int N, int N_left, int N_right, const int *element,
                           const int *metals_host, int num_metals, double Vd, double high_G, double low_G,
                           double *site_CB_edge, int max_iterations, int *iterations_host);
/* Replaces initialize_sparsity_T (src/gpu_solvers.h:57, src/initialize_sparsity_T.cu:948-1153): compacts the atoms
 * (sites that are neither DEFECT nor OXYGEN_DEFECT) and builds the CSR of T_neighbor over N_atom + 1 nodes
 * (0 = extraction, 1 = injection, i >= 2 = atom i - 2; the last atom is the ground node, cut from the graph). */
int kmcb200_initialize_sparsity_T(kmcb200_ctx *ctx, int N, const int *site_element, const double *x, const double *y,
                                  const double *z, double nn_dist, int num_source_inj, int num_ground_ext,
                                  int num_layers_contact, kmcb200_tmat **tmat_out);
int kmcb200_tmat_destroy(kmcb200_tmat *tmat);
int kmcb200_tmat_info(kmcb200_tmat *tmat, int *N_atom, long long *nnz, int *n_tunnel, long long *tunnel_nnz);
/* device pointers owned by tmat (any may be NULL): atom -> site index, T_neighbor CSR, Jacobi M^-1, rhs, tunnel points
 * (atom indices), tunnel CSR (columns = tunnel-point ids), tunnel diagonal */
int kmcb200_tmat_pointers(kmcb200_tmat *tmat, int **atom_ind, int **row_ptr, int **col, double **val, double **inv_diag,
                          double **rhs, int **tunnel_atoms, int **t_row_ptr, int **t_col, double **t_val,
                          double **t_diag);
/* Assembly part of update_power_gpu_sparse_dist (src/gpu_solvers.h:212-218, src/current_solver_gpu.cu:1494-1633):
 * update_atom_arrays, populate_T_dist, update_diagonal_sparse, assemble_sparse_T_submatrix (src/gpu_solvers.h:59-64:
 * tunnel points, WKB tunnel block and its diagonal), preconditioner, rhs = (-loop_G Vd, +loop_G Vd, 0 ...). */
int kmcb200_assemble_T(kmcb200_ctx *ctx, kmcb200_tmat *tmat, const int *site_element, const int *site_charge,
                       const double *site_CB_edge, const int *metals_host, int num_metals, double Vd, double high_G,
                       double low_G, double loop_G, double m_e, double V0);
/* y = T_neighbor x + scatter(T_tunnel gather(x)): dspmv_split_sparse::spmm_split_sparse1
 * (dist_iterative/dist_spmv_split_sparse.cpp:5-79).  x, y: N_atom + 1 doubles. */
int kmcb200_tmat_spmv(kmcb200_ctx *ctx, kmcb200_tmat *tmat, const double *x, double *y);
/* iterative_solver::conjugate_gradient_jacobi_split_sparse (dist_iterative/dist_conjugate_gradient.h:72-94,
 * dist_conjugate_gradient_split_sparse.cpp:18-166) on the assembled T with its Jacobi preconditioner.
 * r: in rhs / out residual; x: warm start in / solution out (N_atom + 1 doubles each). */
int kmcb200_pcg_jacobi_split_sparse(kmcb200_ctx *ctx, kmcb200_tmat *tmat, double *r, double *x,
                                    double relative_tolerance, int max_iterations, int *iterations_host);
/* get_imacro_sparse (src/current_solver_gpu.cu:502-542): sum over the injection row's device columns of
 * T[1][col] * G0 * (V[col] - V[1]) */
int kmcb200_imacro(kmcb200_ctx *ctx, kmcb200_tmat *tmat, const double *virtual_potentials, double G0,
                   double *imacro_host);
/* assemble + solve (tolerance 1e-30 * N_atom, 100 iterations max, src/current_solver_gpu.cu:1455-1456) + I_macro.
 * atom_virtual_potentials: N_atom + 1 doubles, warm start in / solution out (gpubuf.atom_virtual_potentials). */
int kmcb200_update_power_sparse(kmcb200_ctx *ctx, kmcb200_tmat *tmat, const int *site_element, const int *site_charge,
                                const double *site_CB_edge, const int *metals_host, int num_metals, double Vd,
                                double high_G, double low_G, double loop_G, double G0, double m_e, double V0,
                                double *atom_virtual_potentials, double *imacro_host, int *iterations_host);

/* f-4.  Replaces update_temperatureglobal_gpu (src/gpu_solvers.h:229-233, src/heat_solver_gpu.cu:43-69): global
 * temperature recurrence T_bg <- c (1 - a^steps) / (1 - a) + a^steps T_bg with c = b + sum(site_power) / C_thermal *
 * small_step.  T_bg_dev: device scalar, updated in place (host sync). */
int kmcb200_update_temperature_global(kmcb200_ctx *ctx, const double *site_power, double *T_bg_dev, int N,
                                      double a_coeff, double b_coeff, double number_steps, double C_thermal,
                                      double small_step);

/* ------------------------------------------------------------------------------------------------
 * Host model (no GPU needed): the pieces of the reference's host side that feed the path and must be
 * read UNCHANGED: parameters.txt grammar (src/input_parser.cpp:3-399), xyz files (src/utils.cpp:72-98),
 * Device::makeSubstoichiometric (src/Device.cpp:180-211), KMCProcess layer assignment
 * (src/KMCProcess.cpp:17-50, src/structure_input.h), KMC_comm row partition (src/KMC_comm.h:245-263). */
typedef struct {
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
} kmcb200_params;

int kmcb200_parse_parameters(const char *path, kmcb200_params *out);
/* V_switch / t_switch vectors (which = 0 / 1); returns count, copies up to cap values */
int kmcb200_parse_parameter_vector(const char *path, int which, int cap, double *out);
int kmcb200_xyz_count(const char *path);
int kmcb200_read_xyz(const char *path, int cap, int *element, double *x, double *y, double *z);
int kmcb200_make_substoichiometric(int N, int *element, double vacancy_concentration, unsigned rnd_seed);
int kmcb200_num_layers(void);
int kmcb200_layer_table(double *E_gen, double *E_rec, double *E_Vdiff, double *E_Odiff, double *start_x,
                        double *end_x);
int kmcb200_assign_layers(int N, const double *x, int *site_layer);
void kmcb200_partition(int nrows, int nranks, int *counts, int *displs);
/* partition whose boundaries keep multi-GPU dot products bit-identical to 1 GPU: multiples of
 * KMCB200_CHUNK rows for systems of up to 256 chunks, multiples of KMCB200_DOT_GROUP chunks (16 384 rows) above */
void kmcb200_partition_aligned(int nrows, int nranks, int *counts, int *displs);

#ifdef __cplusplus
}
#endif
#endif /* KMC_B200_H */

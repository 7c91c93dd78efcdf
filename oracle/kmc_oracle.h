/*
 * kmc_oracle.h -- CPU ORACLE for the DeviceKMC field-solve + event-selection hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (libkmc_b200.so) never links, loads or calls anything in oracle/.
 *
 * It restates, function by function, what the reference's *GPU* hot path computes
 * (the reference has no CPU implementation of this path), citing the
 * reference file:line each function follows (paths relative to /root/reference).
 *
 * Parity pinning: the oracle is pinned against the reference's own shipped golden run
 * (structures/5nm_device/expected_output): snapshot_init.xyz exactly, the 8 events of
 * the 6 shipped supersteps exactly (snapshot_6.xyz elements), "KMC time is:" lines to
 * 1e-3 relative, potentials to 5e-4 V (the reference's own PCG-order noise floor; see
 * DESIGN.md).  tests/test_oracle_golden.py holds those checks.
 *
 * Where the reference delegates arithmetic to vendor libraries with unspecified
 * summation order (rocSPARSE spmv, hipBLAS ddot, thrust::inclusive_scan) the oracle
 * fixes ONE deterministic association ("summation spec") which the CUDA
 * kernels implement identically, so CG iterates and event choices are bit-comparable.
 */
#ifndef KMC_ORACLE_H
#define KMC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* element / event enums: reference src/utils.h:37-60 */
enum { ORC_DEFECT = 0, ORC_OXYGEN_DEFECT = 1, ORC_VACANCY = 2, ORC_O = 3, ORC_Hf = 4,
       ORC_Ni = 5, ORC_Ti = 6, ORC_Pt = 7, ORC_N = 8, ORC_NULL_ELEMENT = 9 };
enum { ORC_VACANCY_GENERATION = 0, ORC_VACANCY_RECOMBINATION = 1, ORC_VACANCY_DIFFUSION = 2,
       ORC_ION_DIFFUSION = 3, ORC_NULL_EVENT = 4 };

#define ORC_MAX_LAYERS 5
#define ORC_CHUNK 256      /* rows per dot-product chunk (summation spec) */
#define ORC_SPMV_LANES 8   /* virtual lanes per CSR row (summation spec)  */
#define ORC_DOT_GROUP 64   /* chunks per group in the two-level dot combine (problems with more than 256 chunks) */

/* ---- a1: neighbour table (neighbor_lists_gpu.cu:55-78,257-290) ------------------ */
/* rows [row_start,row_start+row_count) ; out is row_count*nn, -1 padded, ascending j,
 * non-PBC distance, i != j, first nn kept.  use_cells=0: O(N^2) brute force exactly as
 * the reference kernel; use_cells=1: cell-list enumeration with the same predicate. */
void orc_neighbor_list(int N, const double *x, const double *y, const double *z,
                       double nn_dist, int nn, int row_start, int row_count,
                       int use_cells, int *neigh_out);

/* ---- a2: cutoff list (neighbor_lists_gpu.cu:80-136,293-373) --------------------- */
void orc_cutoff_count(int N, const int *element, const double *x, const double *y,
                      const double *z, double cutoff, int row_start, int row_count,
                      int *count_out);
void orc_cutoff_list(int N, const int *element, const double *x, const double *y,
                     const double *z, double cutoff, int max_num_cutoff, int row_start,
                     int row_count, int *idx_out /* row_count*max_num_cutoff, -1 pad */);

/* ---- a3: CSR sparsity of one (row block, column block) pair
 *      (iterative_solvers_gpu.cu:96-218: calc_nnz_per_row +
 *       assemble_K_indices_gpu_off_diagonal_block / indices_creation_...) ----------
 * rows  = sites block_start_i .. +block_size_i, cols = sites block_start_j .. +block_size_j,
 * entry when site_dist(pbc) < cutoff (diagonal INCLUDED), col index block-local, ascending.
 * row_ptr has block_size_i+1 entries.  If col_out == NULL only row_ptr is produced.
 * Returns nnz. */
long orc_block_sparsity(const double *x, const double *y, const double *z,
                        const double *lattice, int pbc, double cutoff,
                        int block_size_i, int block_size_j, int block_start_i,
                        int block_start_j, int use_cells, int N_total,
                        int *row_ptr_out, int *col_out);

/* ---- a5: site charges (potential_solver_gpu.cu:12-63) --------------------------- */
void orc_update_charge(int N, int nn, const int *element, int *charge, const int *neigh,
                       const int *metals, int num_metals, int row_start, int row_end);

/* ---- a6: K values, diagonal, Jacobi preconditioner, rhs
 *      (potential_solver_gpu.cu:246-285,323-367,438-454,774-830,846-1029) ---------- */
void orc_assemble_K(int N, int N_left, int N_right, const int *element, const int *charge,
                    const int *metals, int num_metals,
                    const int *row_ptr, const int *col,              /* interior block, n = N-NL-NR rows */
                    const int *left_row_ptr, const int *left_col,    /* cols relative to 0            */
                    const int *right_row_ptr, const int *right_col,  /* cols relative to N_left + n   */
                    double Vd, double high_G, double low_G,
                    double *data_out, double *inv_diag_out, double *rhs_out);

/* ---- summation spec primitives ----------------------------------- */
double orc_dot(const double *u, const double *v, long n);
void orc_spmv(long n, const int *row_ptr, const int *col, const double *data, const double *x,
              double *y, int lanes);

/* ---- a7: Jacobi-PCG (dist_conjugate_gradient.cpp:149-276) ----------------------- */
/* r_io: in = rhs, out = final residual.  x_io: warm start in, solution out.
 * Returns the number of iterations performed (reference's k-1).
 * stats_out[0] = final r.z, stats_out[1] = b.b (may be NULL). */
int orc_pcg_jacobi(long n, const int *row_ptr, const int *col, const double *data,
                   const double *inv_diag, double *r_io, double *x_io,
                   double relative_tolerance, int max_iterations, int lanes, double *stats_out);

/* ---- a8: screened Coulomb sum (potential_solver_gpu.cu:1525-1564; gpu_solvers.h:280-328) */
void orc_coulomb(int N, const double *x, const double *y, const double *z, const int *element,
                 const int *charge, double sigma, double k, double cutoff, int row_start,
                 int row_count, double *pot_out /* indexed by global site id */);

/* same result bit for bit, candidates enumerated through a 20 A cell grid instead of the all-sources loop */
void orc_coulomb_cells(int N, const double *x, const double *y, const double *z, const int *element,
                       const int *charge, double sigma, double k, double cutoff, int row_start,
                       int row_count, double *pot_out);

/* ---- a10: event rates (kmc_events.cu:130-229) ----------------------------------- */
void orc_build_events(int N, int nn, const int *neigh, const int *layer, double T_bg, double freq,
                      double sigma, double k, const double *x, const double *y, const double *z,
                      const double *pot, const int *element, const int *charge,
                      const double *E_gen, const double *E_rec, const double *E_Vdiff,
                      const double *E_Odiff, int row_start, int row_count, int *type_out,
                      double *prob_out);

/* ---- host model: Device::makeSubstoichiometric (Device.cpp:180-211); returns the number of vacancies created */
int orc_make_substoichiometric(int N, int *element, double vacancy_concentration, unsigned rnd_seed);

/* ---- a11: RNG (random_num.h:4-26; libstdc++ mt19937 + uniform_real_distribution) - */
void *orc_rng_create(unsigned seed);
void orc_rng_destroy(void *rng);
double orc_rng_next(void *rng);
/* raw MT19937 state: 624 words + position, for uploading into the CUDA generator */
void orc_rng_get_state(void *rng, unsigned *mt624, int *pos);

/* ---- a10: residence-time event loop (kmc_events.cu:448-516) --------------------- */
/* prob/type are mutated (zeroed slots).  element/charge are mutated by executed events.
 * log_out: max_log rows of 4 ints (i, j, type, slot); psum_out: max_log doubles (Psum before each
 * event).  Returns the number of events executed; *event_time_out = last drawn residence time. */
int orc_event_loop(int N, int nn, const int *neigh, int *type, double *prob, int *element,
                   int *charge, double freq, void *rng, int max_events, int max_log, int *log_out,
                   double *psum_out, double *event_time_out);

/* hierarchical selection primitives exposed for unit tests */
void orc_block_scan_256(const double *v, double *incl);
long orc_select_event(int N, int nn, const double *prob, double number, double *psum_out);

/* ---- whole superstep on 1 rank (kmc_main.cpp:328-540) --------------------------- */
typedef struct {
    int N, nn, N_left, N_right, pbc, num_metals;
    int metals[4];
    double lattice[3];
    double nn_dist, sigma, k, T_bg, freq, high_G, low_G, cutoff_radius, Vd;
    double E_gen[ORC_MAX_LAYERS], E_rec[ORC_MAX_LAYERS], E_Vdiff[ORC_MAX_LAYERS], E_Odiff[ORC_MAX_LAYERS];
    double cg_tol_per_row;   /* 1e-14 (potential_solver_gpu.cu:885) */
    int cg_max_it;           /* 10000 (potential_solver_gpu.cu:886) */
    int spmv_lanes;
} orc_params;

typedef struct {
    int cg_iterations;
    int n_events;
    double event_time;
    double t_charge, t_boundary, t_coulomb, t_events; /* wall seconds */
} orc_step_info;

/* static inputs x,y,z,layer,neigh,CSR; mutable element, charge, pot_boundary (N), pot_total (N) */
void orc_superstep(const orc_params *p, const double *x, const double *y, const double *z,
                   const int *layer, const int *neigh, const int *row_ptr, const int *col,
                   const int *left_row_ptr, const int *left_col, const int *right_row_ptr,
                   const int *right_col, int *element, int *charge, double *pot_boundary,
                   double *pot_total, void *rng, int max_log, int *log_out, orc_step_info *info);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif

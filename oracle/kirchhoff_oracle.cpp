// kirchhoff_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY) for the Kirchhoff / current chain.
//
// Restates the reference's GPU "sparse_dist" current solver, function by function (citations relative to
// /root/reference):
//   atoms compaction            src/current_solver_gpu.cu:1341-1365 (update_atom_arrays), gpu_solvers.h:330-336 (is_defect)
//   T_neighbor sparsity         src/initialize_sparsity_T.cu:10-105 (calc_nnz_per_row_T), :108-210 (assemble_T_col_indices)
//   T_neighbor values           src/current_solver_gpu.cu:1051-1238 (populate_T_dist)
//   T_neighbor diagonal         src/current_solver_gpu.cu:1279-1321,1405-1426 (calc_diagonal_T, insert_diag_T)
//   tunnel points               src/initialize_sparsity_T.cu:618-654 (get_is_tunnel_mpi), :773 (copy_if is_not_zero)
//   tunnel sparsity             src/initialize_sparsity_T.cu:212-290 (calc_nnz_per_row_tunnel), :293-375
//   tunnel values (WKB)         src/initialize_sparsity_T.cu:497-614 (populate_T_tunnel_dist2)
//   tunnel diagonal             src/initialize_sparsity_T.cu:669-689 (calc_diagonal_T_tunnel)
//   preconditioner, rhs         src/current_solver_gpu.cu:1323-1339,1613-1633
//   split-sparse Jacobi-PCG     dist_iterative/dist_conjugate_gradient_split_sparse.cpp:18-166,
//                               dist_iterative/dist_spmv_split_sparse.cpp:5-79 (spmm_split_sparse1)
//   macroscopic current         src/current_solver_gpu.cu:502-542 (get_imacro_sparse), :2036-2049 (its live call site)
//   CB edge (Laplace) solve     src/potential_solver_gpu.cu:287-319,370-436,575-772, src/iterative_solvers_gpu.cu:716-887
//
// PARITY PINNING: the reference ships no golden data for this chain (it is unreachable in the shipped main,
// src/KMC_comm.h:243, and its only driver is a timing harness that exit(1)s, src/current_solver_gpu.cu:1801): parity of
// the CUDA path is against THIS restatement, which is pinned by the reference's own acceptance criteria for it:
// the algebraic invariants of postprocessing/test_matrices.py:38-48 (symmetry, diagonal = -sum of off-diagonals) and
// "split == monolithic" of dist_iterative_test/main_test_cg_split.cpp:1267,1433-1441 (tests/test_kirchhoff_oracle.py).
//
// Summation orders the reference leaves to rocSPARSE / hipBLAS follow the summation spec of DESIGN.md section 4
// (the same spmv_spec / dot_spec as the K solve); everything the reference computes sequentially per thread is
// sequential here.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kmc_oracle.h"

namespace {
const double eV_to_J = 1.60217663e-19;  // initialize_sparsity_T.cu:5
const double h_bar = 1.054571817e-34;   // initialize_sparsity_T.cu:6

inline double dist_nopbc(double x1, double y1, double z1, double x2, double y2, double z2) {  // gpu_solvers.h:280-285
    double dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
    return std::sqrt(dx * dx + dy * dy + dz * dz);
}
// exp / x^1.5 of the WKB coefficients: the summation-spec idea applied to the two transcendental calls of
// populate_T_tunnel_dist2 (the reference leaves them to the device math library).  Only +, *, fma, sqrt, ldexp: the CUDA
// kernels evaluate the identical operation sequence (csrc/common.cuh), so the tunnel block is bit-comparable; both agree
// with libm's exp / pow(x, 1.5) to <= 2 ulp (tests/test_kirchhoff_oracle.py).
inline double det_exp(double x) {
    if (!(x > -745.2)) return 0.0;
    if (x > 709.7) return HUGE_VAL;
    const double k = std::nearbyint(x * 1.4426950408889634);
    double r = std::fma(-k, 0.693147180369123816490e+00, x);
    r = std::fma(-k, 1.90821492927058770002e-10, r);
    double p = 1.0 / 6227020800.0;
    p = std::fma(p, r, 1.0 / 479001600.0);
    p = std::fma(p, r, 1.0 / 39916800.0);
    p = std::fma(p, r, 1.0 / 3628800.0);
    p = std::fma(p, r, 1.0 / 362880.0);
    p = std::fma(p, r, 1.0 / 40320.0);
    p = std::fma(p, r, 1.0 / 5040.0);
    p = std::fma(p, r, 1.0 / 720.0);
    p = std::fma(p, r, 1.0 / 120.0);
    p = std::fma(p, r, 1.0 / 24.0);
    p = std::fma(p, r, 1.0 / 6.0);
    p = std::fma(p, r, 0.5);
    p = std::fma(p, r, 1.0);
    p = std::fma(p, r, 1.0);
    return std::ldexp(p, (int)k);
}
inline double det_pow15(double x) { return x * std::sqrt(x); }

inline bool is_metal(int el, const int *metals, int num_metals) {
    for (int m = 0; m < num_metals; ++m)
        if (metals[m] == el) return true;
    return false;
}
}  // namespace

extern "C" {

// test hooks for the deterministic math routines
double orc_det_exp(double x) { return det_exp(x); }
double orc_det_pow15(double x) { return det_pow15(x); }

// update_atom_arrays: the sites that are neither DEFECT nor OXYGEN_DEFECT, in site order.  atom_ind_out may be NULL.
int orc_atoms_compact(int N, const int *element, int *atom_ind_out) {
    int n = 0;
    for (int i = 0; i < N; ++i)
        if (element[i] != ORC_DEFECT && element[i] != ORC_OXYGEN_DEFECT) {
            if (atom_ind_out) atom_ind_out[n] = i;
            ++n;
        }
    return n;
}

// T_neighbor sparsity of matrix rows [row_start, row_start + row_count), all Nsub = N_atom + 1 columns, ascending.
// Node 0 = extraction, node 1 = injection, node i >= 2 = atom i - 2; the last atom (ground) is cut from the graph.
// col_out == NULL: row_ptr only.  Returns nnz.
long orc_T_sparsity(int N_atom, const double *ax, const double *ay, const double *az, double nn_dist,
                    int num_source_inj, int num_ground_ext, int row_start, int row_count, int *row_ptr, int *col_out) {
    const int Nsub = N_atom + 1;
    // candidate enumeration through a cell grid over the atoms (same predicate afterwards)
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int a = 0; a < N_atom; ++a) {
        lo[0] = std::min(lo[0], ax[a]); hi[0] = std::max(hi[0], ax[a]);
        lo[1] = std::min(lo[1], ay[a]); hi[1] = std::max(hi[1], ay[a]);
        lo[2] = std::min(lo[2], az[a]); hi[2] = std::max(hi[2], az[a]);
    }
    const double h = nn_dist * 1.0001;
    int nc[3];
    for (int d = 0; d < 3; ++d) nc[d] = std::max(1, (int)std::floor((hi[d] - lo[d]) / h) + 1);
    auto cidx = [&](double v, int d) { int c = (int)std::floor((v - lo[d]) / h); return std::min(std::max(c, 0), nc[d] - 1); };
    size_t ncell = (size_t)nc[0] * nc[1] * nc[2];
    std::vector<int> start(ncell + 1, 0), items(N_atom), cid(N_atom);
    for (int a = 0; a < N_atom; ++a) {
        cid[a] = (cidx(ax[a], 0) * nc[1] + cidx(ay[a], 1)) * nc[2] + cidx(az[a], 2);
        start[cid[a] + 1]++;
    }
    for (size_t c = 0; c < ncell; ++c) start[c + 1] += start[c];
    {
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (int a = 0; a < N_atom; ++a) items[fill[cid[a]]++] = a;
    }
    auto row_cols = [&](int i, std::vector<int> &cols) {
        cols.clear();
        if (i == 0) {  // calc_nnz_per_row_T: diagonal, loop connection, extraction terms
            cols.push_back(0);
            cols.push_back(1);
            for (int j = std::max(2, (Nsub + 1) - num_ground_ext + 1); j < Nsub; ++j) cols.push_back(j);
        } else if (i == 1) {  // loop connection, diagonal, injection terms
            cols.push_back(0);
            cols.push_back(1);
            for (int j = 2; j < num_source_inj + 2 && j < Nsub; ++j) cols.push_back(j);
        } else {
            if (i > (Nsub + 1) - num_ground_ext) cols.push_back(0);
            if (i < num_source_inj + 2) cols.push_back(1);
            const int a = i - 2;
            size_t first = cols.size();
            int ca = cidx(ax[a], 0), cb = cidx(ay[a], 1), cc = cidx(az[a], 2);
            for (int da = -1; da <= 1; ++da) for (int db = -1; db <= 1; ++db) for (int dc = -1; dc <= 1; ++dc) {
                int aa = ca + da, bb = cb + db, c2 = cc + dc;
                if (aa < 0 || aa >= nc[0] || bb < 0 || bb >= nc[1] || c2 < 0 || c2 >= nc[2]) continue;
                size_t cell = ((size_t)aa * nc[1] + bb) * nc[2] + c2;
                for (int s = start[cell]; s < start[cell + 1]; ++s) {
                    int b = items[s];
                    if (b + 2 >= Nsub) continue;  // the ground atom is not a column
                    if (b == a || dist_nopbc(ax[a], ay[a], az[a], ax[b], ay[b], az[b]) < nn_dist) cols.push_back(b + 2);
                }
            }
            std::sort(cols.begin() + first, cols.end());
        }
    };
    std::vector<int> cols;
    row_ptr[0] = 0;
    long acc = 0;
    for (int r = 0; r < row_count; ++r) {
        row_cols(row_start + r, cols);
        if (col_out) std::copy(cols.begin(), cols.end(), col_out + acc);
        acc += (long)cols.size();
        row_ptr[r + 1] = (int)acc;
    }
    return acc;
}

// populate_T_dist + update_diagonal_sparse for rows [row_start, row_start+row_count) (col = global matrix column):
// data (CSR values incl. the final diagonal) and diag_out (the neighbour-matrix diagonal, one per row).
void orc_T_values(int N_atom, const double *ax, const double *ay, const double *az, const int *a_element,
                  const int *a_charge, const int *metals, int num_metals, double nn_dist, double high_G, double low_G,
                  double loop_G, int num_source_inj, int num_ground_ext, int row_start, int row_count,
                  const int *row_ptr, const int *col, double *data, double *diag_out) {
    const int Nsub = N_atom + 1;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < row_count; ++r) {
        const int i = row_start + r;
        int diag_slot = -1;
        for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) {
            const int j = col[k];
            double v = 0.0;  // hipMemset of data_d (current_solver_gpu.cu:1384)
            if (i == 0) {
                if (j == 0) v = +high_G; else if (j == 1) v = -loop_G; else v = -high_G;
            }
            if (i == 1) {
                if (j == 0) v = -loop_G;
                if (j > 1) v = -high_G;
            }
            if (i >= 2) {
                if (i == j) {
                    double d = dist_nopbc(ax[i - 2], ay[i - 2], az[i - 2], ax[N_atom - 1], ay[N_atom - 1], az[N_atom - 1]);
                    if (d < nn_dist) v = +high_G;
                }
                if (j == 0 && i > (Nsub + 1) - num_ground_ext) v = -high_G;
                if (j == 1 && i > 1 && i < num_source_inj + 2) v = -high_G;
                if (j >= 2 && j != i) {
                    double d = dist_nopbc(ax[i - 2], ay[i - 2], az[i - 2], ax[j - 2], ay[j - 2], az[j - 2]);
                    if (d < nn_dist) {
                        bool metal1 = is_metal(a_element[i - 2], metals, num_metals);
                        bool metal2 = is_metal(a_element[j - 2], metals, num_metals);
                        bool cv1 = (a_element[i - 2] == ORC_VACANCY) && (a_charge[i - 2] == 0);
                        bool cv2 = (a_element[j - 2] == ORC_VACANCY) && (a_charge[j - 2] == 0);
                        v = ((metal1 && metal2) || (cv1 && cv2)) ? -high_G : -low_G;
                    }
                }
            }
            data[k] = v;
            if (j == i) diag_slot = k;
        }
        // calc_diagonal_T: sequential sum of the off-diagonals in column order; diag = 0 + (-tmp)
        double tmp = 0.0;
        for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k)
            if (col[k] != i) tmp += data[k];
        double diag = 0.0;
        diag += -tmp;
        // insert_diag_T: data[diag] += diag; diag = data[diag]
        if (diag_slot >= 0) {
            data[diag_slot] += diag;
            diag = data[diag_slot];
        }
        diag_out[r] = diag;
    }
}

// get_is_tunnel_mpi + copy_if(is_not_zero): atom indices (ascending) of the tunnel points among atoms 0 .. N_atom-2.
// Atom 0 can never be listed (its index 0 is what is_not_zero drops); returns -1 if it qualifies (the reference would
// then count one point more than it lists).
int orc_tunnel_points(int N_atom, const int *a_element, const double *ax, int *tunnel_atoms_out) {
    int n = 0;
    for (int idx = 0; idx <= N_atom - 2; ++idx) {
        bool yes = a_element[idx] == ORC_VACANCY ||
                   ((a_element[idx] == ORC_Ti || a_element[idx] == ORC_N) && (ax[idx] > -4.2 && ax[idx] < 52.65));
        if (yes) {
            if (idx == 0) return -1;
            if (tunnel_atoms_out) tunnel_atoms_out[n] = idx;
            ++n;
        }
    }
    return n;
}

namespace {
struct TunnelCtx {
    int N_atom, num_layers_contact, num_source_inj, num_ground_ext;
    const double *ax, *ay, *az, *cb;
    const int *el;
    double nn_dist;
};
// the pair predicate shared by calc_nnz_per_row_tunnel / assemble_tunnel_col_indices / populate_T_tunnel_dist2
// (metals are hard-coded to 2 entries by the callers, initialize_sparsity_T.cu:800; tol = eV_to_J * 0.01, :799)
inline bool tunnel_pair(const TunnelCtx &c, int ind_i, int ind_j, const int *metals, int num_metals, bool *contact_to_trap,
                        double *dist_out, double *dE_out) {
    double dist = dist_nopbc(c.ax[ind_i], c.ay[ind_i], c.az[ind_i], c.ax[ind_j], c.ay[ind_j], c.az[ind_j]);
    *dist_out = dist;
    bool v1 = c.el[ind_i] == ORC_VACANCY, v2 = c.el[ind_j] == ORC_VACANCY;
    bool m1 = is_metal(c.el[ind_i], metals, num_metals) && (ind_i > ((c.num_layers_contact - 1) * c.num_source_inj)) &&
              (ind_i < (c.N_atom - (c.num_layers_contact - 1) * c.num_ground_ext));
    bool m2 = is_metal(c.el[ind_j], metals, num_metals) && (ind_j > ((c.num_layers_contact - 1) * c.num_source_inj)) &&
              (ind_j < (c.N_atom - (c.num_layers_contact - 1) * c.num_ground_ext));
    bool tt = v1 && v2, ct = (v1 && m2) || (v2 && m1), cc = m1 && m2;
    double dE = c.cb[ind_i] - c.cb[ind_j];
    *dE_out = dE;
    *contact_to_trap = ct;
    const double tol = eV_to_J * 0.01;
    return (tt || ct || cc) && (std::fabs(dE) > tol);
}
// populate_T_tunnel_dist2: the WKB coefficient of one listed pair (i != j, not neighbours).  The contact-to-trap
// integration loop (:576-590) is restated with an early exit once a term underflows to exactly 0: its exponent
// decreases monotonically with iv, so every later term is exactly 0 and the sum is unchanged (with CB edges in
// joules the loop has a single term anyway: dE = eV_to_J * 0.01 * 1e10 is far larger than any energy window).
inline double tunnel_value(double dist_angstrom, double local_E_drop, bool contact_to_trap, double m_e, double V0,
                           bool *written) {
    double prefac = -(std::sqrt(2 * m_e) / h_bar) * (2.0 / 3.0);
    double dist = (1e-10) * dist_angstrom;
    *written = true;
    if (contact_to_trap) {
        double energy_window = std::fabs(local_E_drop);
        double dV = 0.01;
        double dE = eV_to_J * dV * 10000000000;
        double T = 0.0;
        for (double iv = 0; iv < energy_window; iv += dE) {
            double E1 = eV_to_J * V0 + iv;
            double E2 = E1 - std::fabs(local_E_drop);
            double term = -1.0;
            if (E2 > 0) term = det_exp(prefac * (dist / std::fabs(local_E_drop)) * (det_pow15(E1) - det_pow15(E2)));
            if (E2 < 0) term = det_exp(prefac * (dist / std::fabs(local_E_drop)) * (det_pow15(E1)));
            if (term >= 0.0) {
                T += term;
                if (term == 0.0) break;
            }
        }
        return -T;
    }
    double E1 = eV_to_J * V0;
    double E2 = E1 - std::fabs(local_E_drop);
    if (E2 > 0) return -det_exp(prefac * (dist / std::fabs(E1 - E2)) * (det_pow15(E1) - det_pow15(E2)));
    if (E2 < 0) return -det_exp(prefac * (dist / std::fabs(E1 - E2)) * (det_pow15(E1)));
    *written = false;  // E2 == 0: the reference leaves the (uninitialised) entry untouched; defined as 0 here
    return 0.0;
}
}  // namespace

// Tunnel block for tunnel rows [t_start, t_start + t_count): CSR over the global tunnel-point columns (ascending),
// values, diagonal (calc_diagonal_T_tunnel).  col_out == NULL: row_ptr only.  a_cb = atom_CB_edge [J].
long orc_tunnel_block(int N_atom, const double *ax, const double *ay, const double *az, const int *a_element,
                      const double *a_cb, const int *metals, int num_metals, double nn_dist, int num_layers_contact,
                      int num_source_inj, int num_ground_ext, double m_e, double V0, int n_tunnel,
                      const int *tunnel_atoms, int t_start, int t_count, int *row_ptr, int *col_out, double *data_out,
                      double *diag_out) {
    TunnelCtx c{N_atom, num_layers_contact, num_source_inj, num_ground_ext, ax, ay, az, a_cb, a_element, nn_dist};
    std::vector<int> nnz_row(t_count, 0);
#pragma omp parallel for schedule(dynamic, 16)
    for (int r = 0; r < t_count; ++r) {
        int i = t_start + r, ind_i = tunnel_atoms[i], n = 0;
        for (int j = 0; j < n_tunnel; ++j) {
            int ind_j = tunnel_atoms[j];
            bool ct;
            double dist, dE;
            bool ok = tunnel_pair(c, ind_i, ind_j, metals, num_metals, &ct, &dist, &dE);
            if (i == j) n++;
            if (i != j && dist > nn_dist && ok) n++;
        }
        nnz_row[r] = n;
    }
    row_ptr[0] = 0;
    long acc = 0;
    for (int r = 0; r < t_count; ++r) { acc += nnz_row[r]; row_ptr[r + 1] = (int)acc; }
    if (!col_out) return acc;
#pragma omp parallel for schedule(dynamic, 16)
    for (int r = 0; r < t_count; ++r) {
        int i = t_start + r, ind_i = tunnel_atoms[i];
        int k = row_ptr[r], diag_slot = -1;
        for (int j = 0; j < n_tunnel; ++j) {
            int ind_j = tunnel_atoms[j];
            bool ct;
            double dist, dE;
            bool ok = tunnel_pair(c, ind_i, ind_j, metals, num_metals, &ct, &dist, &dE);
            if (i == j) { col_out[k] = j; data_out[k] = 0.0; diag_slot = k; ++k; }
            if (i != j && dist > nn_dist && ok) {
                col_out[k] = j;
                bool neighbor = (dist < nn_dist) && (i != j);  // populate_T_tunnel_dist2 re-tests (always false here)
                bool written = false;
                data_out[k] = neighbor ? 0.0 : tunnel_value(dist, dE, ct, m_e, V0, &written);
                ++k;
            }
        }
        double tmp = 0.0;  // calc_diagonal_T_tunnel: sequential over the row, skipping the diagonal
        for (int q = row_ptr[r]; q < row_ptr[r + 1]; ++q)
            if (col_out[q] != i) tmp += data_out[q];
        diag_out[r] = -tmp;
        if (diag_slot >= 0) data_out[diag_slot] = -tmp;
    }
    return acc;
}

// y = T_neighbor x + scatter(T_tunnel gather(x))  (spmm_split_sparse1): rows of one rank holding ALL rows.
// tunnel_rows[t] = matrix row of tunnel point t (= tunnel_atoms[t] + 2).
void orc_split_spmv(int n, const int *row_ptr, const int *col, const double *data, int n_tunnel, const int *t_row_ptr,
                    const int *t_col, const double *t_data, const int *tunnel_rows, const double *x, double *y,
                    int lanes, int t_lanes) {
    orc_spmv(n, row_ptr, col, data, x, y, lanes);
    std::vector<double> xs((size_t)std::max(n_tunnel, 1)), ys((size_t)std::max(n_tunnel, 1));
    for (int t = 0; t < n_tunnel; ++t) xs[t] = x[tunnel_rows[t]];  // pack_gpu
    orc_spmv(n_tunnel, t_row_ptr, t_col, t_data, xs.data(), ys.data(), t_lanes);  // tunnel rows are long: 32 lanes per row
    for (int t = 0; t < n_tunnel; ++t) y[tunnel_rows[t]] = y[tunnel_rows[t]] + ys[t];  // unpack_add
}

// conjugate_gradient_jacobi_split_sparse (same update order as the K solve's PCG, dist_..._split_sparse.cpp:52-150)
int orc_pcg_jacobi_split_sparse(int n, const int *row_ptr, const int *col, const double *data, int n_tunnel,
                                const int *t_row_ptr, const int *t_col, const double *t_data, const int *tunnel_rows,
                                const double *inv_diag, double *r, double *x, double tol, int max_it, int lanes,
                                int t_lanes, double *stats) {
    std::vector<double> p((size_t)n), Ap((size_t)n), z((size_t)n);
    double bb = orc_dot(r, r, n);
    orc_split_spmv(n, row_ptr, col, data, n_tunnel, t_row_ptr, t_col, t_data, tunnel_rows, x, Ap.data(), lanes, t_lanes);
    for (int i = 0; i < n; ++i) { r[i] = r[i] - Ap[i]; z[i] = r[i] * inv_diag[i]; }
    double rz = orc_dot(r, z.data(), n), r0 = 0.0;
    int k = 1;
    while (rz / bb > tol * tol && k <= max_it) {
        if (k > 1) {
            double b = rz / r0;
            for (int i = 0; i < n; ++i) { double t = b * p[i]; p[i] = z[i] + t; }
        } else {
            for (int i = 0; i < n; ++i) p[i] = z[i];
        }
        orc_split_spmv(n, row_ptr, col, data, n_tunnel, t_row_ptr, t_col, t_data, tunnel_rows, p.data(), Ap.data(), lanes, t_lanes);
        double pAp = orc_dot(p.data(), Ap.data(), n);
        double a = rz / pAp, na = -a;
        for (int i = 0; i < n; ++i) {
            x[i] = std::fma(a, p[i], x[i]);
            r[i] = std::fma(na, Ap[i], r[i]);
            z[i] = r[i] * inv_diag[i];
        }
        r0 = rz;
        rz = orc_dot(r, z.data(), n);
        k++;
    }
    if (stats) { stats[0] = rz; stats[1] = bb; }
    return k - 1;
}

// get_imacro_sparse on row 1 (injection) of T with m = G0 * virtual potentials: sum over the entries with column >= 2
// of T[1][col] * (m[col] - m[1]).  The reference's tree + atomicAdd order is unspecified: the products are summed with
// the dot-product association of the summation spec, in column order.
double orc_imacro(const int *row_ptr, const int *col, const double *data, const double *virtual_potentials, double G0) {
    std::vector<double> a, b;
    double m1 = virtual_potentials[1] * G0;
    for (int k = row_ptr[1] + 2; k < row_ptr[2]; ++k)
        if (col[k] >= 2) { a.push_back(data[k]); b.push_back(virtual_potentials[col[k]] * G0 - m1); }
    if (a.empty()) return 0.0;
    return orc_dot(a.data(), b.data(), (long)a.size());
}

// update_CB_edge_gpu_sparse: Laplace-type solve on the K sparsity for the conduction-band edge of every site.
// A: off-diagonal -high_G if either site is a metal else -low_G (calc_off_diagonal_A_CB_gpu); diagonal = -(row sum)
// + left + right contact sums (same metal1 || metal2 rule); rhs = left * (Vd/2) + right * (-Vd/2); solved with the
// symmetrically Jacobi-scaled plain CG of solve_sparse_CG_Jacobi (absolute test r.r > tol^2, tol = 1e-14, first test on
// ||r|| itself); contacts fixed to +-Vd/2; everything scaled by eV_to_J.  max_it bounds the loop (the reference only
// warns after 50000 iterations).  site_cb: in = initial guess (interior), out = CB edge [J].  Returns iterations.
int orc_update_CB_edge(int N, int N_left, int N_right, const int *element, const int *metals, int num_metals,
                       const int *row_ptr, const int *col, const int *left_row_ptr, const int *left_col,
                       const int *right_row_ptr, const int *right_col, double Vd, double high_G, double low_G,
                       double *site_cb, int max_it, int lanes) {
    const int n = N - N_left - N_right;
    const long nnz = row_ptr[n];
    std::vector<double> A((size_t)nnz), rhs((size_t)n), dis((size_t)n);
    auto g = [&](int i, int j) { return (is_metal(element[i], metals, num_metals) || is_metal(element[j], metals, num_metals)) ? high_G : low_G; };
    for (int r = 0; r < n; ++r) {
        int i = N_left + r, diag_slot = -1;
        double tmp = 0.0;
        for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) {
            if (col[k] != r) A[k] = -g(i, N_left + col[k]); else { A[k] = 0.0; diag_slot = k; }
            tmp += A[k];
        }
        double left = 0.0, right = 0.0;
        for (int k = left_row_ptr[r]; k < left_row_ptr[r + 1]; ++k) left += g(i, left_col[k]);
        for (int k = right_row_ptr[r]; k < right_row_ptr[r + 1]; ++k) right += g(i, N_left + n + right_col[k]);
        double d = 0.0;
        d -= tmp;       // reduce_rows_into_diag
        d += left;      // add_vector_to_diagonal (left, then right)
        d += right;
        A[diag_slot] = d;
        rhs[r] = left * (Vd / 2) + right * (-Vd / 2);  // calc_rhs_for_A with VL = Vd/2, VR = -Vd/2
        dis[r] = 1.0 / std::sqrt(d);                    // computeDiagonalInvSqrt
    }
    double *y = site_cb + N_left;
    for (int r = 0; r < n; ++r) {
        rhs[r] = rhs[r] * dis[r];                                                        // jacobi_precondition_array(x)
        for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) A[k] = A[k] * dis[r] * dis[col[k]];  // jacobi_precondition_matrix
        y[r] = y[r] / dis[r];                                                            // jacobi_unprecondition_array(y)
    }
    std::vector<double> rr((size_t)n), p((size_t)n), tmpv((size_t)n);
    orc_spmv(n, row_ptr, col, A.data(), y, rr.data(), lanes);   // r = A y
    for (int i = 0; i < n; ++i) { rr[i] = std::fma(-1.0, rhs[i], rr[i]); p[i] = rr[i]; p[i] = -1.0 * p[i]; }
    double h_norm = std::sqrt(orc_dot(rr.data(), rr.data(), n));  // Dnrm2
    const double tol = 1e-14;
    int counter = 0;
    while (h_norm > tol * tol && counter < max_it) {
        double t = orc_dot(rr.data(), rr.data(), n);
        orc_spmv(n, row_ptr, col, A.data(), p.data(), tmpv.data(), lanes);
        double alpha_temp = orc_dot(p.data(), tmpv.data(), n);
        double alpha = t / alpha_temp;
        for (int i = 0; i < n; ++i) { y[i] = std::fma(alpha, p[i], y[i]); rr[i] = std::fma(alpha, tmpv[i], rr[i]); }
        double tnew = orc_dot(rr.data(), rr.data(), n);
        double beta = tnew / t;
        for (int i = 0; i < n; ++i) { double s = p[i] * beta; p[i] = std::fma(-1.0, rr[i], s); }
        h_norm = tnew;  // the reference recomputes r.r (same value)
        counter++;
    }
    for (int r = 0; r < n; ++r) y[r] = y[r] * dis[r];  // jacobi_precondition_array(y)
    for (int i = 0; i < N_left; ++i) site_cb[i] = Vd / 2;
    for (int i = N_left + n; i < N; ++i) site_cb[i] = -Vd / 2;
    for (int i = 0; i < N; ++i) site_cb[i] = site_cb[i] * eV_to_J;  // hipblasDscal
    return counter;
}

// f-4: update_temperatureglobal_gpu (heat_solver_gpu.cu:43-69): returns the new T_bg
double orc_update_temperature_global(const double *site_power, int N, double T_bg, double a_coeff, double b_coeff,
                                     double number_steps, double C_thermal, double small_step) {
    std::vector<double> ones((size_t)N, 1.0);
    double P_tot = orc_dot(site_power, ones.data(), N);
    double c_coeff = b_coeff + P_tot / C_thermal * small_step;
    int step = (int)number_steps;
    return c_coeff * (1.0 - std::pow(a_coeff, (double)step)) / (1.0 - a_coeff) + std::pow(a_coeff, (double)step) * T_bg;
}

}  // extern "C"

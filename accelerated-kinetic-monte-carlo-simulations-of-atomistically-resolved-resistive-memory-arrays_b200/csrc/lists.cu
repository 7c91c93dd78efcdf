// lists.cu -- a1 neighbour table, a2 cutoff list, a3 K sparsity (setup-time kernels).
// Reference: src/neighbor_lists_gpu.cu:55-136,257-373; src/iterative_solvers_gpu.cu:96-218,262-488.
// The reference enumerates all N^2 (resp. N^2/P) pairs; here a cell grid enumerates candidates and the
// identical predicate + ascending-j ordering + caps are applied afterwards.
#include <math.h>

#include "cellgrid.cuh"
#include "kmat.cuh"

namespace {

__global__ void cell_count_kernel(CellGridDev g, const double *__restrict__ x, const double *__restrict__ y,
                                  const double *__restrict__ z, int first, int count, int *__restrict__ cid,
                                  int *__restrict__ counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    int s = first + i;
    int c = g.cell(g.cx(x[s]), g.cy(y[s]), g.cz(z[s]));
    cid[i] = c;
    atomicAdd(counts + c, 1);
}
__global__ void cell_fill_kernel(const int *__restrict__ cid, const int *__restrict__ cell_start,
                                 int *__restrict__ fill, int *__restrict__ items, int first, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    int c = cid[i];
    int pos = atomicAdd(fill + c, 1);
    items[cell_start[c] + pos] = first + i;
}

constexpr int MAX_NN = 64;

// a1: neighbor_lists_gpu.cu:55-78.  One thread per binned site (spatially sorted order); keeps the nn smallest
// qualifying j in ascending order.
__global__ void __launch_bounds__(128) neighbor_kernel(CellGridDev g, const double *__restrict__ x,
                                                      const double *__restrict__ y, const double *__restrict__ z,
                                                      int N, double nn_dist, int nn, int row_start, int row_count,
                                                      int *__restrict__ neigh) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    int i = g.items[t];
    if (i < row_start || i >= row_start + row_count) return;
    double xi = x[i], yi = y[i], zi = z[i];
    int list[MAX_NN];
    int cnt = 0;
    g.for_each_candidate(xi, yi, zi, [&](int j) {
        double d = kmc_dist_nopbc(xi, yi, zi, x[j], y[j], z[j]);
        if (d < nn_dist && i != j) {
            if (cnt < nn) {
                int p = cnt++;
                while (p > 0 && list[p - 1] > j) { list[p] = list[p - 1]; --p; }
                list[p] = j;
            } else if (j < list[nn - 1]) {
                int p = nn - 1;
                while (p > 0 && list[p - 1] > j) { list[p] = list[p - 1]; --p; }
                list[p] = j;
            }
        }
    });
    int *row = neigh + (size_t)(i - row_start) * nn;
    for (int n = 0; n < nn; ++n) row[n] = (n < cnt) ? list[n] : -1;
}

// a2 count: neighbor_lists_gpu.cu:80-104
__global__ void __launch_bounds__(128) cutoff_count_kernel(CellGridDev g, const int *__restrict__ element,
                                                          const double *__restrict__ x, const double *__restrict__ y,
                                                          const double *__restrict__ z, int N, double cutoff,
                                                          int row_start, int row_count, int *__restrict__ counts,
                                                          int *__restrict__ max_count) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int c = 0;
    if (t < N) {
        int i = g.items[t];
        if (i >= row_start && i < row_start + row_count) {
            double xi = x[i], yi = y[i], zi = z[i];
            g.for_each_candidate(xi, yi, zi, [&](int j) {
                double d = kmc_dist_nopbc(xi, yi, zi, x[j], y[j], z[j]);
                if (d < cutoff && i != j && kmc_possibly_charged(element[j])) c++;
            });
            if (counts) counts[i - row_start] = c;
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) c = max(c, __shfl_xor_sync(KMC_FULL_MASK, c, off));
    if ((threadIdx.x & 31) == 0 && c > 0) atomicMax(max_count, c);
}

// a2 list: neighbor_lists_gpu.cu:107-136 (compatibility materialisation; O(N) per row, ordered scan)
__global__ void cutoff_list_kernel(const int *__restrict__ element, const double *__restrict__ x,
                                   const double *__restrict__ y, const double *__restrict__ z, int N, double cutoff,
                                   int max_num_cutoff, int row_start, int row_count, int *__restrict__ out) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= row_count) return;
    int i = idx + row_start;
    double xi = x[i], yi = y[i], zi = z[i];
    int counter = 0;
    int *row = out + (size_t)idx * max_num_cutoff;
    for (int j = 0; j < N; ++j) {
        double d = kmc_dist_nopbc(xi, yi, zi, x[j], y[j], z[j]);
        if (d < cutoff && kmc_possibly_charged(element[j]) && counter < max_num_cutoff && i != j) row[counter++] = j;
    }
    for (; counter < max_num_cutoff; ++counter) row[counter] = -1;
}

constexpr int MAX_ROW_CAND = 128;

// a3 pass 1: calc_nnz_per_row (iterative_solvers_gpu.cu:96-124) for the interior / left / right column ranges.
__global__ void __launch_bounds__(128) ksparsity_count_kernel(CellGridDev g, const double *__restrict__ x,
                                                             const double *__restrict__ y, const double *__restrict__ z,
                                                             int N, int N_left, int N_right, int pbc, double cutoff,
                                                             int row_start, int row_count, int *__restrict__ cnt_int,
                                                             int *__restrict__ cnt_left, int *__restrict__ cnt_right) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= row_count) return;
    int i = N_left + row_start + r;
    double xi = x[i], yi = y[i], zi = z[i];
    int ci = 0, cl = 0, cr = 0;
    int hi = N - N_right;
    g.for_each_candidate(xi, yi, zi, [&](int j) {
        double d = kmc_dist_pbc(xi, yi, zi, x[j], y[j], z[j], g.latty, g.lattz, pbc);
        if (d < cutoff) {
            if (j < N_left) cl++;
            else if (j >= hi) cr++;
            else ci++;
        }
    });
    cnt_int[r] = ci;
    cnt_left[r] = cl;
    cnt_right[r] = cr;
}

// a3 pass 2: assemble_K_indices_gpu_off_diagonal_block (iterative_solvers_gpu.cu:126-157): ascending columns.
__global__ void __launch_bounds__(128) ksparsity_fill_kernel(CellGridDev g, const double *__restrict__ x,
                                                            const double *__restrict__ y, const double *__restrict__ z,
                                                            int N, int N_left, int N_right, int pbc, double cutoff,
                                                            int row_start, int row_count, const int *__restrict__ rp_int,
                                                            const int *__restrict__ rp_left, const int *__restrict__ rp_right,
                                                            int *__restrict__ col_int, int *__restrict__ col_left,
                                                            int *__restrict__ col_right, int *__restrict__ overflow) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= row_count) return;
    int i = N_left + row_start + r;
    double xi = x[i], yi = y[i], zi = z[i];
    int list[MAX_ROW_CAND];
    int cnt = 0;
    bool ovf = false;
    g.for_each_candidate(xi, yi, zi, [&](int j) {
        double d = kmc_dist_pbc(xi, yi, zi, x[j], y[j], z[j], g.latty, g.lattz, pbc);
        if (d < cutoff) {
            if (cnt < MAX_ROW_CAND) {
                int p = cnt++;
                while (p > 0 && list[p - 1] > j) { list[p] = list[p - 1]; --p; }
                list[p] = j;
            } else {
                ovf = true;
            }
        }
    });
    if (ovf) { atomicExch(overflow, 1); return; }
    int hi = N - N_right;
    int oi = rp_int[r], ol = rp_left[r], orr = rp_right[r];
    for (int q = 0; q < cnt; ++q) {
        int j = list[q];
        if (j < N_left) col_left[ol++] = j;
        else if (j >= hi) col_right[orr++] = j - hi;
        else col_int[oi++] = j - N_left;
    }
}

__global__ void block_view_count_kernel(const int *__restrict__ row_ptr, const int *__restrict__ col, int rows,
                                        int col_start, int col_count, int *__restrict__ cnt) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    int c = 0;
    for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) {
        int j = col[k];
        c += (j >= col_start && j < col_start + col_count);
    }
    cnt[r] = c;
}
__global__ void block_view_fill_kernel(const int *__restrict__ row_ptr, const int *__restrict__ col, int rows,
                                       int col_start, int col_count, const int *__restrict__ out_ptr,
                                       int *__restrict__ out_col) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    int o = out_ptr[r];
    for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) {
        int j = col[k];
        if (j >= col_start && j < col_start + col_count) out_col[o++] = j - col_start;
    }
}

}  // namespace

int kmc_build_cellgrid(kmcb200_ctx *ctx, const double *x, const double *y, const double *z, int first, int count,
                       double cutoff, int pbc, const double *lattice_host, CellGridDev *g) {
    double bb[6];
    KMC_TRY(kmc_bbox(ctx, x, y, z, first, count, bb));
    double h = cutoff * 1.0001;
    g->x0 = bb[0]; g->y0 = bb[1]; g->z0 = bb[2];
    g->hx = g->hy = g->hz = h;
    g->nx = (int)floor((bb[3] - bb[0]) / h) + 1;
    g->wrap_yz = (pbc == 1);
    g->latty = lattice_host ? lattice_host[1] : 1.0;
    g->lattz = lattice_host ? lattice_host[2] : 1.0;
    if (g->wrap_yz) {
        g->ny = (int)floor(g->latty / h);
        g->nz = (int)floor(g->lattz / h);
        if (g->ny < 1) g->ny = 1;
        if (g->nz < 1) g->nz = 1;
    } else {
        g->ny = (int)floor((bb[4] - bb[1]) / h) + 1;
        g->nz = (int)floor((bb[5] - bb[2]) / h) + 1;
    }
    if (g->nx < 1) g->nx = 1;
    long long ncell = (long long)g->nx * g->ny * g->nz;
    if (ncell > 400000000LL) {
        kmc_set_error("cell grid too large (%lld cells)", ncell);
        return KMCB200_E_CAPACITY;
    }
    int *cell_start = nullptr, *items = nullptr, *cid = nullptr, *fill = nullptr;
    KMC_TRY(kmc_scratch(ctx, 0, (size_t)(ncell + 1) * sizeof(int), (void **)&cell_start));
    KMC_TRY(kmc_scratch(ctx, 1, (size_t)(count + 1) * sizeof(int), (void **)&items));
    KMC_TRY(kmc_scratch(ctx, 2, (size_t)(count + 1) * sizeof(int), (void **)&cid));
    KMC_TRY(kmc_scratch(ctx, 3, (size_t)(ncell + 1) * sizeof(int), (void **)&fill));
    KMC_CUDA(cudaMemsetAsync(cell_start, 0, (size_t)(ncell + 1) * sizeof(int), ctx->stream));
    KMC_CUDA(cudaMemsetAsync(fill, 0, (size_t)(ncell + 1) * sizeof(int), ctx->stream));
    g->cell_start = cell_start;
    g->items = items;
    if (count > 0) {
        int blocks = (count + 255) / 256;
        kmc_count_launch();
        cell_count_kernel<<<blocks, 256, 0, ctx->stream>>>(*g, x, y, z, first, count, cid, cell_start);
        KMC_CUDA(cudaGetLastError());
        KMC_TRY(kmc_exclusive_scan_i32(ctx, cell_start, cell_start, ncell + 1, 4));
        kmc_count_launch();
        cell_fill_kernel<<<blocks, 256, 0, ctx->stream>>>(cid, cell_start, fill, items, first, count);
        KMC_CUDA(cudaGetLastError());
    }
    return 0;
}

extern "C" int kmcb200_compute_neighbor_list(kmcb200_ctx *ctx, int N, const double *x, const double *y,
                                             const double *z, double nn_dist, int nn, int row_start, int row_count,
                                             int *neigh_out) {
    KMC_CHECK_ARG(ctx && x && y && z && (neigh_out || row_count == 0), "null pointer");
    KMC_CHECK_ARG(N > 0 && nn > 0 && nn <= MAX_NN, "N > 0, 0 < nn <= 64");
    KMC_CHECK_ARG(row_start >= 0 && row_count >= 0 && row_start + row_count <= N, "row range");
    if (row_count == 0) return 0;
    CellGridDev g;
    KMC_TRY(kmc_build_cellgrid(ctx, x, y, z, 0, N, nn_dist, 0, nullptr, &g));
    kmc_count_launch();
    neighbor_kernel<<<(N + 127) / 128, 128, 0, ctx->stream>>>(g, x, y, z, N, nn_dist, nn, row_start, row_count,
                                                            neigh_out);
    KMC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int kmcb200_cutoff_size(kmcb200_ctx *ctx, int N, const int *element, const double *x, const double *y,
                                   const double *z, double cutoff_radius, int row_start, int row_count,
                                   int *counts_out, int *max_count_host) {
    KMC_CHECK_ARG(ctx && element && x && y && z && max_count_host, "null pointer");
    KMC_CHECK_ARG(row_start >= 0 && row_count >= 0 && row_start + row_count <= N, "row range");
    CellGridDev g;
    KMC_TRY(kmc_build_cellgrid(ctx, x, y, z, 0, N, cutoff_radius, 0, nullptr, &g));
    int *d_max = nullptr;
    KMC_TRY(kmc_scratch(ctx, 5, 64, (void **)&d_max));
    KMC_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), ctx->stream));
    kmc_count_launch();
    cutoff_count_kernel<<<(N + 127) / 128, 128, 0, ctx->stream>>>(g, element, x, y, z, N, cutoff_radius, row_start,
                                                                row_count, counts_out, d_max);
    KMC_CUDA(cudaGetLastError());
    KMC_CUDA(cudaMemcpyAsync(max_count_host, d_max, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int kmcb200_cutoff_list(kmcb200_ctx *ctx, int N, const int *element, const double *x, const double *y,
                                   const double *z, double cutoff_radius, int max_num_cutoff, int row_start,
                                   int row_count, int *cutoff_idx_out) {
    KMC_CHECK_ARG(ctx && element && x && y && z && cutoff_idx_out, "null pointer");
    KMC_CHECK_ARG(row_start >= 0 && row_count >= 0 && row_start + row_count <= N && max_num_cutoff >= 0, "range");
    if (row_count == 0 || max_num_cutoff == 0) return 0;
    kmc_count_launch();
    cutoff_list_kernel<<<(row_count + 127) / 128, 128, 0, ctx->stream>>>(element, x, y, z, N, cutoff_radius,
                                                                       max_num_cutoff, row_start, row_count,
                                                                       cutoff_idx_out);
    KMC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int kmcb200_initialize_sparsity_K(kmcb200_ctx *ctx, int N, const double *x, const double *y,
                                             const double *z, const double *lattice_host, int pbc, double nn_dist,
                                             int N_left, int N_right, int row_start, int row_count,
                                             kmcb200_kmat **kmat_out) {
    KMC_CHECK_ARG(ctx && x && y && z && lattice_host && kmat_out, "null pointer");
    int n_int = N - N_left - N_right;
    KMC_CHECK_ARG(n_int > 0 && N_left >= 0 && N_right >= 0, "N_left/N_right");
    KMC_CHECK_ARG(row_start >= 0 && row_count > 0 && row_start + row_count <= n_int, "row range");
    CellGridDev g;
    KMC_TRY(kmc_build_cellgrid(ctx, x, y, z, 0, N, nn_dist, pbc, lattice_host, &g));

    kmcb200_kmat *K = new kmcb200_kmat();
    K->ctx = ctx;
    K->rows = row_count;
    K->row_start = row_start;
    K->cols_global = n_int;
    K->owns_csr = true;
    size_t rp_bytes = (size_t)(row_count + 1) * sizeof(int);
    int rc = 0;
    auto fail = [&](int code) { kmcb200_kmat_destroy(K); return code; };
    if (cudaMalloc(&K->row_ptr, rp_bytes) != cudaSuccess || cudaMalloc(&K->left_row_ptr, rp_bytes) != cudaSuccess ||
        cudaMalloc(&K->right_row_ptr, rp_bytes) != cudaSuccess) {
        kmc_set_error("cudaMalloc(row_ptr) failed");
        return fail(KMCB200_E_CUDA);
    }
    cudaMemsetAsync(K->row_ptr, 0, rp_bytes, ctx->stream);
    cudaMemsetAsync(K->left_row_ptr, 0, rp_bytes, ctx->stream);
    cudaMemsetAsync(K->right_row_ptr, 0, rp_bytes, ctx->stream);
    int blocks = (row_count + 127) / 128;
    kmc_count_launch();
    ksparsity_count_kernel<<<blocks, 128, 0, ctx->stream>>>(g, x, y, z, N, N_left, N_right, pbc, nn_dist, row_start,
                                                          row_count, K->row_ptr, K->left_row_ptr, K->right_row_ptr);
    if (cudaGetLastError() != cudaSuccess) { kmc_set_error("ksparsity_count launch failed"); return fail(KMCB200_E_CUDA); }
    if ((rc = kmc_exclusive_scan_i32(ctx, K->row_ptr, K->row_ptr, row_count + 1, 4))) return fail(rc);
    if ((rc = kmc_exclusive_scan_i32(ctx, K->left_row_ptr, K->left_row_ptr, row_count + 1, 4))) return fail(rc);
    if ((rc = kmc_exclusive_scan_i32(ctx, K->right_row_ptr, K->right_row_ptr, row_count + 1, 4))) return fail(rc);
    int tot[3];
    cudaMemcpyAsync(&tot[0], K->row_ptr + row_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&tot[1], K->left_row_ptr + row_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&tot[2], K->right_row_ptr + row_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { kmc_set_error("sync failed in sparsity K"); return fail(KMCB200_E_CUDA); }
    K->nnz = tot[0]; K->left_nnz = tot[1]; K->right_nnz = tot[2];
    if (cudaMalloc(&K->col, (size_t)(K->nnz + 1) * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&K->left_col, (size_t)(K->left_nnz + 1) * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&K->right_col, (size_t)(K->right_nnz + 1) * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&K->val, (size_t)(K->nnz + 1) * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&K->inv_diag, (size_t)row_count * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&K->rhs, (size_t)row_count * sizeof(double)) != cudaSuccess) {
        kmc_set_error("cudaMalloc(K arrays, nnz=%lld) failed", K->nnz);
        return fail(KMCB200_E_CUDA);
    }
    int *d_ovf = nullptr;
    if ((rc = kmc_scratch(ctx, 5, 64, (void **)&d_ovf))) return fail(rc);
    cudaMemsetAsync(d_ovf, 0, sizeof(int), ctx->stream);
    kmc_count_launch();
    ksparsity_fill_kernel<<<blocks, 128, 0, ctx->stream>>>(g, x, y, z, N, N_left, N_right, pbc, nn_dist, row_start,
                                                         row_count, K->row_ptr, K->left_row_ptr, K->right_row_ptr,
                                                         K->col, K->left_col, K->right_col, d_ovf);
    if (cudaGetLastError() != cudaSuccess) { kmc_set_error("ksparsity_fill launch failed"); return fail(KMCB200_E_CUDA); }
    int ovf = 0;
    cudaMemcpyAsync(&ovf, d_ovf, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { kmc_set_error("sparsity K fill failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(KMCB200_E_CUDA); }
    if (ovf) {
        kmc_set_error("a K row has more than %d entries within nn_dist", MAX_ROW_CAND);
        return fail(KMCB200_E_CAPACITY);
    }
    if ((rc = kmc_kmat_finalize(K))) return fail(rc);
    *kmat_out = K;
    return 0;
}

// Non-zeros per interior row (interior block only) for ALL rows: lets the caller pick nnz-balanced, chunk-aligned rank
// boundaries before any rank builds its shard (the reference splits by row count only, src/KMC_comm.h:249-263).
extern "C" int kmcb200_sparsity_K_row_counts(kmcb200_ctx *ctx, int N, const double *x, const double *y, const double *z,
                                             const double *lattice_host, int pbc, double nn_dist, int N_left,
                                             int N_right, int *row_nnz_dev) {
    KMC_CHECK_ARG(ctx && x && y && z && lattice_host && row_nnz_dev, "null pointer");
    int n_int = N - N_left - N_right;
    KMC_CHECK_ARG(n_int > 0, "N_left/N_right");
    CellGridDev g;
    KMC_TRY(kmc_build_cellgrid(ctx, x, y, z, 0, N, nn_dist, pbc, lattice_host, &g));
    int *tmp = nullptr;
    KMC_TRY(kmc_scratch(ctx, 6, (size_t)2 * n_int * sizeof(int), (void **)&tmp));
    kmc_count_launch();
    ksparsity_count_kernel<<<(n_int + 127) / 128, 128, 0, ctx->stream>>>(g, x, y, z, N, N_left, N_right, pbc, nn_dist, 0,
                                                                        n_int, row_nnz_dev, tmp, tmp + n_int);
    KMC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int kmcb200_kmat_block_view(kmcb200_kmat *K, int col_start, int col_count, int *row_ptr_out, int *col_out,
                                       long long *nnz_host) {
    KMC_CHECK_ARG(K && row_ptr_out && nnz_host, "null pointer");
    kmcb200_ctx *ctx = K->ctx;
    int blocks = (K->rows + 127) / 128;
    KMC_CUDA(cudaMemsetAsync(row_ptr_out, 0, (size_t)(K->rows + 1) * sizeof(int), ctx->stream));
    kmc_count_launch();
    block_view_count_kernel<<<blocks, 128, 0, ctx->stream>>>(K->row_ptr, K->col, K->rows, col_start, col_count, row_ptr_out);
    KMC_CUDA(cudaGetLastError());
    KMC_TRY(kmc_exclusive_scan_i32(ctx, row_ptr_out, row_ptr_out, K->rows + 1, 4));
    int tot = 0;
    KMC_CUDA(cudaMemcpyAsync(&tot, row_ptr_out + K->rows, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ctx->stream));
    *nnz_host = tot;
    if (col_out) {
        kmc_count_launch();
        block_view_fill_kernel<<<blocks, 128, 0, ctx->stream>>>(K->row_ptr, K->col, K->rows, col_start, col_count,
                                                              row_ptr_out, col_out);
        KMC_CUDA(cudaGetLastError());
    }
    return 0;
}

// kmat.cu -- K matrix object + a6 fused assembly kernel.
// Reference: src/potential_solver_gpu.cu:246-285 (calc_off_diagonal_dist), :323-367 (reduce_contact_into_diag),
// :438-454 (calc_rhs_for_A), :774-830 (reduce_rows_into_diag, insert_into_diag, inverse_diag), driver :893-1029.
// The reference runs 7 thread-per-row kernels + 5 memsets per step; here one kernel with 8 lanes per row writes
// values, diagonal, inverse diagonal and rhs in a single pass over the CSR (coalesced col reads / val writes).
#include <stdlib.h>

#include "comm.cuh"
#include "kmat.cuh"

namespace {

// bit0: metal, bit1: uncharged vacancy ("cvacancy" in the reference kernels)
__global__ void site_class_kernel(const int *__restrict__ element, const int *__restrict__ charge, int N,
                                  unsigned metal_mask, unsigned char *__restrict__ cls) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int el = element[i];
    unsigned c = ((metal_mask >> el) & 1u) | ((el == KMCB200_VACANCY && charge[i] == 0) ? 2u : 0u);
    cls[i] = (unsigned char)c;
}

template <int L>
__device__ __forceinline__ double seq_block_sum(const int *__restrict__ rp, const int *__restrict__ colv, int r,
                                                int col_offset, unsigned ci, const unsigned char *__restrict__ cls,
                                                double high_G, double low_G, int lane, int gbase, unsigned gmask) {
    // sequential (column order) sum of the conductances of one contact block row; all L lanes get the sum
    int s = rp[r], e = rp[r + 1];
    double sum = 0.0;
    for (int base = s; base < e; base += L) {
        int k = base + lane;
        double g = 0.0;
        if (k < e) {
            unsigned cj = cls[col_offset + colv[k]];
            g = (ci & cj) ? high_G : low_G;
        }
#pragma unroll
        for (int l = 0; l < L; ++l) {
            double gl = __shfl_sync(gmask, g, gbase + l);
            if (base + l < e) sum += gl;
        }
    }
    return sum;
}

template <int L>
__global__ void __launch_bounds__(256) assemble_kernel(int rows, int row_start, int N_left, int n_int,
                                                      const int *__restrict__ row_ptr, const int *__restrict__ col,
                                                      double *__restrict__ val, const int *__restrict__ lrp,
                                                      const int *__restrict__ lcol, const int *__restrict__ rrp,
                                                      const int *__restrict__ rcol,
                                                      const unsigned char *__restrict__ cls, double VL, double VR,
                                                      double high_G, double low_G, double *__restrict__ inv_diag,
                                                      double *__restrict__ rhs) {
    int gt = blockIdx.x * blockDim.x + threadIdx.x;
    int r = gt / L;
    int lane = threadIdx.x % L;
    int gbase = (threadIdx.x & 31) - lane;  // first lane of this row group inside the warp
    const unsigned gmask = (L == 32 ? 0xffffffffu : ((1u << L) - 1u)) << gbase;  // groups loop independently
    bool active = r < rows;
    int rr = active ? r : rows - 1;  // keep whole warps converged for the shuffles
    int gi = row_start + rr;         // global interior row
    unsigned ci = cls[N_left + gi];
    int s = row_ptr[rr], e = row_ptr[rr + 1];
    double sum = 0.0;
    int diag_k = -1;
    for (int base = s; base < e; base += L) {
        int k = base + lane;
        double g = 0.0;
        if (k < e) {
            int j = col[k];
            if (j != gi) {
                unsigned cj = cls[N_left + j];
                g = (ci & cj) ? high_G : low_G;
                if (active) val[k] = -g;
            } else {
                diag_k = k;
            }
        }
#pragma unroll
        for (int l = 0; l < L; ++l) {
            double gl = __shfl_sync(gmask, g, gbase + l);
            if (base + l < e) sum += gl;
        }
    }
    double left = seq_block_sum<L>(lrp, lcol, rr, 0, ci, cls, high_G, low_G, lane, gbase, gmask);
    double right = seq_block_sum<L>(rrp, rcol, rr, N_left + n_int, ci, cls, high_G, low_G, lane, gbase, gmask);
    double d = sum + left + right;
    if (!active) return;
    if (diag_k >= 0) val[diag_k] = d;
    if (lane == 0) {
        inv_diag[r] = 1.0 / d;
        rhs[r] = left * VL + right * VR;
    }
}

}  // namespace

int kmc_kmat_finalize(kmcb200_kmat *K) {
    KMC_CUDA(cudaMalloc(&K->Ap, (size_t)K->rows * sizeof(double)));
    KMC_CUDA(cudaMalloc(&K->z, (size_t)K->rows * sizeof(double)));
    if (K->rows == K->cols_global) {  // single rank: private exchange plan with empty peer loops
        KMC_TRY(kmc_comm_create_local(K->ctx, K->rows, &K->comm));
        K->owns_comm = true;
    }
    K->plan_max_unique = 0;
    static const bool staged = getenv("KMCB200_SPMV_STAGED") != nullptr;  // opt-in experiment, see pcg.cu
    if (staged) return kmc_build_spmv_plan(K);
    return 0;
}

extern "C" int kmcb200_kmat_from_csr(kmcb200_ctx *ctx, int rows, int cols_global, int row_start,
                                     const int *row_ptr_dev, const int *col_dev, const double *val_dev,
                                     kmcb200_kmat **kmat_out) {
    KMC_CHECK_ARG(ctx && row_ptr_dev && col_dev && val_dev && kmat_out, "null pointer");
    KMC_CHECK_ARG(rows > 0 && cols_global >= rows && row_start >= 0 && row_start + rows <= cols_global, "sizes");
    kmcb200_kmat *K = new kmcb200_kmat();
    K->ctx = ctx;
    K->rows = rows;
    K->row_start = row_start;
    K->cols_global = cols_global;
    K->owns_csr = false;
    K->row_ptr = const_cast<int *>(row_ptr_dev);
    K->col = const_cast<int *>(col_dev);
    K->val = const_cast<double *>(val_dev);
    int nnz = 0;
    KMC_CUDA(cudaMemcpyAsync(&nnz, row_ptr_dev + rows, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ctx->stream));
    K->nnz = nnz;
    int rc = kmc_kmat_finalize(K);
    if (rc) { kmcb200_kmat_destroy(K); return rc; }
    *kmat_out = K;
    return 0;
}

extern "C" int kmcb200_kmat_destroy(kmcb200_kmat *K) {
    if (!K) return 0;
    if (K->ctx) cudaStreamSynchronize(K->ctx->stream);
    if (K->owns_csr) {
        cudaFree(K->row_ptr); cudaFree(K->col); cudaFree(K->val);
        cudaFree(K->left_row_ptr); cudaFree(K->left_col); cudaFree(K->right_row_ptr); cudaFree(K->right_col);
        cudaFree(K->inv_diag); cudaFree(K->rhs);
    }
    if (K->comm && K->owns_comm) kmcb200_comm_destroy(K->comm);
    cudaFree(K->Ap); cudaFree(K->z); cudaFree(K->site_class);
    cudaFree(K->u_ptr); cudaFree(K->u_col); cudaFree(K->lcol);
    delete K;
    return 0;
}

extern "C" int kmcb200_kmat_info(kmcb200_kmat *K, int *rows, long long *nnz, long long *left_nnz, long long *right_nnz) {
    KMC_CHECK_ARG(K != nullptr, "kmat");
    if (rows) *rows = K->rows;
    if (nnz) *nnz = K->nnz;
    if (left_nnz) *left_nnz = K->left_nnz;
    if (right_nnz) *right_nnz = K->right_nnz;
    return 0;
}

extern "C" int kmcb200_kmat_pointers(kmcb200_kmat *K, int **row_ptr, int **col, double **val, int **left_row_ptr,
                                     int **left_col, int **right_row_ptr, int **right_col, double **inv_diag,
                                     double **rhs) {
    KMC_CHECK_ARG(K != nullptr, "kmat");
    if (row_ptr) *row_ptr = K->row_ptr;
    if (col) *col = K->col;
    if (val) *val = K->val;
    if (left_row_ptr) *left_row_ptr = K->left_row_ptr;
    if (left_col) *left_col = K->left_col;
    if (right_row_ptr) *right_row_ptr = K->right_row_ptr;
    if (right_col) *right_col = K->right_col;
    if (inv_diag) *inv_diag = K->inv_diag;
    if (rhs) *rhs = K->rhs;
    return 0;
}

extern "C" int kmcb200_assemble_K(kmcb200_ctx *ctx, kmcb200_kmat *K, int N, int N_left, int N_right,
                                  const int *element, const int *charge, const int *metals_host, int num_metals,
                                  double Vd, double high_G, double low_G) {
    KMC_CHECK_ARG(ctx && K && element && charge, "null pointer");
    KMC_CHECK_ARG(K->owns_csr && K->left_row_ptr && K->right_row_ptr, "kmat was not built by initialize_sparsity_K");
    KMC_CHECK_ARG(N - N_left - N_right == K->cols_global, "N/N_left/N_right do not match the K sparsity");
    KMC_CHECK_ARG(num_metals >= 0 && num_metals <= KMCB200_MAX_METALS && (num_metals == 0 || metals_host), "metals");
    unsigned metal_mask = 0;
    for (int m = 0; m < num_metals; ++m) {
        KMC_CHECK_ARG(metals_host[m] >= 0 && metals_host[m] < 32, "metal id");
        metal_mask |= 1u << metals_host[m];
    }
    if (K->site_class_cap < (size_t)N) {
        if (K->site_class) { KMC_CUDA(cudaStreamSynchronize(ctx->stream)); KMC_CUDA(cudaFree(K->site_class)); }
        KMC_CUDA(cudaMalloc(&K->site_class, (size_t)N));
        K->site_class_cap = (size_t)N;
    }
    kmc_count_launch();
    site_class_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(element, charge, N, metal_mask, K->site_class);
    KMC_CUDA(cudaGetLastError());
    constexpr int L = KMCB200_SPMV_LANES;
    long long threads = (long long)K->rows * L;
    unsigned blocks = (unsigned)((threads + 255) / 256);
    kmc_count_launch();
    assemble_kernel<L><<<blocks, 256, 0, ctx->stream>>>(K->rows, K->row_start, N_left, K->cols_global, K->row_ptr,
                                                       K->col, K->val, K->left_row_ptr, K->left_col,
                                                       K->right_row_ptr, K->right_col, K->site_class, -Vd / 2,
                                                       Vd / 2, high_G, low_G, K->inv_diag, K->rhs);
    KMC_CUDA(cudaGetLastError());
    return 0;
}

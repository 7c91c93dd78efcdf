// coulomb.cu -- a5 site charges, a8 screened-Coulomb potential of the charged defects, a9 potential sum.
// Reference: src/potential_solver_gpu.cu:12-85 (update_charge), :1525-1564 + :1620-1655
// (calculate_pairwise_interaction_indexed / poisson_gridless_gpu), :832-843 + :1130-1151 (sum).
//
// The reference streams a materialised N x N_cutoff int32 list (635 MB at 5 nm) per step with one thread per
// site to find the O(10^2) charged sources of that site.  Here the charged sites (Q << N) are compacted in
// ascending site order each step and an N x Q kernel tiles them through shared memory; the membership test of
// the cutoff list (element class, r < cutoff, i != j) is evaluated inline, and every site sums its sources in
// ascending j like the reference thread does.
#include "common.cuh"

namespace {

// potential_solver_gpu.cu:12-63.  One thread per site; only V / Od sites touch their neighbour row.
__global__ void __launch_bounds__(256) update_charge_kernel(const int *__restrict__ element, int *__restrict__ charge,
                                                           const int *__restrict__ neigh, int nn, unsigned metal_mask,
                                                           int row_start, int row_count) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= row_count) return;
    int i = idx + row_start;
    int el = element[i];
    if (el == KMCB200_VACANCY) {
        int c = 2, Vnn = 0;
        const int *row = neigh + (size_t)idx * nn;
        for (int n = 0; n < nn; ++n) {
            int j = row[n];
            if (j >= 0) {
                int ej = element[j];
                if (ej == KMCB200_VACANCY) Vnn++;
                if ((metal_mask >> ej) & 1u) c = 0;
                if (Vnn >= 2) c = 0;
            }
        }
        charge[i] = c;
    } else if (el == KMCB200_OXYGEN_DEFECT) {
        int c = -2;
        const int *row = neigh + (size_t)idx * nn;
        for (int n = 0; n < nn; ++n) {
            int j = row[n];
            if (j >= 0 && ((metal_mask >> element[j]) & 1u)) c = 0;
        }
        charge[i] = c;
    }
}

__global__ void charged_flag_kernel(const int *__restrict__ element, const int *__restrict__ charge, int N,
                                    int *__restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > N) return;
    flag[i] = (i < N && charge[i] != 0 && kmc_possibly_charged(element[i])) ? 1 : 0;
}

struct __align__(16) Source {
    double x, y, z;
    int charge, idx;
};

__global__ void charged_scatter_kernel(const int *__restrict__ charge, const int *__restrict__ element,
                                       const double *__restrict__ x, const double *__restrict__ y,
                                       const double *__restrict__ z, int N, const int *__restrict__ offs,
                                       Source *__restrict__ src) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (charge[i] != 0 && kmc_possibly_charged(element[i])) {
        Source s;
        s.x = x[i]; s.y = y[i]; s.z = z[i]; s.charge = charge[i]; s.idx = i;
        src[offs[i]] = s;
    }
}

constexpr int CT = 128;    // target sites per CTA
constexpr int TILE = 256;  // sources per shared-memory tile

// potential_solver_gpu.cu:1541-1562: V_i = sum_j v_solve(1e-10*|r_ij|, q_j), ascending j, overwrite.
__global__ void __launch_bounds__(CT) coulomb_kernel(const double *__restrict__ x, const double *__restrict__ y,
                                                    const double *__restrict__ z, const Source *__restrict__ src,
                                                    const int *__restrict__ nsrc_ptr, double sigma, double k,
                                                    double cutoff, int row_start, int row_count,
                                                    double *__restrict__ pot) {
    __shared__ Source tile[TILE];
    const int Q = *nsrc_ptr;
    int idx = blockIdx.x * CT + threadIdx.x;
    bool active = idx < row_count;
    int i = row_start + (active ? idx : 0);
    double xi = x[i], yi = y[i], zi = z[i];
    double local = 0.0;
    for (int base = 0; base < Q; base += TILE) {
        int nt = min(TILE, Q - base);
        __syncthreads();
        for (int t = threadIdx.x; t < nt; t += CT) tile[t] = src[base + t];
        __syncthreads();
        if (active) {
            for (int t = 0; t < nt; ++t) {
                Source s = tile[t];
                double d = kmc_dist_nopbc(xi, yi, zi, s.x, s.y, s.z);
                if (d < cutoff && s.idx != i) {
                    double dist = 1e-10 * d;
                    local += kmc_v_solve(dist, s.charge, sigma, k);
                }
            }
        }
    }
    if (active) pot[i] = local;
}

__global__ void sum_kernel(double *__restrict__ a, const double *__restrict__ b, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += b[i];
}

}  // namespace

extern "C" int kmcb200_update_charge(kmcb200_ctx *ctx, const int *element, int *charge, const int *neigh, int N, int nn,
                                     const int *metals_host, int num_metals, int row_start, int row_count) {
    KMC_CHECK_ARG(ctx && element && charge && neigh, "null pointer");
    KMC_CHECK_ARG(row_start >= 0 && row_count >= 0 && row_start + row_count <= N, "row range");
    KMC_CHECK_ARG(num_metals >= 0 && num_metals <= KMCB200_MAX_METALS && (num_metals == 0 || metals_host), "metals");
    unsigned metal_mask = 0;
    for (int m = 0; m < num_metals; ++m) metal_mask |= 1u << metals_host[m];
    if (row_count == 0) return 0;
    kmc_count_launch();
    update_charge_kernel<<<(row_count + 255) / 256, 256, 0, ctx->stream>>>(element, charge, neigh, nn, metal_mask,
                                                                          row_start, row_count);
    KMC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int kmcb200_poisson_gridless(kmcb200_ctx *ctx, int N, const double *x, const double *y, const double *z,
                                        const int *element, const int *charge, double sigma, double k,
                                        double cutoff_radius, int row_start, int row_count,
                                        double *site_potential_charge) {
    KMC_CHECK_ARG(ctx && x && y && z && element && charge && site_potential_charge, "null pointer");
    KMC_CHECK_ARG(row_start >= 0 && row_count >= 0 && row_start + row_count <= N, "row range");
    if (row_count == 0) return 0;
    int *offs = nullptr;
    Source *src = nullptr;
    KMC_TRY(kmc_scratch(ctx, 6, (size_t)(N + 1) * sizeof(int), (void **)&offs));
    kmc_count_launch();
    charged_flag_kernel<<<(N + 1 + 255) / 256, 256, 0, ctx->stream>>>(element, charge, N, offs);
    KMC_CUDA(cudaGetLastError());
    KMC_TRY(kmc_exclusive_scan_i32(ctx, offs, offs, (long long)N + 1, 4));  // offs[N] = Q
    // capacity: worst case every site charged; the scratch buffer grows lazily to what was needed so far
    int Q = 0;
    KMC_CUDA(cudaMemcpyAsync(ctx->h_mail, offs + N, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ctx->stream));
    Q = *(int *)ctx->h_mail;
    KMC_TRY(kmc_scratch(ctx, 7, (size_t)(Q + 1) * sizeof(Source), (void **)&src));
    kmc_count_launch();
    charged_scatter_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(charge, element, x, y, z, N, offs, src);
    KMC_CUDA(cudaGetLastError());
    kmc_count_launch();
    coulomb_kernel<<<(row_count + CT - 1) / CT, CT, 0, ctx->stream>>>(x, y, z, src, offs + N, sigma, k, cutoff_radius,
                                                                    row_start, row_count, site_potential_charge);
    KMC_CUDA(cudaGetLastError());
    ctx->last_num_charged = Q;
    ctx->last_pair_tests = (long long)Q * row_count;
    return 0;
}

extern "C" int kmcb200_poisson_stats(kmcb200_ctx *ctx, long long *num_charged, long long *pair_tests) {
    KMC_CHECK_ARG(ctx != nullptr, "ctx");
    if (num_charged) *num_charged = ctx->last_num_charged;
    if (pair_tests) *pair_tests = ctx->last_pair_tests;
    return 0;
}

extern "C" int kmcb200_sum_potential(kmcb200_ctx *ctx, int N, double *site_potential_charge,
                                     const double *site_potential_boundary) {
    KMC_CHECK_ARG(ctx && site_potential_charge && site_potential_boundary && N >= 0, "arguments");
    if (N == 0) return 0;
    kmc_count_launch();
    sum_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(site_potential_charge, site_potential_boundary, N);
    KMC_CUDA(cudaGetLastError());
    return 0;
}

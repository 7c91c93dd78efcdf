// spmv_plan.cu -- setup-time analysis of the K sparsity for the shared-memory staged SpMV (pcg.cu).
// For every chunk of 256 rows: sort the chunk's column ids in shared memory (bitonic), compact the unique ones
// (u_col) and binary-search every non-zero's column to get its 16-bit position in that list (lcol).
// Nothing here has a counterpart in the reference (rocsparse_spmv analyses the matrix internally,
// reference dist_iterative/dist_matrix.cpp:546-600 creates descriptors + the rocsparse "preprocess" stage).
#include <limits.h>

#include "kmat.cuh"

namespace {

constexpr int PLAN_CAP = 16384;  // max non-zeros of one 256-row chunk handled by the plan (64 per row)
constexpr int PT = 256;

__device__ __forceinline__ void bitonic_sort_smem(int *keys, int n /* power of two */) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += PT) {
                int ixj = i ^ j;
                if (ixj > i) {
                    int a = keys[i], b = keys[ixj];
                    bool up = ((i & k) == 0);
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// FILL = false: count unique columns per chunk (or flag overflow); FILL = true: write u_col and lcol
template <bool FILL>
__global__ void __launch_bounds__(PT) spmv_plan_kernel(int rows, const int *__restrict__ row_ptr,
                                                      const int *__restrict__ col, int *__restrict__ u_cnt_or_ptr,
                                                      int *__restrict__ u_col, unsigned short *__restrict__ lcol,
                                                      int *__restrict__ overflow) {
    extern __shared__ int keys[];  // PLAN_CAP ints
    __shared__ int warp_cnt[PT / 32];
    __shared__ int total_unique;
    const int r0 = blockIdx.x * 256;
    const int r1 = min(r0 + 256, rows);
    const int s0 = row_ptr[r0], s1 = row_ptr[r1];
    const int m = s1 - s0;
    if (m > PLAN_CAP) {
        if (threadIdx.x == 0) atomicExch(overflow, 1);
        if (!FILL && threadIdx.x == 0) u_cnt_or_ptr[blockIdx.x] = 0;
        return;
    }
    int n2 = 1;
    while (n2 < m) n2 <<= 1;
    for (int i = threadIdx.x; i < n2; i += PT) keys[i] = (i < m) ? col[s0 + i] : INT_MAX;
    __syncthreads();
    bitonic_sort_smem(keys, n2);
    // unique compaction: thread t owns the contiguous slice [t*per, (t+1)*per)
    const int per = (n2 + PT - 1) / PT;
    const int b = min(threadIdx.x * per, m), e = min(b + per, m);
    int cnt = 0;
    for (int i = b; i < e; ++i) cnt += (i == 0 || keys[i] != keys[i - 1]);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = cnt;
#pragma unroll
    for (int d = 1; d <= 16; d <<= 1) {
        int o = __shfl_up_sync(KMC_FULL_MASK, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) warp_cnt[w] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int q = 0; q < PT / 32; ++q) { int t = warp_cnt[q]; warp_cnt[q] = acc; acc += t; }
        total_unique = acc;
    }
    __syncthreads();
    int pos = warp_cnt[w] + incl - cnt;
    if (!FILL) {
        if (threadIdx.x == 0) u_cnt_or_ptr[blockIdx.x] = total_unique;
        return;
    }
    const int ubase = u_cnt_or_ptr[blockIdx.x];
    for (int i = b; i < e; ++i)
        if (i == 0 || keys[i] != keys[i - 1]) u_col[ubase + pos++] = keys[i];
    __syncthreads();  // (global writes of this CTA are visible to it after the barrier)
    const int nu = total_unique;
    for (int i = threadIdx.x; i < nu; i += PT) keys[i] = u_col[ubase + i];
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += PT) {
        int c = col[s0 + i];
        int lo = 0, hi = nu - 1;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (keys[mid] < c) lo = mid + 1; else hi = mid;
        }
        lcol[s0 + i] = (unsigned short)lo;
    }
}

}  // namespace

int kmc_build_spmv_plan(kmcb200_kmat *K) {
    kmcb200_ctx *ctx = K->ctx;
    const int nchunks = (K->rows + 255) / 256;
    K->plan_max_unique = 0;
    if (K->nnz <= 0) return 0;
    int *d_flag = nullptr;
    KMC_TRY(kmc_scratch(ctx, 5, 64, (void **)&d_flag));
    KMC_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream));
    KMC_CUDA(cudaMalloc(&K->u_ptr, (size_t)(nchunks + 1) * sizeof(int)));
    KMC_CUDA(cudaMemsetAsync(K->u_ptr, 0, (size_t)(nchunks + 1) * sizeof(int), ctx->stream));
    if (!ctx->smem_cfg_plan) {  // function attributes are per device: tracked per context, not per process
        KMC_CUDA(cudaFuncSetAttribute(spmv_plan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PLAN_CAP * 4));
        KMC_CUDA(cudaFuncSetAttribute(spmv_plan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PLAN_CAP * 4));
        ctx->smem_cfg_plan = true;
    }
    kmc_count_launch();
    spmv_plan_kernel<false><<<nchunks, PT, PLAN_CAP * 4, ctx->stream>>>(K->rows, K->row_ptr, K->col, K->u_ptr, nullptr,
                                                                       nullptr, d_flag);
    KMC_CUDA(cudaGetLastError());
    std::vector<int> h_cnt((size_t)nchunks);
    int ovf = 0;
    KMC_CUDA(cudaMemcpyAsync(h_cnt.data(), K->u_ptr, (size_t)nchunks * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    KMC_CUDA(cudaMemcpyAsync(&ovf, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ovf) {  // a chunk has more than PLAN_CAP non-zeros: keep the direct-gather kernel
        cudaFree(K->u_ptr);
        K->u_ptr = nullptr;
        return 0;
    }
    int mx = 0;
    long long total = 0;
    for (int v : h_cnt) { mx = v > mx ? v : mx; total += v; }
    KMC_TRY(kmc_exclusive_scan_i32(ctx, K->u_ptr, K->u_ptr, (long long)nchunks + 1, 4));
    KMC_CUDA(cudaMalloc(&K->u_col, (size_t)(total + 1) * sizeof(int)));
    KMC_CUDA(cudaMalloc(&K->lcol, (size_t)(K->nnz + 1) * sizeof(unsigned short)));
    kmc_count_launch();
    spmv_plan_kernel<true><<<nchunks, PT, PLAN_CAP * 4, ctx->stream>>>(K->rows, K->row_ptr, K->col, K->u_ptr, K->u_col,
                                                                      K->lcol, d_flag);
    KMC_CUDA(cudaGetLastError());
    K->plan_max_unique = mx;
    return 0;
}

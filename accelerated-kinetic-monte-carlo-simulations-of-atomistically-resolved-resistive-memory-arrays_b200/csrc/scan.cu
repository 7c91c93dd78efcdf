// scan.cu -- hand-written device-wide primitives (no CUB/Thrust): exclusive int32 scan
// (replaces hipcub::DeviceScan::InclusiveSum, reference src/iterative_solvers_gpu.cu:374-383 and
// thrust::reduce, src/neighbor_lists_gpu.cu:340) and bounding box reduction.
#include <float.h>
#include <string.h>

#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;                        // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048 per CTA

__device__ __forceinline__ int warp_incl_scan_i32(int v) {
    int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d <= 16; d <<= 1) {
        int o = __shfl_up_sync(KMC_FULL_MASK, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// Tile-local exclusive scan; writes per-tile totals.  in may equal out.
__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles(const int *__restrict__ in, int *__restrict__ out,
                                                          int *__restrict__ tile_sums, long long n) {
    __shared__ int warp_tot[SCAN_THREADS / 32];
    long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int tsum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        long long i = base + k;
        v[k] = (i < n) ? in[i] : 0;
        tsum += v[k];
    }
    int incl = warp_incl_scan_i32(tsum);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    if (w == 0) {
        int t = (lane < SCAN_THREADS / 32) ? warp_tot[lane] : 0;
        int ti = warp_incl_scan_i32(t);
        if (lane < SCAN_THREADS / 32) warp_tot[lane] = ti - t;  // exclusive warp offsets
        if (lane == SCAN_THREADS / 32 - 1 && tile_sums) tile_sums[blockIdx.x] = ti;
    }
    __syncthreads();
    int run = warp_tot[w] + incl - tsum;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        long long i = base + k;
        if (i < n) out[i] = run;
        run += v[k];
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) add_tile_offsets(int *__restrict__ out, const int *__restrict__ tile_off,
                                                                long long n) {
    long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    int off = tile_off[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        long long i = base + k;
        if (i < n) out[i] += off;
    }
}

int scan_rec(kmcb200_ctx *ctx, const int *in, int *out, long long n, int *work, int level) {
    long long tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles <= 0) return 0;
    if (tiles == 1) {
        kmc_count_launch();
        scan_tiles<<<1, SCAN_THREADS, 0, ctx->stream>>>(in, out, nullptr, n);
        KMC_CUDA(cudaGetLastError());
        return 0;
    }
    int *tile_sums = work;
    kmc_count_launch();
    scan_tiles<<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(in, out, tile_sums, n);
    KMC_CUDA(cudaGetLastError());
    KMC_TRY(scan_rec(ctx, tile_sums, tile_sums, tiles, work + tiles, level + 1));
    kmc_count_launch();
    add_tile_offsets<<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(out, tile_sums, n);
    KMC_CUDA(cudaGetLastError());
    return 0;
}

// order-preserving double <-> uint64 for atomicMin/Max
__device__ __forceinline__ unsigned long long d2ord(double d) {
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
double ord2d(unsigned long long u) {
    unsigned long long b = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
    double d;
    memcpy(&d, &b, 8);
    return d;
}

__global__ void bbox_kernel(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
                            int first, int count, unsigned long long *out6) {
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        double p[3] = {x[first + i], y[first + i], z[first + i]};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            mn[d] = fmin(mn[d], p[d]);
            mx[d] = fmax(mx[d], p[d]);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            mn[d] = fmin(mn[d], __shfl_xor_sync(KMC_FULL_MASK, mn[d], off));
            mx[d] = fmax(mx[d], __shfl_xor_sync(KMC_FULL_MASK, mx[d], off));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(out6 + d, d2ord(mn[d]));
            atomicMax(out6 + 3 + d, d2ord(mx[d]));
        }
    }
}

}  // namespace

// exclusive scan of n int32; out may alias in.  Uses scratch slot `scratch_slot` for tile sums.
int kmc_exclusive_scan_i32(kmcb200_ctx *ctx, const int *in, int *out, long long n, int scratch_slot) {
    if (n <= 0) return 0;
    long long tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    size_t work_elems = (size_t)tiles + (size_t)(tiles / SCAN_TILE + 2) + 4096;
    int *work = nullptr;
    KMC_TRY(kmc_scratch(ctx, scratch_slot, work_elems * sizeof(int), (void **)&work));
    return scan_rec(ctx, in, out, n, work, 0);
}

// bbox_host6 = {xmin,ymin,zmin,xmax,ymax,zmax}; host sync.
int kmc_bbox(kmcb200_ctx *ctx, const double *x, const double *y, const double *z, int first, int count,
             double *bbox_host6) {
    unsigned long long *d6 = nullptr;
    KMC_TRY(kmc_scratch(ctx, 11, 64, (void **)&d6));
    unsigned long long init[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull};
    KMC_CUDA(cudaMemcpyAsync(d6, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    if (count > 0) {
        int blocks = (count + 255) / 256;
        if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
        kmc_count_launch();
        bbox_kernel<<<blocks, 256, 0, ctx->stream>>>(x, y, z, first, count, d6);
        KMC_CUDA(cudaGetLastError());
    }
    unsigned long long h6[6];
    KMC_CUDA(cudaMemcpyAsync(h6, d6, sizeof(h6), cudaMemcpyDeviceToHost, ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (count <= 0) {
        for (int i = 0; i < 6; ++i) bbox_host6[i] = 0.0;
        return 0;
    }
    for (int i = 0; i < 6; ++i) bbox_host6[i] = ord2d(h6[i]);
    return 0;
}

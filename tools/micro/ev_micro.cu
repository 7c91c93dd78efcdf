// Micro-benchmark of the selector building blocks (one warp, data in shared memory).  Not part of the product.
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
struct Scan256 { double a[8]; double excl, total; };
__device__ __forceinline__ void scan_seq(Scan256 &r) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 1; k < 8; ++k) r.a[k] = r.a[k - 1] + r.a[k];
    double S = r.a[7];
#pragma unroll
    for (int d = 1; d <= 16; d <<= 1) { double o = __shfl_up_sync(FULL, S, d); if (lane >= d) S = o + S; }
    r.excl = __shfl_up_sync(FULL, S, 1);
    r.total = __shfl_sync(FULL, S, 31);
}
__device__ __forceinline__ void scan_tree(Scan256 &r) {
    const int lane = threadIdx.x & 31;
    double a0 = r.a[0], a1 = r.a[1], a2 = r.a[2], a3 = r.a[3], a4 = r.a[4], a5 = r.a[5], a6 = r.a[6], a7 = r.a[7];
    double s01 = a0 + a1, s23 = a2 + a3, s45 = a4 + a5, s67 = a6 + a7;
    double s03 = s01 + s23, s47 = s45 + s67;
    double s07 = s03 + s47;
    r.a[1] = s01; r.a[2] = s01 + a2; r.a[3] = s03; r.a[4] = s03 + a4; r.a[5] = s03 + s45; r.a[6] = (s03 + s45) + a6; r.a[7] = s07;
    double S = s07;
#pragma unroll
    for (int d = 1; d <= 16; d <<= 1) { double o = __shfl_up_sync(FULL, S, d); if (lane >= d) S = o + S; }
    r.excl = __shfl_up_sync(FULL, S, 1);
    r.total = __shfl_sync(FULL, S, 31);
}
__device__ __forceinline__ double scan_incl(const Scan256 &r, int k) { return ((threadIdx.x & 31) > 0) ? (r.excl + r.a[k]) : r.a[k]; }
__device__ __forceinline__ int pick(const Scan256 &sc, const double v[8], double number, double *prev) {
    double inc[8]; int kfirst = 8, klast = -1;
#pragma unroll
    for (int k = 7; k >= 0; --k) { inc[k] = scan_incl(sc, k); if (inc[k] > number) kfirst = k; }
#pragma unroll
    for (int k = 0; k < 8; ++k) if (v[k] > 0.0) klast = k;
    int tsel = -1;
    unsigned bf = __ballot_sync(FULL, kfirst < 8);
    if (bf) { int l = __ffs(bf) - 1; tsel = l * 8 + __shfl_sync(FULL, kfirst, l); }
    else { unsigned bl = __ballot_sync(FULL, klast >= 0); if (bl) { int l = 31 - __clz(bl); tsel = l * 8 + __shfl_sync(FULL, klast, l); } }
    double pv = 0.0;
    if (tsel > 0) { int pl = (tsel - 1) >> 3, pk = (tsel - 1) & 7; double cand = inc[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) if (k == pk) cand = inc[k];
        pv = __shfl_sync(FULL, cand, pl); }
    *prev = pv; return tsel;
}
template <int V> __global__ void bench(const double *in, int iters, long long *cycles, double *out) {
    __shared__ double sm[4096];
    for (int q = threadIdx.x; q < 4096; q += blockDim.x) sm[q] = in[q];
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    double number = 0.3; int base = 0; double accum = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (V == 0 || V == 1) {
            Scan256 sc; double v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { v[k] = sm[base + 8 * lane + k]; sc.a[k] = v[k]; }
            if (V == 0) scan_seq(sc); else scan_tree(sc);
            double prev; int t = pick(sc, v, number * sc.total, &prev);
            base = (t * 16) & 3840; number = 0.1 + 0.5 * (prev / sc.total); accum += prev;
        } else if (V == 2) {  // pick from cached inclusive values: lane compares its 8 values
            double inc[8]; int kfirst = 8;
#pragma unroll
            for (int k = 7; k >= 0; --k) { inc[k] = sm[base + 8 * lane + k]; if (inc[k] > number * 100.0) kfirst = k; }
            unsigned bf = __ballot_sync(FULL, kfirst < 8);
            int t = 255;
            if (bf) { int l = __ffs(bf) - 1; t = l * 8 + __shfl_sync(FULL, kfirst, l); }
            double prev = sm[base + ((t + 255) & 255)];
            base = (t * 16) & 3840; number = 0.1 + 0.5 * (prev / 300.0); accum += prev;
        } else if (V == 3) {  // 16 dependent DADD
#pragma unroll
            for (int k = 0; k < 16; ++k) number = number + 1.0000001;
        } else if (V == 4) {  // 8 dependent (SHFL64 + DADD)
#pragma unroll
            for (int k = 0; k < 8; ++k) { double o = __shfl_up_sync(FULL, number, 1); if (lane >= 1) number = o + number; }
        } else if (V == 5) {  // dependent smem load chain
            base = (int)sm[(base + lane) & 4095] & 4095; 
        } else if (V == 6) {  // 8 dependent xor butterfly steps
#pragma unroll
            for (int k = 0; k < 8; ++k) number = number + __shfl_xor_sync(FULL, number, 1 << (k % 5));
        }
    }
    long long t1 = clock64();
    if (lane == 0) { cycles[0] = t1 - t0; out[0] = number + base + accum; }
}
int main() {
    double *in, *out; long long *cyc;
    cudaMalloc(&in, 4096 * 8); cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    double h[4096]; for (int i = 0; i < 4096; ++i) h[i] = (i * 37 % 11 == 0) ? 0.0 : 1.0 + (i % 7) * 0.01;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    const int iters = 2000; long long c;
#define RUN(V, name, div) bench<V><<<1, 512>>>(in, iters, cyc, out); cudaDeviceSynchronize(); bench<V><<<1, 512>>>(in, iters, cyc, out); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("%-40s %8.1f cycles/iter  (%.1f per unit)\n", name, (double)c / iters, (double)c / iters / div);
    RUN(0, "scan_seq + pick (smem)", 1)
    RUN(1, "scan_tree + pick (smem)", 1)
    RUN(2, "pick from cached inclusive (smem)", 1)
    RUN(3, "16 dependent DADD", 16)
    RUN(4, "8 x (SHFL64 up + DADD)", 8)
    RUN(5, "dependent smem load + cvt", 1)
    RUN(6, "8 x (SHFL64 xor + DADD)", 8)
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

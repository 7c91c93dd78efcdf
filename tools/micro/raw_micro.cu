// Micro-benchmark: latency of a 2 KB warp load (4 x LDG.128 per lane) that follows an 8-byte store into the same
// region by another warp of the same CTA (the R1 -> R2 pattern of the event loop).  Not part of the product.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(double *buf, size_t stride_elems, int iters, long long *out, double *sink) {
    __shared__ int dummy;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long tot = 0; double acc = 0;
    for (int it = 0; it < iters; ++it) {
        double *chunk = buf + (size_t)((it * 7919) % 4096) * stride_elems;  // a different 2 KB chunk each iteration
        if (MODE == 1 && warp == 3 && lane == 0) chunk[77] = (double)it;                 // 8-byte store, other warp
        if (MODE == 2 && warp == 3 && lane == 0) chunk[77 + 256] = (double)it;           // store to the NEXT chunk (other lines)
        if (MODE == 3 && warp == 3) chunk[8 * lane] = (double)it;                        // 32 stores, one per 64 B
        __syncthreads();
        if (warp == 0) {
            long long t0 = clock64();
            const double2 *s2 = reinterpret_cast<const double2 *>(chunk + 8 * lane);
            double2 a = s2[0], b = s2[1], c = s2[2], d = s2[3];
            double v = a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
            acc += v;
            long long t1 = clock64();
            tot += t1 - t0;
        }
        __syncthreads();
    }
    if (tid == 0) { out[0] = tot; sink[0] = acc + dummy * 0; }
}
int main() {
    const size_t stride = 256 * 16;  // chunks 32 KB apart
    double *buf, *sink; long long *out;
    cudaMalloc(&buf, 4096 * stride * 8 + 4096); cudaMemset(buf, 0, 4096 * stride * 8 + 4096);
    cudaMalloc(&sink, 8); cudaMalloc(&out, 8);
    const int iters = 3000; long long c;
#define RUN(M, name) k<M><<<1, 512>>>(buf, stride, iters, out, sink); cudaDeviceSynchronize(); k<M><<<1, 512>>>(buf, stride, iters, out, sink); cudaDeviceSynchronize(); cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost); printf("%-55s %8.1f cycles\n", name, (double)c / iters);
    RUN(0, "2 KB warp load, no store before")
    RUN(1, "2 KB warp load after an 8-byte store into it")
    RUN(2, "2 KB warp load after an 8-byte store elsewhere")
    RUN(3, "2 KB warp load after 32 8-byte stores into it")
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}

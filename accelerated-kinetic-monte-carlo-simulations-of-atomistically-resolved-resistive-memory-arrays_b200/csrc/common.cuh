// common.cuh -- shared declarations for libkmc_b200.so (B200 / sm_100a only).
// All translation units are compiled with --fmad=false: every fused multiply-add in the summation spec
//  is an explicit fma(); everything else rounds separately, exactly like the CPU
// oracle (g++ -ffp-contract=off).  That is what makes CG iterates and event choices bit-comparable.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/kmc_b200.h"

#define KMC_FULL_MASK 0xffffffffu

void kmc_set_error(const char *fmt, ...);
void kmc_count_launch();  // every kernel launch of this library is counted (kmcb200_launch_count)

#define KMC_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            kmc_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__,  \
                          cudaGetErrorString(e_));                                                  \
            return KMCB200_E_CUDA;                                                                  \
        }                                                                                           \
    } while (0)

#define KMC_CHECK_ARG(cond, msg)                                    \
    do {                                                            \
        if (!(cond)) {                                              \
            kmc_set_error("invalid argument: %s (%s)", msg, #cond); \
            return KMCB200_E_ARG;                                   \
        }                                                           \
    } while (0)

#define KMC_TRY(call)            \
    do {                         \
        int rc_ = (call);        \
        if (rc_ != 0) return rc_; \
    } while (0)

// Persistent per-context scratch: grows on demand, never shrinks; per-step calls are allocation free
// once warmed up (the reference hipMalloc/hipFree's its work vectors in every call,
// src/potential_solver_gpu.cu:857-881,1119-1126).
struct ScratchBuf {
    void *ptr = nullptr;
    size_t bytes = 0;
};

struct CgState;  // pcg.cu

struct kmcb200_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    ScratchBuf scratch[12];
    // pinned host mailbox for small D2H results
    void *h_mail = nullptr;
    // coulomb statistics of the last call
    long long last_num_charged = 0, last_pair_tests = 0;
    void *pair_counter_dev = nullptr;
    void *coulomb_plan = nullptr;  // CoulombPlan (coulomb.cu), owned by the context
    // dot/pcg workspace
    CgState *cg_state = nullptr;  // device
    bool cg_done_stale = false;   // a PCG solve ended abnormally: CgState::done may still be set
    // opted-in dynamic shared memory per kernel family (function attributes are per device, a ctx is bound to one)
    size_t smem_cfg_events = 0, smem_cfg_coulomb = 48 * 1024, smem_cfg_staged = 0;
    bool smem_cfg_plan = false;
    double *partials = nullptr;   // device, 2 * max chunks
    size_t partials_cap = 0;
    // persistent PCG loop kernel (pcg.cu): grid barrier counter + per-group arrival counters; co-resident CTAs per SM
    void *pcg_loop_ws = nullptr;
    size_t pcg_loop_ws_bytes = 0;
    int pcg_loop_occ = 0;
};

int kmc_scratch(kmcb200_ctx *ctx, int slot, size_t bytes, void **out);
void kmc_coulomb_plan_free(kmcb200_ctx *ctx);  // coulomb.cu

// ---- device helpers -------------------------------------------------------------------------------
// gpu_solvers.h:280-285 (reference): sqrt(pow(dx,2)+pow(dy,2)+pow(dz,2)); products and sums rounded
// separately (--fmad=false).
__device__ __forceinline__ double kmc_dist_nopbc(double x1, double y1, double z1, double x2, double y2, double z2) {
    double dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
    return sqrt(dx * dx + dy * dy + dz * dz);
}
// gpu_solvers.h:287-319 (reference): y/z minimum image via round()
__device__ __forceinline__ double kmc_dist_pbc(double x1, double y1, double z1, double x2, double y2, double z2,
                                               double latty, double lattz, int pbc) {
    if (pbc == 1) {
        double dist_x = x1 - x2;
        double fy = (y1 - y2) / latty;
        fy -= round(fy);
        double fz = (z1 - z2) / lattz;
        fz -= round(fz);
        double dy = fy * latty, dz = fz * lattz;
        return sqrt(dist_x * dist_x + dy * dy + dz * dz);
    }
    return kmc_dist_nopbc(x1, y1, z1, x2, y2, z2);
}
// gpu_solvers.h:321-328 (reference)
__device__ __forceinline__ double kmc_v_solve(double r_dist, int charge, double sigma, double k) {
    const double q = 1.60217663e-19;
    return (double)charge * erfc(r_dist / (sigma * sqrt(2.0))) * k * q / r_dist;
}
// Deterministic exp / x^1.5 for the WKB tunnel coefficients (initialize_sparsity_T.cu:497-614).  The reference calls the
// device math library there; libm implementations differ in the last ulp, and the macroscopic current is a strongly
// cancelling sum of the resulting potentials, so a 1-ulp difference in a coefficient shows up at 1e-9 relative in I_macro.
// These two routines use only +, *, fma, sqrt, ldexp (all correctly rounded on both sides), are restated operation by
// operation in the CPU oracle, and agree with libm's exp / pow(x, 1.5) to <= 2 ulp.
__device__ __forceinline__ double kmc_det_exp(double x) {
    if (!(x > -745.2)) return 0.0;
    if (x > 709.7) return 1.0 / 0.0;
    const double k = nearbyint(x * 1.4426950408889634);           // round(x / ln 2)
    double r = fma(-k, 0.693147180369123816490e+00, x);            // ln2_hi
    r = fma(-k, 1.90821492927058770002e-10, r);                    // ln2_lo
    double p = 1.0 / 6227020800.0;                                 // Taylor, degree 13, Horner with fma (|r| <= 0.347)
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return ldexp(p, (int)k);
}
__device__ __forceinline__ double kmc_det_pow15(double x) { return x * sqrt(x); }
__device__ __forceinline__ bool kmc_possibly_charged(int el) {
    return el == KMCB200_OXYGEN_DEFECT || el == KMCB200_O || el == KMCB200_VACANCY || el == KMCB200_DEFECT;
}

// Programmatic dependent launch (PDL): a kernel launched with kmc_launch_pdl may become resident while its predecessor
// on the stream is still draining; it must call kmc_pdl_wait() before it touches anything the predecessor reads or
// writes (the wait returns once the predecessor grid has completed and its stores are visible; it is a no-op for an
// ordinary launch).  kmc_pdl_trigger() lets the NEXT kernel on the stream start its own launch as soon as every CTA of
// this grid has started.  Only the launch latency and the CTA ramp-up overlap; all data accesses stay ordered.
__device__ __forceinline__ void kmc_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void kmc_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t kmc_launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem,
                                         cudaStream_t stream, bool pdl, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// summation spec: 32-lane butterfly, every lane ends with the same value
__device__ __forceinline__ double kmc_warp_xor_sum(double v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_xor_sync(KMC_FULL_MASK, v, off);
    return v;
}
// summation spec chunk_reduce_256: blockDim.x == 256, thread t holds v; result valid in thread 0.
// sm: at least 8 doubles of shared memory.
__device__ __forceinline__ double kmc_chunk_reduce_256(double v, double *sm) {
    v = kmc_warp_xor_sum(v);
    int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) sm[w] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
        s = sm[0];
#pragma unroll
        for (int q = 1; q < 8; ++q) s = s + sm[q];
    }
    __syncthreads();
    return s;
}
// summation spec final_reduce over n partials (blockDim.x == 256); result valid in thread 0
__device__ __forceinline__ double kmc_final_reduce(const double *partials, long long n, double *sm) {
    double acc = 0.0;
    long long k = threadIdx.x;
    // 8 loads in flight per trip; the additions keep the sequential order k, k+256, k+512, ...
    for (; k + 7 * 256 < n; k += 8 * 256) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(partials + k + u * 256);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = acc + v[u];
    }
    for (; k < n; k += 256) acc = acc + __ldcg(partials + k);
    return kmc_chunk_reduce_256(acc, sm);
}
// summation spec block_scan_256 (Kogge-Stone in warps, sequential over the 8 warp totals).
// blockDim.x may be a multiple of 256: `t` is the index inside the 256-group and sm8 that group's 8 doubles.
// All threads of the group must call; uses bar_sync callback = __syncthreads by caller convention.
__device__ __forceinline__ double kmc_warp_inclusive_scan(double v) {
    int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d <= 16; d <<= 1) {
        double o = __shfl_up_sync(KMC_FULL_MASK, v, d);
        if (lane >= d) v = o + v;
    }
    return v;
}

// device-wide helpers (scan.cu)
int kmc_exclusive_scan_i32(kmcb200_ctx *ctx, const int *in, int *out, long long n, int scratch_slot);
int kmc_bbox(kmcb200_ctx *ctx, const double *x, const double *y, const double *z, int first, int count,
             double *bbox_host6);

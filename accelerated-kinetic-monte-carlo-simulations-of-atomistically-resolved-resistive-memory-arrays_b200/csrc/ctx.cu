// ctx.cu -- context, error reporting, memory helpers of the C ABI (include/kmc_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void kmc_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static long long g_launches = 0;
void kmc_count_launch() { ++g_launches; }
extern "C" long long kmcb200_launch_count(void) { return g_launches; }

extern "C" const char *kmcb200_last_error(void) { return g_err; }
extern "C" int kmcb200_version(void) { return KMCB200_VERSION; }

int kmc_scratch(kmcb200_ctx *ctx, int slot, size_t bytes, void **out) {
    ScratchBuf &b = ctx->scratch[slot];
    if (b.bytes < bytes) {
        if (b.ptr) {
            KMC_CUDA(cudaStreamSynchronize(ctx->stream));
            KMC_CUDA(cudaFree(b.ptr));
            b.ptr = nullptr;
            b.bytes = 0;
        }
        size_t want = bytes + bytes / 8 + 256;
        KMC_CUDA(cudaMalloc(&b.ptr, want));
        b.bytes = want;
    }
    *out = b.ptr;
    return 0;
}

extern "C" int kmcb200_create(kmcb200_ctx **ctx_out, int device_ordinal, void *stream) {
    KMC_CHECK_ARG(ctx_out != nullptr, "ctx_out");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        kmc_set_error("no CUDA device available (%s); libkmc_b200 has no CPU fallback",
                      e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return KMCB200_E_NOGPU;
    }
    KMC_CHECK_ARG(device_ordinal >= 0 && device_ordinal < ndev, "device ordinal");
    KMC_CUDA(cudaSetDevice(device_ordinal));
    kmcb200_ctx *ctx = new kmcb200_ctx();
    ctx->device = device_ordinal;
    cudaDeviceProp prop;
    KMC_CUDA(cudaGetDeviceProperties(&prop, device_ordinal));
    ctx->sm_count = prop.multiProcessorCount;
    // NULL selects the CUDA legacy default stream (what torch.cuda.current_stream() is unless the caller
    // switched streams), so library work stays ordered with the caller's own default-stream work.
    ctx->stream = (cudaStream_t)stream;
    ctx->own_stream = false;
    KMC_CUDA(cudaHostAlloc(&ctx->h_mail, 4096, cudaHostAllocDefault));
    *ctx_out = ctx;
    return 0;
}

extern "C" int kmcb200_destroy(kmcb200_ctx *ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    kmc_coulomb_plan_free(ctx);
    for (auto &b : ctx->scratch)
        if (b.ptr) cudaFree(b.ptr);
    if (ctx->cg_state) cudaFree(ctx->cg_state);
    if (ctx->partials) cudaFree(ctx->partials);
    if (ctx->pcg_loop_ws) cudaFree(ctx->pcg_loop_ws);
    if (ctx->h_mail) cudaFreeHost(ctx->h_mail);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

extern "C" int kmcb200_set_stream(kmcb200_ctx *ctx, void *stream) {
    KMC_CHECK_ARG(ctx != nullptr, "ctx");
    if (ctx->own_stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
        ctx->own_stream = false;
    }
    ctx->stream = (cudaStream_t)stream;
    return 0;
}

extern "C" int kmcb200_synchronize(kmcb200_ctx *ctx) {
    KMC_CHECK_ARG(ctx != nullptr, "ctx");
    KMC_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int kmcb200_device_info(kmcb200_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem) {
    KMC_CHECK_ARG(ctx != nullptr, "ctx");
    cudaDeviceProp prop;
    KMC_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (total_mem) *total_mem = prop.totalGlobalMem;
    return 0;
}

extern "C" int kmcb200_malloc(kmcb200_ctx *ctx, void **dptr_out, size_t bytes) {
    KMC_CHECK_ARG(ctx && dptr_out, "ctx/dptr_out");
    KMC_CUDA(cudaSetDevice(ctx->device));
    KMC_CUDA(cudaMalloc(dptr_out, bytes ? bytes : 8));
    return 0;
}
extern "C" int kmcb200_free(kmcb200_ctx *ctx, void *dptr) {
    KMC_CHECK_ARG(ctx != nullptr, "ctx");
    if (dptr) {
        KMC_CUDA(cudaStreamSynchronize(ctx->stream));
        KMC_CUDA(cudaFree(dptr));
    }
    return 0;
}
extern "C" int kmcb200_memcpy_h2d(kmcb200_ctx *ctx, void *dst, const void *src, size_t bytes) {
    KMC_CHECK_ARG(ctx != nullptr, "ctx");
    KMC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
extern "C" int kmcb200_memcpy_d2h(kmcb200_ctx *ctx, void *dst, const void *src, size_t bytes) {
    KMC_CHECK_ARG(ctx != nullptr, "ctx");
    KMC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
}
extern "C" int kmcb200_memcpy_d2d(kmcb200_ctx *ctx, void *dst, const void *src, size_t bytes) {
    KMC_CHECK_ARG(ctx != nullptr, "ctx");
    KMC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}
extern "C" int kmcb200_memset(kmcb200_ctx *ctx, void *dst, int value, size_t bytes) {
    KMC_CHECK_ARG(ctx != nullptr, "ctx");
    KMC_CUDA(cudaMemsetAsync(dst, value, bytes, ctx->stream));
    return 0;
}
extern "C" int kmcb200_host_alloc_pinned(void **hptr_out, size_t bytes) {
    KMC_CHECK_ARG(hptr_out != nullptr, "hptr_out");
    KMC_CUDA(cudaHostAlloc(hptr_out, bytes ? bytes : 8, cudaHostAllocDefault));
    return 0;
}
extern "C" int kmcb200_host_free_pinned(void *hptr) {
    if (hptr) KMC_CUDA(cudaFreeHost(hptr));
    return 0;
}


// ---- measurement helper: FP64 FMA peak of this device (the roofline denominator of the Coulomb sum; MEASURED_PEAKS.json
// has no FP64 entry).  8 independent FMA chains per thread, 256 threads per CTA, 16 CTAs per SM.
__global__ void __launch_bounds__(256) fp64_peak_kernel(int iters, double seed, double *sink) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == -1.2345) sink[0] = r;  // never true: keeps the chains alive
}

extern "C" int kmcb200_fp64_peak(kmcb200_ctx *ctx, double *tflops_host) {
    KMC_CHECK_ARG(ctx && tflops_host, "arguments");
    void *sink = nullptr;
    KMC_TRY(kmc_scratch(ctx, 11, 64, &sink));
    const int iters = 1 << 14, blocks = ctx->sm_count * 16;
    cudaEvent_t e0, e1;
    KMC_CUDA(cudaEventCreate(&e0));
    KMC_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {  // first repetition = warm-up
        KMC_CUDA(cudaEventRecord(e0, ctx->stream));
        kmc_count_launch();
        fp64_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(iters, 1.0 + rep, (double *)sink);
        KMC_CUDA(cudaEventRecord(e1, ctx->stream));
        KMC_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        KMC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
        if (rep > 0 && ms > 0.f && flops / (ms * 1e-3) > best) best = flops / (ms * 1e-3);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops_host = best * 1e-12;
    return 0;
}

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
    for (auto &b : ctx->scratch)
        if (b.ptr) cudaFree(b.ptr);
    if (ctx->cg_state) cudaFree(ctx->cg_state);
    if (ctx->partials) cudaFree(ctx->partials);
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

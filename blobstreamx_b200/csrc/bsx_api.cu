// Context management for libbsx (see include/bsx.h).
#include "common.cuh"

#include <stdlib.h>

#include <new>

extern "C" int bsx_version(void) { return BSX_VERSION; }

// stream2 carries the Ed25519 half of bsx_header_range.  That kernel is latency-bound with few, fat CTAs (202
// registers/thread); giving it the highest priority makes the block scheduler place its CTAs first, and the SHA-256
// map kernel (ALU pipe, thousands of small CTAs) fills the rest of every SM around them.
static cudaError_t create_priority_stream(cudaStream_t *st) {
    int lo = 0, hi = 0;
    cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (e != cudaSuccess) return e;
    return cudaStreamCreateWithPriority(st, cudaStreamNonBlocking, hi);
}

extern "C" int bsx_init(int device, bsx_ctx **out) {
    if (!out) return BSX_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) return BSX_ERR_NODEVICE;
    bsx_ctx *ctx = new (std::nothrow) bsx_ctx();
    if (!ctx) return BSX_ERR_NOMEM;
    memset(ctx, 0, sizeof *ctx);
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        create_priority_stream(&ctx->stream2) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fork2, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join2, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return BSX_ERR_CUDA;
    }
    {   // stream-ordered scratch (cudaMallocAsync in k_ed25519.cu) stays cached in the pool across synchronizes
        cudaMemPool_t pool;
        uint64_t keep = UINT64_MAX;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    for (int i = 0; i < BSX_PIPE_STREAMS; i++) {
        if (cudaStreamCreateWithFlags(&ctx->pipe[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev_pipe[i], cudaEventDisableTiming) != cudaSuccess) {
            bsx_destroy(ctx);
            return BSX_ERR_CUDA;
        }
    }
    *out = ctx;
    return BSX_OK;
}

extern "C" void bsx_destroy(bsx_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
    }
    if (ctx->stream2) {
        cudaStreamSynchronize(ctx->stream2);
        cudaStreamDestroy(ctx->stream2);
    }
    for (int i = 0; i < BSX_PIPE_STREAMS; i++) {
        if (ctx->pipe[i]) {
            cudaStreamSynchronize(ctx->pipe[i]);
            cudaStreamDestroy(ctx->pipe[i]);
        }
        if (ctx->ev_pipe[i]) cudaEventDestroy(ctx->ev_pipe[i]);
    }
    for (uint32_t i = 0; i < ctx->n_ev_chunk; i++) cudaEventDestroy(ctx->ev_chunk[i]);
    free(ctx->ev_chunk);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_fork2) cudaEventDestroy(ctx->ev_fork2);
    if (ctx->ev_join2) cudaEventDestroy(ctx->ev_join2);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->ed_table) cudaFree(ctx->ed_table);
    delete ctx;
}

extern "C" const char *bsx_last_error(const bsx_ctx *ctx) { return ctx ? ctx->err : "null ctx"; }

extern "C" uint64_t bsx_launch_count(const bsx_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int bsx_sync(bsx_ctx *ctx) {
    BSX_REQUIRE(ctx, ctx != nullptr);
    BSX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BSX_OK;
}

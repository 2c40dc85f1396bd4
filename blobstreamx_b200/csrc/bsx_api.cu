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

static const struct { const char *name; int dflt; } TUNABLES[BSX_TUN_COUNT] = {
    {"ED_MODE", 0},      {"ED_QUAD_MAX", 16384}, {"ED_INLINE", -1},  {"ED_OCC", 0},     {"ED_REGS", -1},
    {"ED_FP64", -1},     {"HR_HASH_STREAM", -1}, {"HR_TRACE", 0},    {"PIPE_CHUNK", 0}, {"PIPE_ED", 0},
    {"PIPE_TRACE", 0},   {"PROOFS_OCC", 8},      {"SUBCHAIN_FUSED", 0}, {"COMMIT_THREADS", 128},
    {"ED_KEYTAB", -1},   {"ED_KOCC", 0},        {"ED_PAIR", -1},       {"ED_RESIDENT", 0},
    {"ED_TRACE_LANES", 0},
};

static int tunable_index(const char *name) {
    if (!name) return -1;
    if (!strncmp(name, "BSX_", 4)) name += 4;
    for (int i = 0; i < BSX_TUN_COUNT; i++)
        if (!strcmp(name, TUNABLES[i].name)) return i;
    return -1;
}

extern "C" int bsx_set_tunable(bsx_ctx *ctx, const char *name, int value) {
    BSX_REQUIRE(ctx, ctx != nullptr);
    const int i = tunable_index(name);
    if (i < 0) return bsx::fail(ctx, BSX_ERR_INVALID, "unknown tunable %s%s", name ? name : "(null)");
    ctx->tun[i] = value;
    return BSX_OK;
}

extern "C" int bsx_get_tunable(const bsx_ctx *ctx, const char *name, int *value) {
    const int i = tunable_index(name);
    if (!ctx || !value || i < 0) return BSX_ERR_INVALID;
    *value = ctx->tun[i];
    return BSX_OK;
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
    for (int i = 0; i < BSX_TUN_COUNT; i++) {
        char var[64];
        snprintf(var, sizeof var, "BSX_%s", TUNABLES[i].name);
        const char *e = getenv(var);
        ctx->tun[i] = e ? atoi(e) : TUNABLES[i].dflt;
    }
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
    if (ctx->ev_table) cudaEventDestroy(ctx->ev_table);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->ed_table) cudaFree(ctx->ed_table);
    bsx_plonk_cache_free(ctx);
    delete ctx;
}

extern "C" const char *bsx_last_error(const bsx_ctx *ctx) { return ctx ? ctx->err : "null ctx"; }

extern "C" uint64_t bsx_launch_count(const bsx_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int bsx_sync(bsx_ctx *ctx) {
    BSX_REQUIRE(ctx, ctx != nullptr);
    BSX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BSX_OK;
}

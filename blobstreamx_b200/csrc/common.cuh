// Internal helpers shared by the kernels: ctx, error plumbing, TMA (cp.async.bulk) staging.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/bsx.h"

#define BSX_PIPE_STREAMS 3

// Measurement knobs.  Defaults are the measured best; the environment variable BSX_<name> overrides a default when the
// ctx is created and bsx_set_tunable(ctx, "<name>", v) changes it afterwards (per ctx, so every path can be exercised
// in one process).  DESIGN.md section 4 lists what each one selects.
enum bsx_tun_id {
    BSX_TUN_ED_MODE,         // 0 by batch size, 1 three-stage quad-lane path, 2 one thread per signature
    BSX_TUN_ED_QUAD_MAX,     // largest batch that takes the quad-lane path (16384)
    BSX_TUN_ED_INLINE,       // thread-per-signature kernel: -1 by call site, 0 compact, 1 inlined point arithmetic
    BSX_TUN_ED_OCC,          // 0 default (4 CTAs/SM), 6 -> 168 registers, 8 -> 128 registers
    BSX_TUN_ED_REGS,         // 192-register build beside the hash kernels: -1 by wave fill, 0 never, >0 always
    BSX_TUN_ED_FP64,         // field arithmetic of the scalar multiplications on the FP64 pipe: -1 default, 0 off, 1 on
    BSX_TUN_HR_HASH_STREAM,  // skip hashes: -1 by wave fill, 0 caller's stream, 1 own stream
    BSX_TUN_HR_TRACE,        // device-resident step: print when each half finishes (synchronises)
    BSX_TUN_PIPE_CHUNK,      // host path: ranges per chunk (0 = growing chunks)
    BSX_TUN_PIPE_ED,         // host path: 0/1 skip half first, 2 last
    BSX_TUN_PIPE_TRACE,      // host path: per-chunk timeline on stderr
    BSX_TUN_PROOFS_OCC,      // proofs kernel: 8 (64 registers) or 6 (80 registers)
    BSX_TUN_SUBCHAIN_FUSED,  // 1 = one-CTA-per-job map kernel with TMA-staged inputs (A/B only)
    BSX_TUN_COMMIT_THREADS,  // threads per CTA of the commit kernel (B = 32 / 64): 128, 256, 512 or 1024
    BSX_TUN_ED_KEYTAB,       // per-key window tables for h*A (FP64 build): -1 when keys repeat enough, 0 never, 1 whenever they fit
    BSX_TUN_ED_KOCC,         // register budget of the table-path kernel: 0 = as the general kernel's, 4 / 6 / 8 CTAs per SM
    BSX_TUN_ED_PAIR,         // table path: two signatures per thread with one shared inversion: -1 by batch size, 0 never, 1 always
    BSX_TUN_ED_RESIDENT,     // thread-per-signature kernel: at most this many CTAs per SM (dynamic shared memory as ballast), 0 = no limit
    BSX_TUN_ED_TRACE_LANES,  // Ed25519 trace, lanes per multiplication in the chain kernel: 0 by batch size, 1 or 8
    BSX_TUN_COUNT
};

struct bsx_ctx {
    int device;
    int sm_count;
    cudaStream_t stream;       // the ctx's own stream (host entry points run here)
    char err[512];
    uint64_t launches;
    uint8_t *ws;               // grow-only device workspace for the host entry points
    size_t ws_cap;
    size_t ws_off;
    void *ed_table;            // s*G window table (k_ed25519.cu), built on first use
    cudaEvent_t ev_table;      // recorded after the table build: every consumer stream waits on it (the build runs once,
    int ed_table_pending;      // on whichever stream made the first Ed25519 call); pending until the event has been seen complete
    struct bsx_plonk_cache *plonk;   // twiddle tables of the transforms (k_plonk.cu), built on first use
    int tun[BSX_TUN_COUNT];    // measurement knobs (bsx_set_tunable; defaults from the BSX_* environment at bsx_init)
    cudaStream_t stream2;      // second stream + events: bsx_header_range runs its two halves concurrently
    cudaEvent_t ev_fork, ev_join;     // bsx_header_range (host path: two copy+compute pipelines)
    cudaEvent_t ev_fork2, ev_join2;   // verify_*: Ed25519 kernel on stream2 beside the SHA-256 schedule
    // bsx_header_range (host buffers): the map half is cut into chunks of ranges that flow through three streams by
    // role (0: H2D copies, 1: kernels, 2: D2H copies), so the H2D copy of chunk k+1, the kernels of chunk k and the
    // D2H copy of chunk k-1 run at the same time; ev_chunk[2k] = chunk k uploaded, ev_chunk[2k+1] = chunk k computed
    cudaStream_t pipe[BSX_PIPE_STREAMS];
    cudaEvent_t ev_pipe[BSX_PIPE_STREAMS];
    cudaEvent_t *ev_chunk;
    uint32_t n_ev_chunk;
};

// Does a batch of n signatures fill whole waves of the thread-per-signature Ed25519 kernel at `occ` CTAs of 64 per SM?  Then
// every SM's register file is full for the kernel's whole run (last wave at least 90 % full; 378 ranges: 591 of 592 CTAs).
static inline bool bsx_ed_fills_waves_at(const bsx_ctx *ctx, uint64_t n, int occ) {
    const uint64_t wave = (uint64_t)ctx->sm_count * occ, ctas = (n + 63) / 64, tail = ctas % wave;
    return ctas * 10 >= wave * 9 && (tail == 0 || tail * 10 >= wave * 9);
}
// The register budget (CTAs per SM: 8 -> 128 registers, 6 -> 168, 4 -> 216 / 240) whose waves the batch fills, largest first;
// 0 if none.  r02g (profiles/r02g_ed_occ_wave_sizes.txt, FP64 limbs): a build with more resident warps only pays when the
// batch fills ITS wave -- 757 ranges per step: 128-register build 156.1 M headers/s against 151.1 M with the 240-register
// one; 568 ranges: 168 registers 151.8 M against 150.0 M; at 378 ranges (one wave of 4 CTAs per SM) the wider builds
// would run two thirds or half empty.
static inline int bsx_ed_wave_occ(const bsx_ctx *ctx, uint64_t n) {
    if (bsx_ed_fills_waves_at(ctx, n, 8)) return 8;
    if (bsx_ed_fills_waves_at(ctx, n, 6)) return 6;
    if (bsx_ed_fills_waves_at(ctx, n, 4)) return 4;
    return 0;
}
// The step is arranged to leave room for the SHA-256 kernels beside a full Ed25519 wave (skip hashes on their own stream:
// +1..2.6 % at 378 / 756 / 1134 ranges per step); with a partly filled last wave the SMs have room anyway and the same
// arrangement costs 2..4 % (256 / 512 ranges).
static inline bool bsx_ed_fills_waves(const bsx_ctx *ctx, uint64_t n) { return bsx_ed_wave_occ(ctx, n) != 0; }

void bsx_plonk_cache_free(bsx_ctx *ctx);   // k_plonk.cu

// stages of verify_* (k_verify.cu), reused by bsx_header_range_dev
int bsx_verify_launch_ed(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, const uint8_t *validators, uint8_t *ed_out, int ed_corun = 0);
int bsx_verify_skip_corun_dev(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, const bsx_header_in *hdr, const uint8_t *validators,
                              const bsx_skip_in *skip, const uint8_t *trusted_pubkeys, const uint64_t *trusted_powers,
                              const uint32_t *trusted_byte_lengths, uint8_t *digests, uint8_t *ed_out, uint32_t *fail);
int bsx_verify_launch_hash(bsx_ctx *ctx, void *stream, int mode, uint32_t n, uint32_t N, const bsx_header_in *hdr,
                           const uint8_t *validators, const bsx_skip_in *skip, const uint8_t *trusted_pubkeys,
                           const uint64_t *trusted_powers, const uint32_t *trusted_byte_lengths, const bsx_step_in *step,
                           uint8_t *digests, uint8_t *data_commitments, uint32_t *fail);
int bsx_verify_launch_flags(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, const uint8_t *ed_out, uint32_t *fail);

namespace bsx {

inline int fail(bsx_ctx *ctx, int code, const char *fmt, const char *a = "", const char *b = "") {
    if (ctx) snprintf(ctx->err, sizeof ctx->err, fmt, a, b);
    return code;
}

#define BSX_CUDA(ctx, call)                                                                           \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return bsx::fail((ctx), BSX_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

#define BSX_LAUNCHED(ctx)                                                                             \
    do {                                                                                              \
        (ctx)->launches++;                                                                            \
        cudaError_t e_ = cudaGetLastError();                                                          \
        if (e_ != cudaSuccess) return bsx::fail((ctx), BSX_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(e_)); \
    } while (0)

#define BSX_REQUIRE(ctx, cond)                                                                        \
    do {                                                                                              \
        if (!(cond)) return bsx::fail((ctx), BSX_ERR_INVALID, "invalid argument: %s", #cond);         \
    } while (0)

// One shared-memory carveout for every kernel that can share an SM with another one.  An SM re-partitions its
// L1/shared memory only when it is empty, so two kernels whose preferred carveouts differ cannot be co-resident: the
// Ed25519 kernel (no shared memory, prefers all-L1) then keeps the SHA-256 commit/reduce/verify kernels (shared-memory
// trees) off every SM it occupies.  Pinning all of them to the same preference lets the halves of bsx_header_range
// really overlap.  BSX_CARVEOUT overrides the percentage (negative = leave the driver's default per kernel).
template <typename K>
inline void pin_carveout(K kernel) {
    static const int pct = [] { const char *e = getenv("BSX_CARVEOUT"); return e ? atoi(e) : 50; }();
    if (pct >= 0) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}
#define BSX_PIN_CARVEOUT(kernel)                       \
    do {                                               \
        static const bool once_ = (bsx::pin_carveout(kernel), true); \
        (void)once_;                                   \
    } while (0)

// ---- workspace (host entry points only) ----
inline int ws_begin(bsx_ctx *ctx, size_t total) {
    total += 4096;
    if (total > ctx->ws_cap) {
        if (ctx->ws) cudaFree(ctx->ws);
        ctx->ws = nullptr;
        ctx->ws_cap = 0;
        size_t cap = total + total / 4;
        cudaError_t e = cudaMalloc(&ctx->ws, cap);
        if (e != cudaSuccess) return fail(ctx, BSX_ERR_NOMEM, "cudaMalloc workspace: %s", cudaGetErrorString(e));
        ctx->ws_cap = cap;
    }
    ctx->ws_off = 0;
    return BSX_OK;
}
template <typename T>
inline T *ws_take(bsx_ctx *ctx, size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    T *p = reinterpret_cast<T *>(ctx->ws + ctx->ws_off);
    ctx->ws_off += bytes;
    return p;
}
inline size_t ws_size(size_t bytes) { return (bytes + 255) & ~size_t(255); }

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    }
}
// shared -> global bulk store (async proxy); caller fences + waits
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace bsx

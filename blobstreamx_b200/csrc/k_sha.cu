// K1: batched SHA-256 / SHA-512 over variable-length messages (one thread per message).
// Replaces the per-request HashDigestHint loop (PX/frontend/hash/curta/builder.rs:26-49,
// PX/frontend/hash/curta/digest_hint.rs:30-38).
#include "common.cuh"
#include "sha256.cuh"
#include "sha512.cuh"

namespace bsx {

__global__ void __launch_bounds__(128) sha256_batch_kernel(const uint8_t *__restrict__ msgs,
                                                           const uint32_t *__restrict__ offsets, uint32_t n,
                                                           uint8_t *__restrict__ digests) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t off = offsets[i], len = offsets[i + 1] - off;
    const uint8_t *m = msgs + off;
    uint32_t st[8];
    if ((reinterpret_cast<uintptr_t>(m) & 3) == 0) {
        // word-aligned message: 4-byte loads, byte order fixed with PRMT
        const uint32_t *mw = reinterpret_cast<const uint32_t *>(m);
        sha256_init(st);
        uint32_t nblk = (len + 9 + 63) >> 6;
        for (uint32_t b = 0; b < nblk; b++) {
            uint32_t w[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                uint32_t base = b * 64 + k * 4;
                uint32_t v;
                if (base + 4 <= len) {
                    v = bswap32(__ldg(mw + (base >> 2)));
                } else {
                    v = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        uint32_t idx = base + j;
                        uint32_t byte = idx < len ? (uint32_t)__ldg(m + idx) : (idx == len ? 0x80u : 0u);
                        v = (v << 8) | byte;
                    }
                }
                w[k] = v;
            }
            if (b == nblk - 1) { w[14] = 0; w[15] = len << 3; }
            sha256_compress(st, w);
        }
    } else {
        sha256_bytes([&](uint32_t k) -> uint8_t { return __ldg(m + k); }, len, st);
    }
    store_digest_be(digests + 32 * (size_t)i, st);
}

__global__ void __launch_bounds__(128) sha512_batch_kernel(const uint8_t *__restrict__ msgs,
                                                           const uint32_t *__restrict__ offsets, uint32_t n,
                                                           uint8_t *__restrict__ digests) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t off = offsets[i], len = offsets[i + 1] - off;
    const uint8_t *m = msgs + off;
    uint64_t st[8];
    sha512_bytes([&](uint32_t k) -> uint8_t { return __ldg(m + k); }, len, st);
    uint2 *out = reinterpret_cast<uint2 *>(digests + 64 * (size_t)i);
#pragma unroll
    for (int k = 0; k < 8; k++) out[k] = make_uint2(bswap32((uint32_t)(st[k] >> 32)), bswap32((uint32_t)st[k]));
}

}  // namespace bsx

extern "C" int bsx_sha256_batch_dev(bsx_ctx *ctx, void *stream, const uint8_t *msgs, const uint32_t *offsets,
                                    uint32_t n, uint8_t *digests) {
    BSX_REQUIRE(ctx, ctx && offsets && digests);
    if (n == 0) return BSX_OK;
    BSX_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(digests) & 15) == 0);
    bsx::sha256_batch_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(msgs, offsets, n, digests);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

extern "C" int bsx_sha512_batch_dev(bsx_ctx *ctx, void *stream, const uint8_t *msgs, const uint32_t *offsets,
                                    uint32_t n, uint8_t *digests) {
    BSX_REQUIRE(ctx, ctx && offsets && digests);
    if (n == 0) return BSX_OK;
    BSX_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(digests) & 7) == 0);
    bsx::sha512_batch_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(msgs, offsets, n, digests);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

static int sha_batch_host(bsx_ctx *ctx, const uint8_t *msgs, const uint32_t *offsets, uint32_t n, uint8_t *digests,
                          int dsz) {
    BSX_REQUIRE(ctx, ctx && offsets && digests);
    if (n == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t total = offsets[n];
    BSX_REQUIRE(ctx, total == 0 || msgs);
    int rc = bsx::ws_begin(ctx, bsx::ws_size(total + 16) + bsx::ws_size(4 * (size_t)(n + 1)) + bsx::ws_size((size_t)n * dsz));
    if (rc) return rc;
    uint8_t *d_msgs = bsx::ws_take<uint8_t>(ctx, total + 16);
    uint32_t *d_off = bsx::ws_take<uint32_t>(ctx, n + 1);
    uint8_t *d_dig = bsx::ws_take<uint8_t>(ctx, (size_t)n * dsz);
    if (total) BSX_CUDA(ctx, cudaMemcpyAsync(d_msgs, msgs, total, cudaMemcpyHostToDevice, ctx->stream));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_off, offsets, 4 * (size_t)(n + 1), cudaMemcpyHostToDevice, ctx->stream));
    rc = dsz == 32 ? bsx_sha256_batch_dev(ctx, ctx->stream, d_msgs, d_off, n, d_dig)
                   : bsx_sha512_batch_dev(ctx, ctx->stream, d_msgs, d_off, n, d_dig);
    if (rc) return rc;
    BSX_CUDA(ctx, cudaMemcpyAsync(digests, d_dig, (size_t)n * dsz, cudaMemcpyDeviceToHost, ctx->stream));
    BSX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BSX_OK;
}

extern "C" int bsx_sha256_batch(bsx_ctx *ctx, const uint8_t *msgs, const uint32_t *offsets, uint32_t n,
                                uint8_t *digests) {
    return sha_batch_host(ctx, msgs, offsets, n, digests, 32);
}
extern "C" int bsx_sha512_batch(bsx_ctx *ctx, const uint8_t *msgs, const uint32_t *offsets, uint32_t n,
                                uint8_t *digests) {
    return sha_batch_host(ctx, msgs, offsets, n, digests, 64);
}

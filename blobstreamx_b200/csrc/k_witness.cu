// Witness-side data formats (SURVEY 8a row a18 and 8b "data convention"):
//   * HashInputData -- the STARK public-input layout plonky2x derives from an accelerator's request list:
//     padded_chunks / end_bits / digest_bits / digest_indices   PX/frontend/hash/curta/mod.rs:95-192,
//     stream order PX/frontend/hash/curta/data.rs:63-78; padding PX/frontend/hash/sha/sha256/pad.rs:15-157,
//     PX/frontend/hash/sha/sha512/pad.rs:13-58.
//   * value <-> field-element encodings: ByteVariable = 8 big-endian bit elements (PX/frontend/vars/byte.rs:49-66),
//     SHA-256 digest = 8 big-endian u32 words, one element each (PX/frontend/hash/sha/sha256/curta.rs:81-92).
// These are pure layout kernels: one thread per output word, fully coalesced stores; the bit expansion writes
// 64 bytes per input byte and is HBM-write bound.
#include "common.cuh"

namespace bsx {

// CHUNK = 64 (SHA-256, u32 words) or 128 (SHA-512, u64 words)
template <int CHUNK, typename Word>
__global__ void __launch_bounds__(256) hash_input_data_kernel(uint32_t n_req, const uint8_t *__restrict__ bufs,
                                                              const uint32_t *__restrict__ buf_offsets,
                                                              const uint32_t *__restrict__ lens, const uint8_t *__restrict__ kinds,
                                                              const uint32_t *__restrict__ chunk_offsets, Word *__restrict__ padded,
                                                              uint8_t *__restrict__ end_bits, uint8_t *__restrict__ digest_bits,
                                                              uint32_t *__restrict__ digest_indices) {
    constexpr int WB = sizeof(Word), WPC = CHUNK / WB, LENB = CHUNK / 8;  // length field: 8 (16) bytes, low 8 used
    const uint32_t total = chunk_offsets[n_req];
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)total * WPC) return;
    const uint32_t c = (uint32_t)(idx / WPC), w = (uint32_t)(idx % WPC);
    // request owning chunk c: last r with chunk_offsets[r] <= c
    uint32_t lo = 0, hi = n_req;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (chunk_offsets[mid] <= c) lo = mid; else hi = mid;
    }
    const uint32_t r = lo, j = c - chunk_offsets[r], nch = chunk_offsets[r + 1] - chunk_offsets[r];
    const uint8_t *b = bufs + buf_offsets[r];
    const uint32_t blen = buf_offsets[r + 1] - buf_offsets[r];
    const bool variable = kinds[r] != 0;
    const uint32_t len = variable ? lens[r] : blen;
    const uint32_t lc = variable ? (len + LENB) / CHUNK : nch - 1;   // chunk that carries the bit length / the digest
    const uint64_t bits = (uint64_t)len * 8;
    Word v = 0;
#pragma unroll
    for (int k = 0; k < WB; k++) {
        const uint32_t p = j * CHUNK + w * WB + k;
        uint32_t byte = (p < len && p < blen) ? b[p] : (p == len ? 0x80u : 0u);
        const uint32_t q = p % CHUNK;
        if (j == lc && q >= CHUNK - 8) byte = (uint32_t)(bits >> (8 * (CHUNK - 1 - q))) & 0xffu;
        v = (Word)(v << 8) | (Word)byte;
    }
    padded[idx] = v;
    if (w == 0) {
        end_bits[c] = (j == nch - 1);
        digest_bits[c] = (j == lc);
        if (j == 0) digest_indices[r] = chunk_offsets[r] + lc;
    }
}

// bytes -> 8 big-endian bit elements each (ByteVariable): the witness is 64x the payload, a pure HBM-write stream.
// One thread per PAIR of elements (16 bytes): consecutive lanes write consecutive 16-byte pieces, so every store
// instruction of a warp is one contiguous 512-byte segment; the four lanes of a byte read it once through L1.
__global__ void __launch_bounds__(256) pack_bytes_kernel(const uint8_t *__restrict__ bytes, size_t n, uint64_t *__restrict__ el) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // pair index: byte t / 4, bits 7-2k and 6-2k, k = t % 4
    if (t >= 4 * n) return;
    const uint32_t b = __ldg(bytes + (t >> 2)), k = (uint32_t)t & 3u;
    __stcs(reinterpret_cast<ulonglong2 *>(el) + t, make_ulonglong2((b >> (7 - 2 * k)) & 1, (b >> (6 - 2 * k)) & 1));
}
// inverse; bad[0] is set to 1 when some element is not a bit.  Same mapping: a lane loads 16 contiguous bytes, the four
// lanes of a byte combine their two bits with shuffles.
__global__ void __launch_bounds__(256) unpack_bytes_kernel(const uint64_t *__restrict__ el, size_t n, uint8_t *__restrict__ bytes,
                                                           uint32_t *__restrict__ bad) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < 4 * n;
    ulonglong2 e = make_ulonglong2(0, 0);
    if (live) e = __ldcs(reinterpret_cast<const ulonglong2 *>(el) + t);
    const uint32_t k = (uint32_t)t & 3u;
    uint32_t v = ((uint32_t)(e.x & 1) << (7 - 2 * k)) | ((uint32_t)(e.y & 1) << (6 - 2 * k));
    uint32_t nb = (e.x > 1 || e.y > 1) ? 1u : 0u;
    v |= __shfl_xor_sync(0xffffffffu, v, 1);
    v |= __shfl_xor_sync(0xffffffffu, v, 2);
    if (live && k == 0) bytes[t >> 2] = (uint8_t)v;
    nb = __any_sync(0xffffffffu, nb);
    if (nb && bad && (threadIdx.x & 31) == 0) atomicOr(bad, 1u);
}
// big-endian 4-byte (8-byte) groups -> one element each ([U32Variable; 8] digests; u64 state words as (lo, hi) u32 limbs)
__global__ void __launch_bounds__(256) pack_u32_be_kernel(const uint8_t *__restrict__ bytes, size_t n_words, uint64_t *__restrict__ el) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    const uint8_t *p = bytes + 4 * i;
    el[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

}  // namespace bsx

using namespace bsx;

static uint32_t chunks_of(uint32_t blen, int variable, int chunk) {
    const uint32_t lenb = chunk / 8 + 1;  // 0x80 + length field
    if (!variable) return (blen + lenb + chunk - 1) / chunk;
    // sha256: the buffer is first rounded up to whole chunks (sha256/curta.rs:154-161); sha512: used as is
    const uint32_t eff = chunk == 64 ? ((blen + 63) / 64) * 64 : blen;
    return (eff + lenb + chunk - 1) / chunk;
}

extern "C" uint32_t bsx_hash_input_chunks(int sha512, uint32_t buf_len, int variable) {
    return chunks_of(buf_len, variable, sha512 ? 128 : 64);
}

extern "C" int bsx_hash_input_data_dev(bsx_ctx *ctx, void *stream, int sha512, uint32_t n_req, const uint8_t *bufs,
                                       const uint32_t *buf_offsets, const uint32_t *lens, const uint8_t *kinds,
                                       const uint32_t *chunk_offsets, uint32_t total_chunks, void *padded_chunks,
                                       uint8_t *end_bits, uint8_t *digest_bits, uint32_t *digest_indices) {
    BSX_REQUIRE(ctx, ctx && buf_offsets && lens && kinds && chunk_offsets && padded_chunks && end_bits && digest_bits && digest_indices);
    if (n_req == 0 || total_chunks == 0) return BSX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (sha512) {
        const size_t words = (size_t)total_chunks * 16;
        hash_input_data_kernel<128, uint64_t><<<(unsigned)((words + 255) / 256), 256, 0, st>>>(
            n_req, bufs, buf_offsets, lens, kinds, chunk_offsets, reinterpret_cast<uint64_t *>(padded_chunks), end_bits, digest_bits,
            digest_indices);
    } else {
        const size_t words = (size_t)total_chunks * 16;
        hash_input_data_kernel<64, uint32_t><<<(unsigned)((words + 255) / 256), 256, 0, st>>>(
            n_req, bufs, buf_offsets, lens, kinds, chunk_offsets, reinterpret_cast<uint32_t *>(padded_chunks), end_bits, digest_bits,
            digest_indices);
    }
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

extern "C" int bsx_hash_input_data(bsx_ctx *ctx, int sha512, uint32_t n_req, const uint8_t *bufs, const uint32_t *buf_offsets,
                                   const uint32_t *lens, const uint8_t *kinds, void *padded_chunks, uint8_t *end_bits,
                                   uint8_t *digest_bits, uint32_t *digest_indices, uint32_t *total_chunks) {
    BSX_REQUIRE(ctx, ctx && buf_offsets && lens && kinds && total_chunks);
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    const int chunk = sha512 ? 128 : 64;
    // chunk offsets are circuit constants (buffer lengths and request kinds are fixed at build time)
    uint32_t *h_off = new uint32_t[n_req + 1];
    h_off[0] = 0;
    for (uint32_t r = 0; r < n_req; r++) h_off[r + 1] = h_off[r] + chunks_of(buf_offsets[r + 1] - buf_offsets[r], kinds[r] != 0, chunk);
    const uint32_t total = h_off[n_req];
    *total_chunks = total;
    if (!padded_chunks || n_req == 0) { delete[] h_off; return BSX_OK; }  // size query
    const size_t s_buf = buf_offsets[n_req], s_pad = (size_t)total * chunk;
    int rc = ws_begin(ctx, ws_size(s_buf + 16) + 3 * ws_size(4 * (size_t)(n_req + 1)) + ws_size(n_req) + ws_size(s_pad) + 2 * ws_size(total) +
                               ws_size(4 * (size_t)n_req));
    if (rc) { delete[] h_off; return rc; }
    uint8_t *d_buf = ws_take<uint8_t>(ctx, s_buf + 16);
    uint32_t *d_bo = ws_take<uint32_t>(ctx, n_req + 1), *d_len = ws_take<uint32_t>(ctx, n_req + 1), *d_co = ws_take<uint32_t>(ctx, n_req + 1);
    uint8_t *d_kind = ws_take<uint8_t>(ctx, n_req), *d_pad = ws_take<uint8_t>(ctx, s_pad);
    uint8_t *d_eb = ws_take<uint8_t>(ctx, total), *d_db = ws_take<uint8_t>(ctx, total);
    uint32_t *d_di = ws_take<uint32_t>(ctx, n_req);
    cudaStream_t st = ctx->stream;
    cudaError_t e = cudaSuccess;
    if (s_buf) e = cudaMemcpyAsync(d_buf, bufs, s_buf, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_bo, buf_offsets, 4 * (size_t)(n_req + 1), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_len, lens, 4 * (size_t)n_req, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_kind, kinds, n_req, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_co, h_off, 4 * (size_t)(n_req + 1), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // h_off is freed below
    delete[] h_off;
    BSX_CUDA(ctx, e);
    rc = bsx_hash_input_data_dev(ctx, st, sha512, n_req, d_buf, d_bo, d_len, d_kind, d_co, total, d_pad, d_eb, d_db, d_di);
    if (rc) return rc;
    BSX_CUDA(ctx, cudaMemcpyAsync(padded_chunks, d_pad, s_pad, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(end_bits, d_eb, total, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(digest_bits, d_db, total, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(digest_indices, d_di, 4 * (size_t)n_req, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaStreamSynchronize(st));
    return BSX_OK;
}

extern "C" int bsx_witness_pack_bytes_dev(bsx_ctx *ctx, void *stream, const uint8_t *bytes, size_t n, uint64_t *elements) {
    BSX_REQUIRE(ctx, ctx && (n == 0 || (bytes && elements)) && (reinterpret_cast<uintptr_t>(elements) & 15) == 0);
    if (n == 0) return BSX_OK;
    pack_bytes_kernel<<<(unsigned)((4 * n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bytes, n, elements);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}
extern "C" int bsx_witness_unpack_bytes_dev(bsx_ctx *ctx, void *stream, const uint64_t *elements, size_t n, uint8_t *bytes,
                                            uint32_t *not_bits) {
    BSX_REQUIRE(ctx, ctx && (n == 0 || (bytes && elements)) && (reinterpret_cast<uintptr_t>(elements) & 15) == 0);
    if (n == 0) return BSX_OK;
    unpack_bytes_kernel<<<(unsigned)((4 * n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(elements, n, bytes, not_bits);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}
extern "C" int bsx_witness_pack_u32_be_dev(bsx_ctx *ctx, void *stream, const uint8_t *bytes, size_t n_words, uint64_t *elements) {
    BSX_REQUIRE(ctx, ctx && (n_words == 0 || (bytes && elements)));
    if (n_words == 0) return BSX_OK;
    pack_u32_be_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bytes, n_words, elements);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

// host-buffer forms
extern "C" int bsx_witness_pack_bytes(bsx_ctx *ctx, const uint8_t *bytes, size_t n, uint64_t *elements) {
    BSX_REQUIRE(ctx, ctx && (n == 0 || (bytes && elements)));
    if (n == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = ws_begin(ctx, ws_size(n) + ws_size(64 * n));
    if (rc) return rc;
    uint8_t *d_b = ws_take<uint8_t>(ctx, n);
    uint64_t *d_e = ws_take<uint64_t>(ctx, 8 * n);
    BSX_CUDA(ctx, cudaMemcpyAsync(d_b, bytes, n, cudaMemcpyHostToDevice, ctx->stream));
    rc = bsx_witness_pack_bytes_dev(ctx, ctx->stream, d_b, n, d_e);
    if (rc) return rc;
    BSX_CUDA(ctx, cudaMemcpyAsync(elements, d_e, 64 * n, cudaMemcpyDeviceToHost, ctx->stream));
    BSX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BSX_OK;
}
extern "C" int bsx_witness_unpack_bytes(bsx_ctx *ctx, const uint64_t *elements, size_t n, uint8_t *bytes, uint32_t *not_bits) {
    BSX_REQUIRE(ctx, ctx && (n == 0 || (bytes && elements)));
    if (not_bits) *not_bits = 0;
    if (n == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = ws_begin(ctx, ws_size(n) + ws_size(64 * n) + ws_size(4));
    if (rc) return rc;
    uint8_t *d_b = ws_take<uint8_t>(ctx, n);
    uint64_t *d_e = ws_take<uint64_t>(ctx, 8 * n);
    uint32_t *d_bad = ws_take<uint32_t>(ctx, 1);
    BSX_CUDA(ctx, cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_e, elements, 64 * n, cudaMemcpyHostToDevice, ctx->stream));
    rc = bsx_witness_unpack_bytes_dev(ctx, ctx->stream, d_e, n, d_b, d_bad);
    if (rc) return rc;
    BSX_CUDA(ctx, cudaMemcpyAsync(bytes, d_b, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (not_bits) BSX_CUDA(ctx, cudaMemcpyAsync(not_bits, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
    BSX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BSX_OK;
}

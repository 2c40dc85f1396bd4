// K2/K3: Tendermint Merkle proofs and trees, and get_data_commitment<N>.
#include "common.cuh"
#include "sha256.cuh"
#include "tm_tree.cuh"

namespace bsx {

// K3: one thread per inclusion proof (PX/frontend/merkle/tendermint.rs:62-93).
__global__ void __launch_bounds__(128) tm_merkle_proofs_kernel(const uint8_t *__restrict__ leaves, uint32_t leaf_len,
                                                               const uint8_t *__restrict__ aunts, uint32_t depth,
                                                               const uint32_t *__restrict__ path_bits, uint32_t n,
                                                               int hashed_leaf, uint8_t *__restrict__ digests,
                                                               uint8_t *__restrict__ roots) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t nd = 2 * depth + (hashed_leaf ? 0 : 1);
    uint8_t *out = digests + 32 * (size_t)nd * i;
    uint32_t h[8];
    if (hashed_leaf) {
        const uint8_t *p = leaves + 32 * (size_t)i;
#pragma unroll
        for (int k = 0; k < 8; k++)
            h[k] = ((uint32_t)p[4 * k] << 24) | ((uint32_t)p[4 * k + 1] << 16) | ((uint32_t)p[4 * k + 2] << 8) | p[4 * k + 3];
    } else {
        const uint8_t *p = leaves + (size_t)leaf_len * i;
        tm_leaf_hash([&](uint32_t k) -> uint8_t { return __ldg(p + k); }, leaf_len, h);
        store_digest_be(out, h);
        out += 32;
    }
    uint32_t bits = path_bits[i];
    for (uint32_t l = 0; l < depth; l++) {
        uint32_t a[8], left[8], right[8];
        load_digest_be(aunts + 32 * ((size_t)i * depth + l), a);
        tm_inner_hash(h, a, left);
        tm_inner_hash(a, h, right);
        store_digest_be(out, left);
        store_digest_be(out + 32, right);
        out += 64;
        bool sel = (bits >> l) & 1;
#pragma unroll
        for (int k = 0; k < 8; k++) h[k] = sel ? right[k] : left[k];
    }
    store_digest_be(roots + 32 * (size_t)i, h);
}

// K2: one CTA per tree.  dynamic smem: P*32 + (P/2)*32 bytes.
__global__ void tm_merkle_tree_kernel(const uint8_t *__restrict__ leaf_digests, uint32_t N, uint32_t P,
                                      const uint64_t *__restrict__ nb_enabled, uint8_t *__restrict__ inner,
                                      uint8_t *__restrict__ roots) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *A = smem, *Bf = smem + 8 * (size_t)P;
    uint32_t t = blockIdx.x;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(leaf_digests + 32 * (size_t)N * t);
    for (uint32_t k = threadIdx.x; k < 8 * P; k += blockDim.x) A[k] = k < 8 * N ? bswap32(src[k]) : 0u;  // zero-digest padding
    __syncthreads();
    uint32_t root[8];
    tm_tree_cta(A, Bf, P, nb_enabled[t], inner ? inner + 32 * (size_t)(P - 1) * t : nullptr, root);
    if (threadIdx.x == 0) store_digest_be(roots + 32 * (size_t)t, root);
}

// get_data_commitment<N> (BX/circuits/builder.rs:105-148): tuple leaf hashes + tree, one CTA per tree.
__global__ void data_commitment_kernel(const uint8_t *__restrict__ data_hashes, uint32_t N, uint32_t P,
                                       const uint64_t *__restrict__ start_blocks, const uint64_t *__restrict__ end_blocks,
                                       uint8_t *__restrict__ digests, uint8_t *__restrict__ roots,
                                       uint32_t *__restrict__ fail) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *A = smem, *Bf = smem + 8 * (size_t)P;
    uint32_t t = blockIdx.x;
    uint64_t start = start_blocks[t], end = end_blocks[t];
    uint64_t nb_blocks = end - start;  // wrapping, as the U64 gadget
    uint8_t *out = digests + 32 * (size_t)(N + P - 1) * t;
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
        uint32_t d[8];
        if (i < N) {
            // leaf = 0x00 ‖ 0^24 ‖ u64be(start+i) ‖ data_hash  (65 bytes, builder.rs:82-103)
            const uint32_t *dh = reinterpret_cast<const uint32_t *>(data_hashes + 32 * ((size_t)N * t + i));
            uint64_t hgt = start + i;
            uint32_t x[8];
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = bswap32(__ldg(dh + k));
            uint32_t w[16];
#pragma unroll
            for (int k = 0; k < 6; k++) w[k] = 0;
            w[6] = (uint32_t)(hgt >> 40);                        // bytes 24..27 = 00 h7 h6 h5
            w[7] = (uint32_t)(hgt >> 8);                         // h4 h3 h2 h1
            w[8] = ((uint32_t)hgt << 24) | (x[0] >> 8);          // h0 d0 d1 d2
#pragma unroll
            for (int k = 1; k < 8; k++) w[8 + k] = __funnelshift_r(x[k], x[k - 1], 8);
            sha256_init(d);
            sha256_compress(d, w);
            sha256_tail65(d, x[7] & 0xffu);
            store_digest_be(out + 32 * (size_t)i, d);
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) d[k] = 0;
        }
#pragma unroll
        for (int k = 0; k < 8; k++) A[8 * i + k] = d[k];
    }
    __syncthreads();
    uint32_t root[8];
    tm_tree_cta(A, Bf, P, nb_blocks & 0xffffffffull, out + 32 * (size_t)N, root);
    if (threadIdx.x == 0) {
        store_digest_be(roots + 32 * (size_t)t, root);
        if (fail) fail[t] = (end < start || (nb_blocks >> 32)) ? BSX_FAIL_END_LT_START : 0u;
    }
}

static inline uint32_t pow2_ceil(uint32_t n) {
    uint32_t p = 1;
    while (p < n) p <<= 1;
    return p;
}

}  // namespace bsx

using namespace bsx;

extern "C" int bsx_tm_merkle_proofs_dev(bsx_ctx *ctx, void *stream, const uint8_t *leaves, uint32_t leaf_len,
                                        const uint8_t *aunts, uint32_t depth, const uint32_t *path_bits, uint32_t n,
                                        int hashed_leaf, uint8_t *digests, uint8_t *roots) {
    BSX_REQUIRE(ctx, ctx && leaves && path_bits && digests && roots && depth <= 32 && (depth == 0 || aunts));
    BSX_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(aunts) | reinterpret_cast<uintptr_t>(digests) |
                       reinterpret_cast<uintptr_t>(roots)) & 15) == 0);
    if (n == 0) return BSX_OK;
    tm_merkle_proofs_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(leaves, leaf_len, aunts, depth, path_bits,
                                                                               n, hashed_leaf, digests, roots);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

extern "C" int bsx_tm_merkle_proofs(bsx_ctx *ctx, const uint8_t *leaves, uint32_t leaf_len, const uint8_t *aunts,
                                    uint32_t depth, const uint32_t *path_bits, uint32_t n, int hashed_leaf,
                                    uint8_t *digests, uint8_t *roots) {
    BSX_REQUIRE(ctx, ctx && leaves && path_bits && digests && roots);
    if (n == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t ll = hashed_leaf ? 32 : leaf_len, nd = 2 * depth + (hashed_leaf ? 0 : 1);
    size_t s_leaf = ll * n, s_aunt = 32 * (size_t)depth * n, s_dig = 32 * nd * n, s_root = 32 * (size_t)n;
    int rc = ws_begin(ctx, ws_size(s_leaf) + ws_size(s_aunt) + ws_size(4 * (size_t)n) + ws_size(s_dig) + ws_size(s_root));
    if (rc) return rc;
    uint8_t *d_leaf = ws_take<uint8_t>(ctx, s_leaf), *d_aunt = ws_take<uint8_t>(ctx, s_aunt);
    uint32_t *d_bits = ws_take<uint32_t>(ctx, n);
    uint8_t *d_dig = ws_take<uint8_t>(ctx, s_dig), *d_root = ws_take<uint8_t>(ctx, s_root);
    BSX_CUDA(ctx, cudaMemcpyAsync(d_leaf, leaves, s_leaf, cudaMemcpyHostToDevice, ctx->stream));
    if (s_aunt) BSX_CUDA(ctx, cudaMemcpyAsync(d_aunt, aunts, s_aunt, cudaMemcpyHostToDevice, ctx->stream));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_bits, path_bits, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    rc = bsx_tm_merkle_proofs_dev(ctx, ctx->stream, d_leaf, leaf_len, d_aunt, depth, d_bits, n, hashed_leaf, d_dig, d_root);
    if (rc) return rc;
    if (s_dig) BSX_CUDA(ctx, cudaMemcpyAsync(digests, d_dig, s_dig, cudaMemcpyDeviceToHost, ctx->stream));
    BSX_CUDA(ctx, cudaMemcpyAsync(roots, d_root, s_root, cudaMemcpyDeviceToHost, ctx->stream));
    BSX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BSX_OK;
}

static int tree_launch_cfg(bsx_ctx *ctx, const void *kernel, uint32_t P, uint32_t *threads, size_t *smem) {
    *threads = P / 2 < 32 ? 32 : (P / 2 > 512 ? 512 : P / 2);  // <= 512 threads: up to 128 registers each
    *smem = 32 * (size_t)P + 16 * (size_t)P;
    if (*smem > 48 * 1024) BSX_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem));
    return BSX_OK;
}

extern "C" int bsx_tm_merkle_tree_dev(bsx_ctx *ctx, void *stream, const uint8_t *leaf_digests, uint32_t N, uint32_t t,
                                      const uint64_t *nb_enabled, uint8_t *inner, uint8_t *roots) {
    BSX_REQUIRE(ctx, ctx && leaf_digests && nb_enabled && roots && N >= 1 && N <= 4096);
    BSX_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(leaf_digests) & 3) | ((reinterpret_cast<uintptr_t>(inner) |
                      reinterpret_cast<uintptr_t>(roots)) & 15)) == 0);
    if (t == 0) return BSX_OK;
    uint32_t P = pow2_ceil(N), threads;
    size_t smem;
    int rc = tree_launch_cfg(ctx, (const void *)tm_merkle_tree_kernel, P, &threads, &smem);
    if (rc) return rc;
    tm_merkle_tree_kernel<<<t, threads, smem, (cudaStream_t)stream>>>(leaf_digests, N, P, nb_enabled, inner, roots);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

extern "C" int bsx_tm_merkle_tree(bsx_ctx *ctx, const uint8_t *leaf_digests, uint32_t N, uint32_t t,
                                  const uint64_t *nb_enabled, uint8_t *inner, uint8_t *roots) {
    BSX_REQUIRE(ctx, ctx && leaf_digests && nb_enabled && roots && N >= 1 && N <= 4096);
    if (t == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t P = pow2_ceil(N);
    size_t s_leaf = 32 * (size_t)N * t, s_inner = 32 * (size_t)(P - 1) * t, s_root = 32 * (size_t)t;
    int rc = ws_begin(ctx, ws_size(s_leaf) + ws_size(8 * (size_t)t) + ws_size(s_inner) + ws_size(s_root));
    if (rc) return rc;
    uint8_t *d_leaf = ws_take<uint8_t>(ctx, s_leaf);
    uint64_t *d_nb = ws_take<uint64_t>(ctx, t);
    uint8_t *d_inner = ws_take<uint8_t>(ctx, s_inner), *d_root = ws_take<uint8_t>(ctx, s_root);
    BSX_CUDA(ctx, cudaMemcpyAsync(d_leaf, leaf_digests, s_leaf, cudaMemcpyHostToDevice, ctx->stream));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_nb, nb_enabled, 8 * (size_t)t, cudaMemcpyHostToDevice, ctx->stream));
    rc = bsx_tm_merkle_tree_dev(ctx, ctx->stream, d_leaf, N, t, d_nb, d_inner, d_root);
    if (rc) return rc;
    if (inner && s_inner) BSX_CUDA(ctx, cudaMemcpyAsync(inner, d_inner, s_inner, cudaMemcpyDeviceToHost, ctx->stream));
    BSX_CUDA(ctx, cudaMemcpyAsync(roots, d_root, s_root, cudaMemcpyDeviceToHost, ctx->stream));
    BSX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BSX_OK;
}

extern "C" int bsx_data_commitment_batch_dev(bsx_ctx *ctx, void *stream, const uint8_t *data_hashes, uint32_t N,
                                             uint32_t t, const uint64_t *start_blocks, const uint64_t *end_blocks,
                                             uint8_t *digests, uint8_t *roots, uint32_t *fail) {
    BSX_REQUIRE(ctx, ctx && data_hashes && start_blocks && end_blocks && digests && roots && N >= 1 && N <= 4096);
    BSX_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(data_hashes) & 3) | ((reinterpret_cast<uintptr_t>(digests) |
                      reinterpret_cast<uintptr_t>(roots)) & 15)) == 0);
    if (t == 0) return BSX_OK;
    uint32_t P = pow2_ceil(N), threads;
    size_t smem;
    int rc = tree_launch_cfg(ctx, (const void *)data_commitment_kernel, P, &threads, &smem);
    if (rc) return rc;
    if (threads < P && P <= 512) threads = P;  // one tuple leaf per thread
    data_commitment_kernel<<<t, threads, smem, (cudaStream_t)stream>>>(data_hashes, N, P, start_blocks, end_blocks, digests,
                                                                       roots, fail);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

extern "C" int bsx_data_commitment_batch(bsx_ctx *ctx, const uint8_t *data_hashes, uint32_t N, uint32_t t,
                                         const uint64_t *start_blocks, const uint64_t *end_blocks, uint8_t *digests,
                                         uint8_t *roots, uint32_t *fail) {
    BSX_REQUIRE(ctx, ctx && data_hashes && start_blocks && end_blocks && digests && roots && N >= 1 && N <= 4096);
    if (t == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t P = pow2_ceil(N);
    size_t s_in = 32 * (size_t)N * t, s_dig = 32 * (size_t)(N + P - 1) * t, s_root = 32 * (size_t)t;
    int rc = ws_begin(ctx, ws_size(s_in) + 2 * ws_size(8 * (size_t)t) + ws_size(s_dig) + ws_size(s_root) + ws_size(4 * (size_t)t));
    if (rc) return rc;
    uint8_t *d_in = ws_take<uint8_t>(ctx, s_in);
    uint64_t *d_s = ws_take<uint64_t>(ctx, t), *d_e = ws_take<uint64_t>(ctx, t);
    uint8_t *d_dig = ws_take<uint8_t>(ctx, s_dig), *d_root = ws_take<uint8_t>(ctx, s_root);
    uint32_t *d_fail = ws_take<uint32_t>(ctx, t);
    BSX_CUDA(ctx, cudaMemcpyAsync(d_in, data_hashes, s_in, cudaMemcpyHostToDevice, ctx->stream));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_s, start_blocks, 8 * (size_t)t, cudaMemcpyHostToDevice, ctx->stream));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_e, end_blocks, 8 * (size_t)t, cudaMemcpyHostToDevice, ctx->stream));
    rc = bsx_data_commitment_batch_dev(ctx, ctx->stream, d_in, N, t, d_s, d_e, d_dig, d_root, d_fail);
    if (rc) return rc;
    BSX_CUDA(ctx, cudaMemcpyAsync(digests, d_dig, s_dig, cudaMemcpyDeviceToHost, ctx->stream));
    BSX_CUDA(ctx, cudaMemcpyAsync(roots, d_root, s_root, cudaMemcpyDeviceToHost, ctx->stream));
    if (fail) BSX_CUDA(ctx, cudaMemcpyAsync(fail, d_fail, 4 * (size_t)t, cudaMemcpyDeviceToHost, ctx->stream));
    BSX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BSX_OK;
}

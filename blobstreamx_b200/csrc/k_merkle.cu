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

// Attestation proofs: the Merkle inclusion proof of one data-root tuple in a data commitment, in the
// form BlobstreamX.verifyAttestation consumes (BX/contracts/src/BlobstreamX.sol, BinaryMerkleProof: sideNodes leaf side
// first, key = height - start, numLeaves = end - start) -- the aunts of compute_hash_from_aunts
// (TX/input/tendermint_utils.rs:225-273).  A pure gather over the digests data_commitment_kernel already wrote:
//   sel_s[j], the value the fixed-shape tree moves up for node j of level s, is the raw inner hash of the first
//   node on j's leftmost path whose two children are both enabled, or the leaf digest if there is none;
//   level s contributes the aunt sel_s[(i >> s) ^ 1] unless that sibling subtree is entirely beyond the last leaf
//   (then the variable-shape Tendermint tree has no node at that level: the value passes up unchanged).
// One thread per (query, level).
__global__ void __launch_bounds__(128) attestation_proofs_kernel(const uint8_t *__restrict__ digests, uint32_t N, uint32_t P, uint32_t logP,
                                                                 const uint64_t *__restrict__ start_blocks,
                                                                 const uint64_t *__restrict__ end_blocks, uint32_t n_q,
                                                                 const uint32_t *__restrict__ q_tree, const uint64_t *__restrict__ q_height,
                                                                 uint8_t *__restrict__ side_nodes, uint32_t *__restrict__ depth,
                                                                 uint32_t *__restrict__ key, uint32_t *__restrict__ num_leaves) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t q = t / (logP ? logP : 1), s = t % (logP ? logP : 1);
    if (q >= n_q) return;
    const uint32_t tr = q_tree[q];
    const uint64_t start = start_blocks[tr], end = end_blocks[tr], h = q_height[q];
    const uint64_t span = end > start ? end - start : 0;
    const uint32_t n = span < (uint64_t)N ? (uint32_t)span : N;
    const bool valid = h >= start && h - start < (uint64_t)n;
    const uint32_t i = valid ? (uint32_t)(h - start) : 0;
    if (s == 0) {
        key[q] = valid ? i : 0xFFFFFFFFu;
        num_leaves[q] = n;
    }
    uint8_t *out = side_nodes + (size_t)q * logP * 32;
    // position of level s among the levels that contribute an aunt, and the total
    uint32_t pos = 0, total = 0;
    bool mine = false;
    for (uint32_t l = 0; l < logP; l++) {
        const uint32_t sib = (i >> l) ^ 1u;
        const bool has = valid && (((uint64_t)sib << l) < (uint64_t)n);
        if (l == s) { mine = has; pos = total; }
        total += has ? 1u : 0u;
    }
    if (s == 0) depth[q] = total;
    if (logP == 0) return;
    uint4 *o = reinterpret_cast<uint4 *>(out + 32 * (size_t)(mine ? pos : 0));
    if (!mine) {
        // zero the unused tail slots: slot (total + k) for the k-th level without an aunt
        uint32_t k = 0;
        for (uint32_t l = 0; l < s; l++) {
            const uint32_t sib = (i >> l) ^ 1u;
            k += (valid && (((uint64_t)sib << l) < (uint64_t)n)) ? 0u : 1u;
        }
        o = reinterpret_cast<uint4 *>(out + 32 * (size_t)(total + k));
        o[0] = make_uint4(0, 0, 0, 0);
        o[1] = make_uint4(0, 0, 0, 0);
        return;
    }
    // descend from node j of level s along left children until a node with both children enabled (or a leaf)
    const uint8_t *base = digests + 32 * (size_t)(N + P - 1) * tr;
    uint32_t j = (i >> s) ^ 1u, lvl = s;
    while (lvl > 0 && !((((uint64_t)(2 * j + 1)) << (lvl - 1)) < (uint64_t)n)) { j <<= 1; lvl--; }
    // layer-major offsets: inner hashes of level 1 (P/2 nodes) start at N, level l at N + P - (P >> (l - 1))
    const size_t idx = lvl == 0 ? (size_t)j : (size_t)N + (size_t)P - (size_t)(P >> (lvl - 1)) + j;
    const uint4 *src = reinterpret_cast<const uint4 *>(base + 32 * idx);
    o[0] = src[0];
    o[1] = src[1];
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

// ---- attestation proofs ----
extern "C" int bsx_attestation_proofs_dev(bsx_ctx *ctx, void *stream, const uint8_t *digests, uint32_t N,
                                          const uint64_t *start_blocks, const uint64_t *end_blocks, uint32_t n_q,
                                          const uint32_t *q_tree, const uint64_t *q_height, uint8_t *side_nodes, uint32_t *depth,
                                          uint32_t *key, uint32_t *num_leaves) {
    BSX_REQUIRE(ctx, ctx && digests && start_blocks && end_blocks && q_tree && q_height && side_nodes && depth && key && num_leaves);
    BSX_REQUIRE(ctx, N >= 1 && N <= 4096);
    BSX_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(digests) | reinterpret_cast<uintptr_t>(side_nodes)) & 15) == 0);
    if (n_q == 0) return BSX_OK;
    const uint32_t P = pow2_ceil(N);
    uint32_t logP = 0;
    while ((1u << logP) < P) logP++;
    const size_t threads = (size_t)n_q * (logP ? logP : 1);
    attestation_proofs_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        digests, N, P, logP, start_blocks, end_blocks, n_q, q_tree, q_height, side_nodes, depth, key, num_leaves);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

extern "C" uint32_t bsx_attestation_max_depth(uint32_t N) {
    uint32_t logP = 0;
    while ((1u << logP) < N) logP++;
    return logP;
}

extern "C" int bsx_attestation_proofs(bsx_ctx *ctx, const uint8_t *data_hashes, uint32_t N, uint32_t t, const uint64_t *start_blocks,
                                      const uint64_t *end_blocks, uint32_t n_q, const uint32_t *q_tree, const uint64_t *q_height,
                                      uint8_t *side_nodes, uint32_t *depth, uint32_t *key, uint32_t *num_leaves, uint8_t *roots) {
    BSX_REQUIRE(ctx, ctx && data_hashes && start_blocks && end_blocks && q_tree && q_height && side_nodes && depth && key && num_leaves);
    BSX_REQUIRE(ctx, N >= 1 && N <= 4096 && t >= 1);
    for (uint32_t q = 0; q < n_q; q++) BSX_REQUIRE(ctx, q_tree[q] < t);
    if (n_q == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t P = pow2_ceil(N), logP = bsx_attestation_max_depth(N);
    const size_t T = t, Q = n_q, s_in = T * N * 32, s_dig = T * (N + P - 1) * 32, s_side = Q * (logP ? logP : 1) * 32;
    int rc = ws_begin(ctx, ws_size(s_in) + 2 * ws_size(8 * T) + ws_size(s_dig) + ws_size(32 * T) + ws_size(4 * T) + ws_size(4 * Q) + ws_size(8 * Q) +
                               ws_size(s_side) + 3 * ws_size(4 * Q));
    if (rc) return rc;
    uint8_t *d_in = ws_take<uint8_t>(ctx, s_in);
    uint64_t *d_s = ws_take<uint64_t>(ctx, T), *d_e = ws_take<uint64_t>(ctx, T);
    uint8_t *d_dig = ws_take<uint8_t>(ctx, s_dig), *d_root = ws_take<uint8_t>(ctx, 32 * T);
    uint32_t *d_fail = ws_take<uint32_t>(ctx, T), *d_qt = ws_take<uint32_t>(ctx, Q);
    uint64_t *d_qh = ws_take<uint64_t>(ctx, Q);
    uint8_t *d_side = ws_take<uint8_t>(ctx, s_side);
    uint32_t *d_depth = ws_take<uint32_t>(ctx, Q), *d_key = ws_take<uint32_t>(ctx, Q), *d_nl = ws_take<uint32_t>(ctx, Q);
    cudaStream_t st = ctx->stream;
    BSX_CUDA(ctx, cudaMemcpyAsync(d_in, data_hashes, s_in, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_s, start_blocks, 8 * T, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_e, end_blocks, 8 * T, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_qt, q_tree, 4 * Q, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_qh, q_height, 8 * Q, cudaMemcpyHostToDevice, st));
    rc = bsx_data_commitment_batch_dev(ctx, st, d_in, N, t, d_s, d_e, d_dig, d_root, d_fail);
    if (rc) return rc;
    rc = bsx_attestation_proofs_dev(ctx, st, d_dig, N, d_s, d_e, n_q, d_qt, d_qh, d_side, d_depth, d_key, d_nl);
    if (rc) return rc;
    BSX_CUDA(ctx, cudaMemcpyAsync(side_nodes, d_side, s_side, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(depth, d_depth, 4 * Q, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(key, d_key, 4 * Q, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(num_leaves, d_nl, 4 * Q, cudaMemcpyDeviceToHost, st));
    if (roots) BSX_CUDA(ctx, cudaMemcpyAsync(roots, d_root, 32 * T, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaStreamSynchronize(st));
    return BSX_OK;
}

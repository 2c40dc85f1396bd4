// verify_header / verify_skip / verify_step (+ prove_next_header_data_commitment): the SHA-256 request
// schedule and assertion logic of the tendermintx light-client gadgets, one CTA per instance.
//   verify_header               TX/builder/verify.rs:225-329
//   verify_trusted_validators   TX/builder/verify.rs:356-433      verify_skip  :527-564
//   verify_step                 TX/builder/verify.rs:468-505
//   hash_validator_set          TX/builder/validator.rs:209-252 (marshal :185-207, varint TX/builder/shared.rs:67-156)
//   prove_next_header_data_commitment   BX/circuits/builder.rs:411-443
// The Ed25519 records of the instance's validators are produced by ed25519_batch_kernel (k_ed25519.cu),
// launched on the same stream just before; this kernel only reads their flag words.
// Digest order = Curta request order.
#include "common.cuh"
// the verify_* schedules run beside the Ed25519 kernel, which keeps the FMA pipe busy: all additions stay on the ALU pipe (r02j)
#ifndef BSX_SHA_FMA_ADDS
#define BSX_SHA_FMA_ADDS 0
#endif
#include "sha256.cuh"
#include "tm_tree.cuh"

namespace bsx {

enum { MODE_HEADER = 0, MODE_SKIP = 1, MODE_STEP = 2 };
static_assert(sizeof(bsx_header_in) == 616 && sizeof(bsx_skip_in) == 224 && sizeof(bsx_step_in) == 576, "bsx.h struct layout");

struct VerifyArgs {
    uint32_t N, P;
    const bsx_header_in *hdr;
    const uint8_t *validators;   // n * N * BSX_VAL_IN_BYTES
    const bsx_skip_in *skip;
    const uint8_t *trusted_pubkeys;
    const uint64_t *trusted_powers;
    const uint32_t *trusted_byte_lengths;
    const bsx_step_in *step;
    uint8_t *digests;            // per instance: see bsx_*_digest_count
    const uint8_t *ed_out;       // n * N * BSX_SIG_OUT_BYTES
    uint8_t *data_commitments;   // MODE_STEP
    uint32_t *fail;
};

__device__ __forceinline__ uint64_t ld_u64(const uint8_t *p) {
    uint64_t v = 0;
#pragma unroll
    for (int i = 7; i >= 0; i--) v = (v << 8) | p[i];
    return v;
}
__device__ __forceinline__ uint32_t ld_u32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ void ld_words_be(const uint8_t *p, uint32_t d[8]) {
#pragma unroll
    for (int k = 0; k < 8; k++)
        d[k] = ((uint32_t)p[4 * k] << 24) | ((uint32_t)p[4 * k + 1] << 16) | ((uint32_t)p[4 * k + 2] << 8) | p[4 * k + 3];
}
__device__ __forceinline__ bool bytes_eq(const uint8_t *a, const uint8_t *b, uint32_t n) {
    uint32_t x = 0;
    for (uint32_t i = 0; i < n; i++) x |= a[i] ^ b[i];
    return x == 0;
}

// marshal_int64_varint (shared.rs:67-156): 9 septets, continuation bit below the last non-zero one
__device__ __forceinline__ void marshal_varint9(uint64_t v, uint8_t out[9]) {
    uint32_t last = 0;
#pragma unroll
    for (int i = 0; i < 9; i++)
        if ((v >> (7 * i)) & 0x7f) last = i;
#pragma unroll
    for (int i = 0; i < 9; i++) out[i] = (uint8_t)(((v >> (7 * i)) & 0x7f) | ((uint32_t)i < last ? 0x80 : 0));
}

// validator leaf: sha256 of the first 1+byte_length bytes of  0x00 ‖ 0a 22 0a 20 ‖ pk ‖ 10 ‖ varint9 ‖ 0...
__device__ __forceinline__ void validator_leaf_hash(const uint8_t *pk, uint64_t power, uint32_t byte_length, uint32_t out[8]) {
    uint8_t buf[64];
#pragma unroll
    for (int i = 0; i < 64; i++) buf[i] = 0;
    buf[1] = 10; buf[2] = 34; buf[3] = 10; buf[4] = 32;
    for (int i = 0; i < 32; i++) buf[5 + i] = pk[i];
    buf[37] = 16;
    marshal_varint9(power, buf + 38);
    uint32_t len = 1 + byte_length;
    if (len > 64) len = 64;
    sha256_bytes([&](uint32_t k) -> uint8_t { return buf[k]; }, len, out);
}

// Merkle inclusion proof by one thread (tendermint.rs:62-93); returns the root in h.
__device__ __forceinline__ void proof_chain(uint32_t h[8], const uint8_t *aunts, uint32_t path_bits, uint8_t *out) {
#pragma unroll 1
    for (int l = 0; l < 4; l++) {
        uint32_t au[8], left[8], right[8];
        ld_words_be(aunts + 32 * l, au);
        tm_inner_hash(h, au, left);
        tm_inner_hash(au, h, right);
        store_digest_be(out + 64 * l, left);
        store_digest_be(out + 64 * l + 32, right);
        const bool sel = (path_bits >> l) & 1;
#pragma unroll
        for (int k = 0; k < 8; k++) h[k] = sel ? right[k] : left[k];
    }
}
// leaf ‖ aunts record (leaf_len 34 or 72): 9 digests; returns root
__device__ __forceinline__ void proof_record(const uint8_t *rec, uint32_t leaf_len, uint32_t path_bits, uint8_t *out, uint32_t root[8]) {
    tm_leaf_hash([&](uint32_t k) -> uint8_t { return rec[k]; }, leaf_len, root);
    store_digest_be(out, root);
    proof_chain(root, rec + leaf_len, path_bits, out + 32);
}
__device__ __forceinline__ bool root_is(const uint32_t root[8], const uint8_t *expect) {
    uint32_t e[8];
    ld_words_be(expect, e);
    return digest_eq(root, e);
}

// hash_validator_set<N> by the whole CTA: N leaf digests + P-1 inner to `out`, root to every lane of warp 0
__device__ __forceinline__ void validator_set_cta(uint32_t N, uint32_t P, const uint8_t *pk, uint32_t pk_stride,
                                                  const uint8_t *power, uint32_t power_stride, const uint8_t *blen,
                                                  uint32_t blen_stride, uint64_t nb_enabled, uint32_t *sA, uint32_t *sB,
                                                  uint8_t *out, uint32_t root[8], uint32_t *s_fail, uint32_t msb_bit) {
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
        uint32_t d[8];
        if (i < N) {
            const uint64_t pw = ld_u64(power + (size_t)power_stride * i);
            // marshal_int64_varint asserts that bit 63 of the voting power is zero (TX/builder/shared.rs:77-80)
            if (pw >> 63) atomicOr(s_fail, msb_bit);
            validator_leaf_hash(pk + (size_t)pk_stride * i, pw, ld_u32(blen + (size_t)blen_stride * i), d);
            store_digest_be(out + 32 * (size_t)i, d);
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) d[k] = 0;
        }
#pragma unroll
        for (int k = 0; k < 8; k++) sA[8 * i + k] = d[k];
    }
    __syncthreads();
    tm_tree_cta(sA, sB, P, nb_enabled, out + 32 * (size_t)N, root);
    __syncthreads();
}

// voting threshold (TX/builder/voting.rs:31-79): sum(include) * den > sum(enabled) * num, u64 wrapping like the gadget
__device__ __forceinline__ bool voting_threshold(uint32_t N, const uint8_t *vals, uint64_t nb_enabled, uint64_t num, uint64_t den,
                                                 uint32_t flag_off) {
    uint64_t total = 0, acc = 0;
    for (uint32_t i = 0; i < N; i++) {
        const uint8_t *v = vals + BSX_VAL_IN_BYTES * (size_t)i;
        uint64_t p = ld_u64(v + 224);
        if ((uint64_t)i < nb_enabled) total += p;
        if (v[flag_off]) acc += p;
    }
    return acc * den > total * num;
}

template <int MODE>
__global__ void __launch_bounds__(256) verify_kernel(VerifyArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t N = a.N, P = a.P, tid = threadIdx.x;
    uint32_t *sA = smem, *sB = smem + 8 * (size_t)P;
    __shared__ uint32_t s_fail;
    const size_t inst = blockIdx.x;
    const bsx_header_in *H = a.hdr + inst;
    const uint8_t *vals = a.validators + inst * N * BSX_VAL_IN_BYTES;
    const uint32_t n_hdr = N + P - 1 + 27;
    const uint32_t n_dig = MODE == MODE_SKIP ? 9 + N + P - 1 + n_hdr : (MODE == MODE_STEP ? n_hdr + 28 : n_hdr);
    uint8_t *out = a.digests + inst * (size_t)n_dig * 32;
    if (tid == 0) s_fail = 0;
    __syncthreads();
    uint32_t root[8];

    if (MODE == MODE_SKIP) {
        const bsx_skip_in *S = a.skip + inst;
        // verify_skip_distance (:507-525), trusted validators-hash proof against the trusted header
        if (tid == 0) {
            uint32_t f = 0;
            const uint64_t target = H->height;
            if (!(target > S->trusted_block + 1)) f |= BSX_VFAIL_SKIP_DISTANCE;
            if (!(target <= S->trusted_block + S->skip_max)) f |= BSX_VFAIL_SKIP_DISTANCE;
            uint32_t r[8];
            proof_record(S->trusted_validators_hash_proof, 34, 7, out, r);
            if (!root_is(r, S->trusted_header)) f |= BSX_VFAIL_TRUSTED_PROOF;
            if (f) atomicOr(&s_fail, f);
        }
        out += 9 * 32;
        validator_set_cta(N, P, a.trusted_pubkeys + inst * N * 32, 32, reinterpret_cast<const uint8_t *>(a.trusted_powers + inst * N), 8,
                          reinterpret_cast<const uint8_t *>(a.trusted_byte_lengths + inst * N), 4, S->trusted_nb_enabled, sA, sB,
                          out, root, &s_fail, BSX_VFAIL_TRUSTED_VALHASH);
        out += 32 * (size_t)(N + P - 1);
        if (tid == 0 && !root_is(root, S->trusted_validators_hash_proof + 2)) atomicOr(&s_fail, BSX_VFAIL_TRUSTED_VALHASH);
        // present_on_trusted_header => signed, and the pubkey really is in the trusted set (O(N^2), :381-415)
        for (uint32_t i = tid; i < N; i += blockDim.x) {
            const uint8_t *v = vals + BSX_VAL_IN_BYTES * (size_t)i;
            if (v[237]) {
                uint32_t f = v[236] ? 0u : BSX_VFAIL_TRUSTED_PRESENT;
                bool found = false;
                const uint8_t *tp = a.trusted_pubkeys + inst * N * 32;
                if ((((uintptr_t)v | (uintptr_t)tp) & 3) == 0) {   // word-aligned inputs: 8 word compares per pair instead of 32 byte pairs
                    const uint32_t *vw = reinterpret_cast<const uint32_t *>(v), *tw = reinterpret_cast<const uint32_t *>(tp);
                    uint32_t k8[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) k8[k] = vw[k];
                    for (uint32_t j = 0; j < N; j++) {
                        uint32_t x = 0;
#pragma unroll
                        for (int k = 0; k < 8; k++) x |= k8[k] ^ __ldg(tw + 8 * j + k);
                        found = found || x == 0;
                    }
                } else {
                    for (uint32_t j = 0; j < N; j++) found = found || bytes_eq(v, tp + j * 32, 32);
                }
                if (!found) f |= BSX_VFAIL_TRUSTED_PRESENT;
                if (f) atomicOr(&s_fail, f);
            }
        }
        if (tid == 0 && !voting_threshold(N, vals, H->nb_enabled, 1, 3, 237)) atomicOr(&s_fail, BSX_VFAIL_TRUSTED_THRESHOLD);
    }

    // ---- verify_header ----
    // (1) EdDSA batch (:239-251): the records come from ed25519_batch_kernel, which runs CONCURRENTLY on the ctx's
    //     high-priority stream; ed25519_flags_kernel folds their flags into fail[] after both have finished.
    // (2) validators hash (:253-267)
    validator_set_cta(N, P, vals, BSX_VAL_IN_BYTES, vals + 224, BSX_VAL_IN_BYTES, vals + 232, BSX_VAL_IN_BYTES, H->nb_enabled, sA,
                      sB, out, root, &s_fail, BSX_VFAIL_VALHASH);
    out += 32 * (size_t)(N + P - 1);
    if (tid == 0 && !root_is(root, H->validators_hash_proof + 2)) atomicOr(&s_fail, BSX_VFAIL_VALHASH);
    // (3) validators-hash proof, (6) chain id, (7) height: three independent chains, one thread each
    if (tid < 3) {
        uint32_t f = 0, r[8];
        if (tid == 0) {
            proof_record(H->validators_hash_proof, 34, 7, out, r);
            if (!root_is(r, H->header)) f |= BSX_VFAIL_VALHASH_PROOF;
        } else {
            uint8_t buf[64];
            for (int i = 0; i < 64; i++) buf[i] = 0;
            uint32_t len;
            if (tid == 1) {  // verify.rs:181-223
                for (int i = 0; i < 52; i++) buf[1 + i] = H->chain_id_enc[i];
                len = 1 + H->chain_id_enc_len;
            } else {  // shared.rs:169-207
                buf[1] = 0x08;
                marshal_varint9(H->height, buf + 2);
                len = 1 + H->height_enc_len;
            }
            if (len > 64) len = 64;
            sha256_bytes([&](uint32_t k) -> uint8_t { return buf[k]; }, len, r);
            uint8_t *o = out + 32 * (size_t)(9 + 9 * (tid - 1));
            store_digest_be(o, r);
            proof_chain(r, tid == 1 ? H->chain_id_aunts : H->height_aunts, tid == 1 ? 1u : 2u, o + 32);
            if (!root_is(r, H->header)) f |= (tid == 1 ? BSX_VFAIL_CHAIN_ID : BSX_VFAIL_HEIGHT);
            if (tid == 1) {
                uint32_t el = H->expected_chain_id_len;
                if (el > sizeof H->expected_chain_id) el = sizeof H->expected_chain_id;
                if (!bytes_eq(H->chain_id_enc + 2, H->expected_chain_id, el)) f |= BSX_VFAIL_CHAIN_ID;
            }
        }
        if (f) atomicOr(&s_fail, f);
    }
    out += 27 * 32;
    // marshal_int64_varint(height) asserts bit 63 == 0 (shared.rs:77-80, called from verify_block_height :178);
    // verify_non_negative_round asserts the round's sign bit == 0 (validator.rs:73-78, once per validator)
    if (tid == 0 && ((H->height >> 63) || (H->round >> 63)))
        atomicOr(&s_fail, ((H->height >> 63) ? BSX_VFAIL_HEIGHT : 0u) | ((H->round >> 63) ? BSX_VFAIL_MESSAGE : 0u));
    // (4) 2/3 threshold over `signed` (:279-288)
    if (tid == 32 && !voting_threshold(N, vals, H->nb_enabled, 2, 3, 236)) atomicOr(&s_fail, BSX_VFAIL_THRESHOLD);
    // (5) per-validator message checks (validator.rs:80-183, verify.rs:290-312)
    for (uint32_t i = tid; i < N; i += blockDim.x) {
        const uint8_t *v = vals + BSX_VAL_IN_BYTES * (size_t)i, *msg = v + 96;
        const bool is_signed = v[236] != 0, enabled = (uint64_t)i < H->nb_enabled;
        const uint64_t round = H->round, height = H->height;
        bool ok = bytes_eq(round == 0 ? msg + 16 : msg + 25, H->header, 32) && msg[1] == 8 && msg[2] == 2;
        ok = ok && ld_u64(msg + 4) == height && (round == 0 || ld_u64(msg + 13) == round);
        const bool valid = is_signed && enabled && ok;
        if (is_signed != valid) atomicOr(&s_fail, BSX_VFAIL_MESSAGE);
    }

    if (MODE == MODE_STEP) {
        const bsx_step_in *S = a.step + inst;
        if (tid < 3) {
            uint32_t f = 0, r[8];
            if (tid == 0) {  // verify_prev_header_in_header (:137-154), LAST_BLOCK_ID_INDEX = 4
                proof_record(S->last_block_id_proof, 72, 4, out, r);
                if (!root_is(r, H->header)) f |= BSX_VFAIL_PREV_HEADER;
                if (!bytes_eq(S->last_block_id_proof + 2, S->prev_header, 32)) f |= BSX_VFAIL_PREV_HEADER;
            } else if (tid == 1) {  // verify_prev_header_next_validators_hash (:156-179), index 8
                proof_record(S->prev_next_validators_proof, 34, 8, out + 9 * 32, r);
                if (!root_is(r, S->prev_header)) f |= BSX_VFAIL_NEXT_VALS;
                if (!bytes_eq(H->validators_hash_proof + 2, S->prev_next_validators_proof + 2, 32)) f |= BSX_VFAIL_NEXT_VALS;
            } else {  // prove_next_header_data_commitment (BX/circuits/builder.rs:411-443), DATA_HASH_INDEX = 6
                proof_record(S->data_hash_proof, 34, 6, out + 18 * 32, r);
                if (!root_is(r, S->prev_header)) f |= BSX_VFAIL_DATA_HASH_PROOF;
                // leaf_hash(encode_data_root_tuple(data_hash, prev_block)) = sha256(0x00 ‖ 0^24 ‖ u64be ‖ data_hash)
                uint8_t t[65];
                for (int i = 0; i < 25; i++) t[i] = 0;
                for (int i = 0; i < 8; i++) t[25 + i] = (uint8_t)(S->prev_block >> (56 - 8 * i));
                for (int i = 0; i < 32; i++) t[33 + i] = S->data_hash_proof[2 + i];
                uint32_t d[8];
                sha256_bytes([&](uint32_t k) -> uint8_t { return t[k]; }, 65, d);
                store_digest_be(out + 27 * 32, d);
                store_digest_be(a.data_commitments + 32 * inst, d);
            }
            if (f) atomicOr(&s_fail, f);
        }
    }
    __syncthreads();
    if (tid == 0) a.fail[inst] = s_fail;
}

// fail[inst] |= BSX_VFAIL_SIG unless every lane's record says "s < l, A and R decompressed, sG == R + hA"
__global__ void ed25519_flags_kernel(uint32_t n, uint32_t N, const uint8_t *__restrict__ ed_out, uint32_t *__restrict__ fail) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * N) return;
    if ((ed_out[(size_t)i * BSX_SIG_OUT_BYTES + 520] & 0xf) != 0xf) atomicOr(fail + i / N, BSX_VFAIL_SIG);
}

static int launch_verify(bsx_ctx *ctx, cudaStream_t st, int mode, uint32_t n, VerifyArgs a) {
    uint32_t P = 1;
    while (P < a.N) P <<= 1;
    a.P = P;
    const size_t smem = 4 * (8 * (size_t)P + 4 * (size_t)P + 16);
    const uint32_t threads = P < 64 ? 64 : (P > 256 ? 256 : P);
    if (smem > 48 * 1024) {   // N > 512: above the default dynamic shared-memory limit
        BSX_CUDA(ctx, cudaFuncSetAttribute(mode == MODE_HEADER ? (const void *)verify_kernel<MODE_HEADER>
                                           : mode == MODE_SKIP ? (const void *)verify_kernel<MODE_SKIP>
                                                               : (const void *)verify_kernel<MODE_STEP>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    switch (mode) {
        case MODE_HEADER: BSX_PIN_CARVEOUT(verify_kernel<MODE_HEADER>); verify_kernel<MODE_HEADER><<<n, threads, smem, st>>>(a); break;
        case MODE_SKIP: BSX_PIN_CARVEOUT(verify_kernel<MODE_SKIP>); verify_kernel<MODE_SKIP><<<n, threads, smem, st>>>(a); break;
        default: BSX_PIN_CARVEOUT(verify_kernel<MODE_STEP>); verify_kernel<MODE_STEP><<<n, threads, smem, st>>>(a); break;
    }
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

}  // namespace bsx

using namespace bsx;

int bsx_ed25519_strided_corun(bsx_ctx *ctx, void *stream, uint32_t n, const uint8_t *pks, uint32_t pk_stride, const uint8_t *sigs,
                              uint32_t sig_stride, const uint8_t *msgs, uint32_t msg_stride, uint32_t msg_max, const uint8_t *msg_lens,
                              uint32_t len_stride, const uint8_t *active, uint32_t active_stride, uint8_t *out, int corun);

extern "C" uint32_t bsx_verify_digest_count(int mode, uint32_t N) {
    uint32_t P = 1;
    while (P < N) P <<= 1;
    const uint32_t n_hdr = N + P - 1 + 27;
    return mode == MODE_SKIP ? 9 + N + P - 1 + n_hdr : (mode == MODE_STEP ? n_hdr + 28 : n_hdr);
}

static int verify_dev(bsx_ctx *ctx, void *stream, int mode, uint32_t n, uint32_t N, const bsx_header_in *hdr,
                      const uint8_t *validators, const bsx_skip_in *skip, const uint8_t *trusted_pubkeys,
                      const uint64_t *trusted_powers, const uint32_t *trusted_byte_lengths, const bsx_step_in *step,
                      uint8_t *digests, uint8_t *ed_out, uint8_t *data_commitments, uint32_t *fail, int ed_corun = 0) {
    BSX_REQUIRE(ctx, ctx && hdr && validators && digests && ed_out && fail);
    BSX_REQUIRE(ctx, N >= 1 && N <= 4096);
    BSX_REQUIRE(ctx, mode != MODE_SKIP || (skip && trusted_pubkeys && trusted_powers && trusted_byte_lengths));
    BSX_REQUIRE(ctx, mode != MODE_STEP || (step && data_commitments));
    BSX_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(digests) | reinterpret_cast<uintptr_t>(data_commitments)) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(hdr) & 7) == 0);
    if (n == 0) return BSX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // fork: Ed25519 (FMA pipe, latency-bound) on the high-priority stream, the SHA-256 schedule on the caller's stream
    BSX_CUDA(ctx, cudaEventRecord(ctx->ev_fork2, st));
    BSX_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork2, 0));
    int rc = bsx_verify_launch_ed(ctx, ctx->stream2, n, N, validators, ed_out, ed_corun);
    if (rc) return rc;
    rc = bsx_verify_launch_hash(ctx, stream, mode, n, N, hdr, validators, skip, trusted_pubkeys, trusted_powers,
                                trusted_byte_lengths, step, digests, data_commitments, fail);
    if (rc) return rc;
    BSX_CUDA(ctx, cudaEventRecord(ctx->ev_join2, ctx->stream2));
    BSX_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join2, 0));
    return bsx_verify_launch_flags(ctx, stream, n, N, ed_out, fail);
}

// the three stages, also used by bsx_header_range_dev to interleave them with the map/reduce kernels
int bsx_verify_launch_ed(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, const uint8_t *validators, uint8_t *ed_out, int ed_corun) {
    // curta_eddsa_verify_sigs_conditional over the validators (is_active = signed), verify.rs:239-251
    return bsx_ed25519_strided_corun(ctx, stream, n * N, validators, BSX_VAL_IN_BYTES, validators + 32, BSX_VAL_IN_BYTES,
                                     validators + 96, BSX_VAL_IN_BYTES, 124, validators + 220, BSX_VAL_IN_BYTES,
                                     validators + 236, BSX_VAL_IN_BYTES, ed_out, ed_corun);
}
// verify_skip with the Ed25519 batch in its co-run register budget (the pipelined host path of bsx_header_range)
int bsx_verify_skip_corun_dev(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, const bsx_header_in *hdr, const uint8_t *validators,
                              const bsx_skip_in *skip, const uint8_t *trusted_pubkeys, const uint64_t *trusted_powers,
                              const uint32_t *trusted_byte_lengths, uint8_t *digests, uint8_t *ed_out, uint32_t *fail) {
    return verify_dev(ctx, stream, MODE_SKIP, n, N, hdr, validators, skip, trusted_pubkeys, trusted_powers, trusted_byte_lengths,
                      nullptr, digests, ed_out, nullptr, fail, 1);
}
int bsx_verify_launch_hash(bsx_ctx *ctx, void *stream, int mode, uint32_t n, uint32_t N, const bsx_header_in *hdr,
                           const uint8_t *validators, const bsx_skip_in *skip, const uint8_t *trusted_pubkeys,
                           const uint64_t *trusted_powers, const uint32_t *trusted_byte_lengths, const bsx_step_in *step,
                           uint8_t *digests, uint8_t *data_commitments, uint32_t *fail) {
    VerifyArgs a{N, 0, hdr, validators, skip, trusted_pubkeys, trusted_powers, trusted_byte_lengths, step, digests, nullptr,
                 data_commitments, fail};
    return launch_verify(ctx, (cudaStream_t)stream, mode, n, a);
}
int bsx_verify_launch_flags(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, const uint8_t *ed_out, uint32_t *fail) {
    const uint32_t total = n * N;
    bsx::ed25519_flags_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, N, ed_out, fail);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

extern "C" int bsx_verify_header_dev(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, const bsx_header_in *hdr,
                                     const uint8_t *validators, uint8_t *digests, uint8_t *ed_out, uint32_t *fail) {
    return verify_dev(ctx, stream, MODE_HEADER, n, N, hdr, validators, nullptr, nullptr, nullptr, nullptr, nullptr, digests,
                      ed_out, nullptr, fail);
}
extern "C" int bsx_verify_skip_dev(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, const bsx_header_in *hdr,
                                   const uint8_t *validators, const bsx_skip_in *skip, const uint8_t *trusted_pubkeys,
                                   const uint64_t *trusted_powers, const uint32_t *trusted_byte_lengths, uint8_t *digests,
                                   uint8_t *ed_out, uint32_t *fail) {
    return verify_dev(ctx, stream, MODE_SKIP, n, N, hdr, validators, skip, trusted_pubkeys, trusted_powers,
                      trusted_byte_lengths, nullptr, digests, ed_out, nullptr, fail);
}
extern "C" int bsx_next_header_dev(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, const bsx_header_in *hdr,
                                   const uint8_t *validators, const bsx_step_in *step, uint8_t *digests, uint8_t *ed_out,
                                   uint8_t *data_commitments, uint32_t *fail) {
    return verify_dev(ctx, stream, MODE_STEP, n, N, hdr, validators, nullptr, nullptr, nullptr, nullptr, step, digests,
                      ed_out, data_commitments, fail);
}

// ---- host-buffer entry points ----
static int verify_host(bsx_ctx *ctx, int mode, uint32_t n, uint32_t N, const bsx_header_in *hdr, const uint8_t *validators,
                       const bsx_skip_in *skip, const uint8_t *trusted_pubkeys, const uint64_t *trusted_powers,
                       const uint32_t *trusted_byte_lengths, const bsx_step_in *step, uint8_t *digests, uint8_t *ed_out,
                       uint8_t *data_commitments, uint32_t *fail) {
    BSX_REQUIRE(ctx, ctx && hdr && validators && digests && ed_out && fail);
    BSX_REQUIRE(ctx, mode != MODE_SKIP || (skip && trusted_pubkeys && trusted_powers && trusted_byte_lengths));
    BSX_REQUIRE(ctx, mode != MODE_STEP || (step && data_commitments));
    if (n == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n_ = n, nN = n_ * N;
    const size_t s_hdr = n_ * sizeof(bsx_header_in), s_val = nN * BSX_VAL_IN_BYTES, s_skip = n_ * sizeof(bsx_skip_in),
                 s_step = n_ * sizeof(bsx_step_in), s_dig = n_ * bsx_verify_digest_count(mode, N) * 32,
                 s_ed = nN * BSX_SIG_OUT_BYTES;
    int rc = ws_begin(ctx, ws_size(s_hdr) + ws_size(s_val) + ws_size(s_skip) + ws_size(s_step) + ws_size(32 * nN) +
                               ws_size(8 * nN) + ws_size(4 * nN) + ws_size(s_dig) + ws_size(s_ed) + ws_size(32 * n_) +
                               ws_size(4 * n_));
    if (rc) return rc;
    auto *d_hdr = reinterpret_cast<bsx_header_in *>(ws_take<uint8_t>(ctx, s_hdr));
    uint8_t *d_val = ws_take<uint8_t>(ctx, s_val);
    auto *d_skip = reinterpret_cast<bsx_skip_in *>(ws_take<uint8_t>(ctx, s_skip));
    auto *d_step = reinterpret_cast<bsx_step_in *>(ws_take<uint8_t>(ctx, s_step));
    uint8_t *d_tpk = ws_take<uint8_t>(ctx, 32 * nN);
    uint64_t *d_tpw = ws_take<uint64_t>(ctx, nN);
    uint32_t *d_tbl = ws_take<uint32_t>(ctx, nN);
    uint8_t *d_dig = ws_take<uint8_t>(ctx, s_dig), *d_ed = ws_take<uint8_t>(ctx, s_ed), *d_dc = ws_take<uint8_t>(ctx, 32 * n_);
    uint32_t *d_fail = ws_take<uint32_t>(ctx, n_);
    cudaStream_t st = ctx->stream;
    BSX_CUDA(ctx, cudaMemcpyAsync(d_hdr, hdr, s_hdr, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_val, validators, s_val, cudaMemcpyHostToDevice, st));
    if (mode == MODE_SKIP) {
        BSX_CUDA(ctx, cudaMemcpyAsync(d_skip, skip, s_skip, cudaMemcpyHostToDevice, st));
        BSX_CUDA(ctx, cudaMemcpyAsync(d_tpk, trusted_pubkeys, 32 * nN, cudaMemcpyHostToDevice, st));
        BSX_CUDA(ctx, cudaMemcpyAsync(d_tpw, trusted_powers, 8 * nN, cudaMemcpyHostToDevice, st));
        BSX_CUDA(ctx, cudaMemcpyAsync(d_tbl, trusted_byte_lengths, 4 * nN, cudaMemcpyHostToDevice, st));
    }
    if (mode == MODE_STEP) BSX_CUDA(ctx, cudaMemcpyAsync(d_step, step, s_step, cudaMemcpyHostToDevice, st));
    rc = verify_dev(ctx, st, mode, n, N, d_hdr, d_val, d_skip, d_tpk, d_tpw, d_tbl, d_step, d_dig, d_ed, d_dc, d_fail);
    if (rc) return rc;
    BSX_CUDA(ctx, cudaMemcpyAsync(digests, d_dig, s_dig, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(ed_out, d_ed, s_ed, cudaMemcpyDeviceToHost, st));
    if (mode == MODE_STEP) BSX_CUDA(ctx, cudaMemcpyAsync(data_commitments, d_dc, 32 * n_, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(fail, d_fail, 4 * n_, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaStreamSynchronize(st));
    return BSX_OK;
}

extern "C" int bsx_verify_header(bsx_ctx *ctx, uint32_t n, uint32_t N, const bsx_header_in *hdr, const uint8_t *validators,
                                 uint8_t *digests, uint8_t *ed_out, uint32_t *fail) {
    return verify_host(ctx, MODE_HEADER, n, N, hdr, validators, nullptr, nullptr, nullptr, nullptr, nullptr, digests, ed_out,
                       nullptr, fail);
}
extern "C" int bsx_verify_skip(bsx_ctx *ctx, uint32_t n, uint32_t N, const bsx_header_in *hdr, const uint8_t *validators,
                               const bsx_skip_in *skip, const uint8_t *trusted_pubkeys, const uint64_t *trusted_powers,
                               const uint32_t *trusted_byte_lengths, uint8_t *digests, uint8_t *ed_out, uint32_t *fail) {
    return verify_host(ctx, MODE_SKIP, n, N, hdr, validators, skip, trusted_pubkeys, trusted_powers, trusted_byte_lengths,
                       nullptr, digests, ed_out, nullptr, fail);
}
extern "C" int bsx_next_header(bsx_ctx *ctx, uint32_t n, uint32_t N, const bsx_header_in *hdr, const uint8_t *validators,
                               const bsx_step_in *step, uint8_t *digests, uint8_t *ed_out, uint8_t *data_commitments,
                               uint32_t *fail) {
    return verify_host(ctx, MODE_STEP, n, N, hdr, validators, nullptr, nullptr, nullptr, nullptr, step, digests, ed_out,
                       data_commitments, fail);
}

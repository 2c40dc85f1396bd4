// Input shaping, part 2: the protobuf field encoders of a header, CanonicalVote sign-bytes, validator
// records and the present_on_trusted_header walk -- what the reference's off-chain input fetcher computes per header
// and per validator before the circuits see anything:
//   generate_proofs_from_header         TX/input/tendermint_utils.rs:374-393  (the 14 encode_vec calls)
//   get_signed_message_data             TX/input/conversion.rs:20-57
//   get_validator_data_from_block       TX/input/conversion.rs:59-140
//   validator_hash_field_from_block     TX/input/conversion.rs:142-184
//   update_present_on_trusted_header    TX/input/conversion.rs:186-240
// All byte-serial work on a few hundred bytes per item: one thread per header / validator slot; the header records are
// staged through shared memory on both sides, the (smaller) validator records are built in local memory and stored with
// 16-byte writes; the trusted-set walk is sequential per commit and takes one warp.
#include <cstddef>

#include "common.cuh"

// the layouts the Python binding (blobstreamx_b200/inputs.py) and the Rust FFI crate mirror
static_assert(sizeof(bsx_header_fields) == 464 && offsetof(bsx_header_fields, chain_id) == 40 && offsetof(bsx_header_fields, hash_len) == 101 &&
                  offsetof(bsx_header_fields, last_block_hash) == 112 && offsetof(bsx_header_fields, hashes) == 176,
              "bsx_header_fields layout");
static_assert(sizeof(bsx_commit_in) == 152 && offsetof(bsx_commit_in, block_hash) == 16 && offsetof(bsx_commit_in, chain_id) == 88, "bsx_commit_in layout");
static_assert(sizeof(bsx_commit_sig_in) == 160 && offsetof(bsx_commit_sig_in, voting_power) == 96 && offsetof(bsx_commit_sig_in, address) == 120,
              "bsx_commit_sig_in layout");

namespace bsx {
namespace {

__constant__ uint8_t DUMMY_PK[32] = {138, 136, 227, 221, 116, 9,  241, 149, 253, 82,  219, 45, 60,  186, 93, 114,
                                     202, 103, 9,   191, 29,  148, 18,  27,  243, 116, 136, 1,  180, 15,  111, 92};
__constant__ uint8_t DUMMY_SIGN[64] = {55,  20,  104, 158, 84,  120, 194, 17,  6,   237, 157, 164, 85,  88,  158, 137,
                                       187, 119, 187, 240, 159, 73,  80,  63,  133, 162, 74,  91,  48,  53,  6,   138,
                                       1,   41,  22,  121, 249, 46,  198, 145, 155, 102, 3,   210, 168, 135, 173, 55,
                                       252, 72,  45,  126, 169, 178, 191, 7,   153, 67,  112, 90,  150, 33,  140, 7};

// proto3 writer over a byte buffer: W = 1 a linear buffer (local memory); W > 1 a buffer whose 32-bit words are interleaved
// with those of W - 1 other writers (word w of this writer at word index w * W), the layout of the header kernel's stage
template <int W>
struct PbW {
    uint8_t *p;
    uint32_t n;
    __device__ void byte(uint32_t b) {
        p[W == 1 ? n : (n >> 2) * (4 * W) + (n & 3)] = (uint8_t)b;
        n++;
    }
    __device__ void varint(uint64_t v) {
        while (v >= 0x80) {
            byte((uint32_t)(v & 0x7F) | 0x80);
            v >>= 7;
        }
        byte((uint32_t)v);
    }
    __device__ void raw(const uint8_t *s, uint32_t len) {
        uint32_t i = 0;
        if (W > 1) {
            // interleaved stage, source 4-byte aligned (every byte array of bsx_header_fields is): whole words, shifted into
            // place; the buffer starts zeroed and later byte stores only touch their own byte, so a partial word can be stored as is
            const uint32_t *sw = reinterpret_cast<const uint32_t *>(s);
            uint32_t *dw = reinterpret_cast<uint32_t *>(p);
            const uint32_t sh = 8 * (n & 3), words = len >> 2;
            uint32_t at = (n >> 2) * W, carry = sh ? dw[at] : 0;
            for (uint32_t k = 0; k < words; k++, at += W) {
                const uint32_t w = sw[k];
                dw[at] = carry | (w << sh);
                carry = sh ? w >> (32 - sh) : 0;
            }
            if (sh && words) dw[at] = carry;
            i = 4 * words;
            n += i;
        }
        for (; i < len; i++) byte(s[i]);
    }
    __device__ void fixed64(uint64_t v) {
        for (int i = 0; i < 8; i++) byte((uint32_t)(v >> (8 * i)) & 0xFF);
    }
    // varint field, omitted when zero / bytes field, omitted when empty
    __device__ void vi(uint32_t tag, uint64_t v) {
        if (v) {
            byte(tag);
            varint(v);
        }
    }
    __device__ void ld(uint32_t tag, const uint8_t *s, uint32_t len) {
        if (len) {
            byte(tag);
            varint(len);
            raw(s, len);
        }
    }
};
using Pb = PbW<1>;
__device__ __forceinline__ uint32_t varint_len(uint64_t v) {
    uint32_t n = 1;
    while (v >= 0x80) {
        v >>= 7;
        n++;
    }
    return n;
}
// BlockId / CanonicalBlockId (same field numbers): hash = 1, part_set_header = 2 {total = 1, hash = 2}
template <class O>
__device__ void pb_block_id(O &o, const uint8_t *hash, uint32_t parts_total, const uint8_t *parts_hash) {
    o.ld(0x0A, hash, 32);
    o.byte(0x12);
    o.varint((parts_total ? 1 + varint_len(parts_total) : 0) + 34);
    o.vi(0x08, parts_total);
    o.ld(0x12, parts_hash, 32);
}
__device__ __forceinline__ uint32_t block_id_len(uint32_t parts_total) { return 34 + 2 + (parts_total ? 1 + varint_len(parts_total) : 0) + 34; }
template <class O>
__device__ void pb_timestamp(O &o, int64_t seconds, uint32_t nanos) {
    o.vi(0x08, (uint64_t)seconds);
    o.vi(0x10, nanos);
}
__device__ __forceinline__ uint32_t timestamp_len(int64_t seconds, uint32_t nanos) {
    return (seconds ? 1 + varint_len((uint64_t)seconds) : 0) + (nanos ? 1 + varint_len(nanos) : 0);
}

// header record: lengths of the 14 fields in bytes [0,14), the fields back to back from byte 16 (bsx.h).
// One thread encodes one header, one warp per CTA, and neither side touches global or local memory byte-wise: the 32
// input structs come in as coalesced asynchronous copies into shared memory (stride padded to 117 words: a thread's byte
// reads stay in its own bank), and the record is written byte by byte into a second shared buffer with the threads' words
// interleaved (word w of thread t at w*33 + t: conflict-free while the threads are at the same offset, a few-way conflict
// once their field lengths differ), from which the warp stores each record as four 128-byte rows.
// (First version -- fields read from global, record built in local memory: 32 sectors per byte access, 0.68 ms for 262 k
// headers = 6 % of the copy bandwidth.)
constexpr int ENC_T = 32, ENC_IN_WORDS = sizeof(bsx_header_fields) / 4, ENC_IN_STRIDE = ENC_IN_WORDS + 1, ENC_OUT_WORDS = BSX_HEADER_LEAVES_BYTES / 4;
static_assert(sizeof(bsx_header_fields) % 16 == 0, "bulk loads of the input structs");

__global__ void __launch_bounds__(ENC_T) encode_headers_kernel(uint32_t n, const bsx_header_fields *__restrict__ fields, uint8_t *__restrict__ headers) {
    __shared__ __align__(16) uint32_t sm_in[ENC_T * ENC_IN_STRIDE];
    __shared__ __align__(16) uint32_t sm_out[ENC_OUT_WORDS * (ENC_T + 1)];
    const uint32_t base = blockIdx.x * ENC_T, t = threadIdx.x, cnt = n - base < (uint32_t)ENC_T ? n - base : (uint32_t)ENC_T;
    {
        // asynchronous 4-byte copies (the padded stride rules out 16-byte ones): all 116 of a lane are in flight at once.
        // (A register-staged loop -- load 16 bytes, store 4 words -- had ONE load in flight per warp: 29 DRAM round trips.)
        const uint32_t *src = reinterpret_cast<const uint32_t *>(fields + base);
        for (uint32_t q = t; q < cnt * ENC_IN_WORDS; q += ENC_T) {
            const uint32_t s = q / ENC_IN_WORDS, w = q % ENC_IN_WORDS;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm_in + s * ENC_IN_STRIDE + w);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src + q) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        for (int w = 0; w < ENC_OUT_WORDS; w++) sm_out[w * (ENC_T + 1) + t] = 0;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    if (t < cnt) {
        // field offsets as in bsx_header_fields; the struct sits at a 4-byte aligned shared address
        const uint8_t *hb = reinterpret_cast<const uint8_t *>(sm_in + t * ENC_IN_STRIDE);
        auto u32at = [&](int off) { return *reinterpret_cast<const uint32_t *>(hb + off); };
        auto u64at = [&](int off) { return (uint64_t)u32at(off) | ((uint64_t)u32at(off + 4) << 32); };
        const uint64_t version_block = u64at(offsetof(bsx_header_fields, version_block)), version_app = u64at(offsetof(bsx_header_fields, version_app)),
                       height = u64at(offsetof(bsx_header_fields, height));
        const int64_t time_seconds = (int64_t)u64at(offsetof(bsx_header_fields, time_seconds));
        const uint32_t time_nanos = u32at(offsetof(bsx_header_fields, time_nanos)), chain_id_len = u32at(offsetof(bsx_header_fields, chain_id_len)),
                       parts_total = u32at(offsetof(bsx_header_fields, parts_total));
        const uint8_t *hash_len = hb + offsetof(bsx_header_fields, hash_len);
        PbW<ENC_T + 1> o{reinterpret_cast<uint8_t *>(sm_out + t), 16};
        uint32_t at = o.n;
        auto close = [&](uint32_t f) {
            o.p[(f >> 2) * (4 * (ENC_T + 1)) + (f & 3)] = (uint8_t)(o.n - at);
            at = o.n;
        };
        o.vi(0x08, version_block);
        o.vi(0x10, version_app);
        close(0);
        o.ld(0x0A, hb + offsetof(bsx_header_fields, chain_id), chain_id_len > 56 ? 56 : chain_id_len);
        close(1);
        o.vi(0x08, height);
        close(2);
        pb_timestamp(o, time_seconds, time_nanos);
        close(3);
        if (hb[offsetof(bsx_header_fields, has_last_block_id)])
            pb_block_id(o, hb + offsetof(bsx_header_fields, last_block_hash), parts_total, hb + offsetof(bsx_header_fields, parts_hash));
        close(4);
        for (int k = 0; k < 9; k++) {
            o.ld(0x0A, hb + offsetof(bsx_header_fields, hashes) + 32 * k, hash_len[k] > 32 ? 32 : hash_len[k]);
            close(5 + k);
        }
    }
    __syncwarp();
    uint32_t *dst = reinterpret_cast<uint32_t *>(headers) + (size_t)base * ENC_OUT_WORDS;
    for (uint32_t j = 0; j < cnt; j++)
        for (int k = 0; k < ENC_OUT_WORDS / 32; k++) dst[(size_t)j * ENC_OUT_WORDS + 32 * k + t] = sm_out[(32 * k + t) * (ENC_T + 1) + j];
}

// one thread per validator slot
__global__ void __launch_bounds__(128) validator_records_kernel(uint32_t n, uint32_t N, const bsx_commit_in *__restrict__ commits,
                                                                 const bsx_commit_sig_in *__restrict__ sigs, uint8_t *__restrict__ validators,
                                                                 uint8_t *__restrict__ pubkeys, uint64_t *__restrict__ powers,
                                                                 uint32_t *__restrict__ byte_lengths, uint32_t *__restrict__ fail) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n * N) return;
    const uint32_t c = (uint32_t)(t / N), i = (uint32_t)(t % N);
    const bsx_commit_in &cm = commits[c];
    const bsx_commit_sig_in &sg = sigs[t];
    if (i == 0 && cm.n_signatures > N) atomicOr(&fail[c], BSX_FAIL_INPUT_SIGN_BYTES);
    const bool in_set = i < cm.n_signatures, is_signed = in_set && sg.block_id_flag == 2;
    const uint8_t *pk = in_set ? sg.pubkey : DUMMY_PK;
    const uint64_t power = in_set ? sg.voting_power : 0;
    const uint32_t vlen = in_set ? 37 + varint_len(power) : 46;      // 0a 22 0a 20 pk 10 varint(power); VALIDATOR_BYTE_LENGTH_MAX
    if (pubkeys) {
        for (int k = 0; k < 32; k++) pubkeys[32 * t + k] = pk[k];
        powers[t] = power;
        byte_lengths[t] = vlen;
    }
    if (!validators) return;
    __align__(16) uint8_t rec[BSX_VAL_IN_BYTES];
    for (int k = 0; k < BSX_VAL_IN_BYTES / 4; k++) reinterpret_cast<uint32_t *>(rec)[k] = 0;
    for (int k = 0; k < 32; k++) rec[k] = pk[k];
    const uint8_t *sig = is_signed ? sg.signature : DUMMY_SIGN;
    for (int k = 0; k < 64; k++) rec[32 + k] = sig[k];
    uint32_t msg_len = 32;
    if (is_signed) {
        // CanonicalVote: type = 1 (precommit = 2), height = 2 sfixed64, round = 3 sfixed64, block_id = 4, timestamp = 5,
        // chain_id = 6; SignedVote::sign_bytes is the length-delimited encoding
        const uint32_t cid = cm.chain_id_len > 56 ? 56 : cm.chain_id_len, bid = cm.has_block_id ? block_id_len(cm.parts_total) : 0,
                       ts = timestamp_len(sg.ts_seconds, sg.ts_nanos);
        const uint32_t body = 2 + (cm.height ? 9 : 0) + (cm.round ? 9 : 0) + (bid ? 1 + varint_len(bid) + bid : 0) + 2 + ts + (cid ? 2 + cid : 0);
        msg_len = varint_len(body) + body;
        if (msg_len > 124) {
            atomicOr(&fail[c], BSX_FAIL_INPUT_SIGN_BYTES);
            msg_len = 0;
        } else {
            Pb o{rec + 96, 0};
            o.varint(body);
            o.byte(0x08);
            o.byte(0x02);
            if (cm.height) {
                o.byte(0x11);
                o.fixed64(cm.height);
            }
            if (cm.round) {
                o.byte(0x19);
                o.fixed64(cm.round);
            }
            if (bid) {
                o.byte(0x22);
                o.varint(bid);
                pb_block_id(o, cm.block_hash, cm.parts_total, cm.parts_hash);
            }
            o.byte(0x2A);
            o.varint(ts);
            pb_timestamp(o, sg.ts_seconds, sg.ts_nanos);
            o.ld(0x32, cm.chain_id, cid);
        }
    }
    reinterpret_cast<uint32_t *>(rec)[220 / 4] = msg_len;
    reinterpret_cast<uint32_t *>(rec)[224 / 4] = (uint32_t)power;
    reinterpret_cast<uint32_t *>(rec)[228 / 4] = (uint32_t)(power >> 32);
    reinterpret_cast<uint32_t *>(rec)[232 / 4] = vlen;
    rec[236] = is_signed ? 1 : 0;
    uint4 *dst = reinterpret_cast<uint4 *>(validators + (size_t)BSX_VAL_IN_BYTES * t);
    for (int k = 0; k < BSX_VAL_IN_BYTES / 16; k++) dst[k] = reinterpret_cast<const uint4 *>(rec)[k];
}

__device__ __forceinline__ bool addr_eq(const uint8_t *a, const uint8_t *b) {
    bool eq = true;
    for (int k = 0; k < 20; k++) eq &= a[k] == b[k];
    return eq;
}

// one warp per commit: lanes search the target set / the commit's signatures, the walk over the trusted set is serial
__global__ void __launch_bounds__(128) present_on_trusted_kernel(uint32_t n, uint32_t N, const bsx_commit_sig_in *__restrict__ target,
                                                                  const uint32_t *__restrict__ n_target, const bsx_commit_sig_in *__restrict__ trusted,
                                                                  const uint32_t *__restrict__ n_trusted, uint8_t *__restrict__ validators,
                                                                  uint32_t *__restrict__ fail) {
    const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x & 31;
    if (c >= n) return;
    const bsx_commit_sig_in *tg = target + (size_t)c * N, *tr = trusted + (size_t)c * N;
    const uint32_t nt = n_target[c] < N ? n_target[c] : N, ns = n_trusted[c] < N ? n_trusted[c] : N;
    unsigned long long total = 0;
    for (uint32_t i = lane; i < nt; i += 32) total += tg[i].voting_power;
    for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    const double third = __dmul_rn((double)total, 1.0 / 3.0);
    unsigned long long shared = 0;
    for (uint32_t s = 0; s < ns && third > (double)shared; s++) {
        const uint8_t *addr = tr[s].address;
        uint32_t idx = 0xFFFFFFFFu;
        for (uint32_t i = lane; i < nt && idx == 0xFFFFFFFFu; i += 32)
            if (addr_eq(tg[i].address, addr)) idx = i;
        idx = __reduce_min_sync(0xffffffffu, idx);
        if (idx == 0xFFFFFFFFu) continue;
        uint32_t cnt = 0;
        for (uint32_t j = lane; j < nt; j += 32)
            cnt += (tg[j].block_id_flag == 2 || tg[j].block_id_flag == 3) && addr_eq(tg[j].sig_address, addr);
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (cnt) {
            shared += (unsigned long long)cnt * tg[idx].voting_power;
            if (lane == 0) validators[((size_t)c * N + idx) * BSX_VAL_IN_BYTES + 237] = 1;
        }
    }
    if (lane == 0 && third > (double)shared) atomicOr(&fail[c], BSX_FAIL_INPUT_THRESHOLD);
}

}  // namespace
}  // namespace bsx

extern "C" int bsx_encode_headers_dev(bsx_ctx *ctx, void *stream, uint32_t n, const bsx_header_fields *fields, uint8_t *headers) {
    BSX_REQUIRE(ctx, ctx && fields && headers);
    BSX_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(fields) & 15) == 0 && (reinterpret_cast<uintptr_t>(headers) & 15) == 0);
    if (n == 0) return BSX_OK;
    // not part of the witness pipeline (never beside the Ed25519 wave): takes the largest shared-memory carveout, 7 CTAs per SM
    static const bool carve = (cudaFuncSetAttribute(bsx::encode_headers_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100), true);
    (void)carve;
    bsx::encode_headers_kernel<<<(n + bsx::ENC_T - 1) / bsx::ENC_T, bsx::ENC_T, 0, (cudaStream_t)stream>>>(n, fields, headers);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

extern "C" int bsx_validator_records_dev(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, const bsx_commit_in *commits,
                                         const bsx_commit_sig_in *sigs, uint8_t *validators, uint8_t *pubkeys, uint64_t *powers,
                                         uint32_t *byte_lengths, uint32_t *fail) {
    BSX_REQUIRE(ctx, ctx && commits && sigs && fail && N >= 1);
    BSX_REQUIRE(ctx, validators || pubkeys);
    BSX_REQUIRE(ctx, (pubkeys != nullptr) == (powers != nullptr) && (pubkeys != nullptr) == (byte_lengths != nullptr));
    BSX_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(commits) | reinterpret_cast<uintptr_t>(sigs)) & 7) == 0 &&
                         (reinterpret_cast<uintptr_t>(validators) & 15) == 0);
    if (n == 0) return BSX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    BSX_CUDA(ctx, cudaMemsetAsync(fail, 0, 4 * (size_t)n, st));
    const size_t slots = (size_t)n * N;
    BSX_PIN_CARVEOUT(bsx::validator_records_kernel);
    bsx::validator_records_kernel<<<(unsigned)((slots + 127) / 128), 128, 0, st>>>(n, N, commits, sigs, validators, pubkeys, powers, byte_lengths, fail);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

extern "C" int bsx_present_on_trusted_dev(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, const bsx_commit_sig_in *target_sigs,
                                          const uint32_t *n_target, const bsx_commit_sig_in *trusted_sigs, const uint32_t *n_trusted,
                                          uint8_t *validators, uint32_t *fail) {
    BSX_REQUIRE(ctx, ctx && target_sigs && n_target && trusted_sigs && n_trusted && validators && fail && N >= 1);
    if (n == 0) return BSX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    BSX_CUDA(ctx, cudaMemsetAsync(fail, 0, 4 * (size_t)n, st));
    BSX_PIN_CARVEOUT(bsx::present_on_trusted_kernel);
    bsx::present_on_trusted_kernel<<<(n + 3) / 4, 128, 0, st>>>(n, N, target_sigs, n_target, trusted_sigs, n_trusted, validators, fail);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

// ---- host-buffer forms ----
using bsx::ws_begin;
using bsx::ws_size;
using bsx::ws_take;

extern "C" int bsx_encode_headers(bsx_ctx *ctx, uint32_t n, const bsx_header_fields *fields, uint8_t *headers) {
    BSX_REQUIRE(ctx, ctx && fields && headers);
    if (n == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t s_in = sizeof(bsx_header_fields) * (size_t)n, s_out = (size_t)BSX_HEADER_LEAVES_BYTES * n;
    int rc = ws_begin(ctx, ws_size(s_in) + ws_size(s_out));
    if (rc) return rc;
    auto *d_in = ws_take<bsx_header_fields>(ctx, n);
    uint8_t *d_out = ws_take<uint8_t>(ctx, s_out);
    cudaStream_t st = ctx->stream;
    BSX_CUDA(ctx, cudaMemcpyAsync(d_in, fields, s_in, cudaMemcpyHostToDevice, st));
    rc = bsx_encode_headers_dev(ctx, st, n, d_in, d_out);
    if (rc) return rc;
    BSX_CUDA(ctx, cudaMemcpyAsync(headers, d_out, s_out, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaStreamSynchronize(st));
    return BSX_OK;
}

extern "C" int bsx_validator_records(bsx_ctx *ctx, uint32_t n, uint32_t N, const bsx_commit_in *commits, const bsx_commit_sig_in *sigs,
                                     uint8_t *validators, uint8_t *pubkeys, uint64_t *powers, uint32_t *byte_lengths, uint32_t *fail) {
    BSX_REQUIRE(ctx, ctx && commits && sigs && fail && N >= 1);
    BSX_REQUIRE(ctx, validators || pubkeys);
    BSX_REQUIRE(ctx, (pubkeys != nullptr) == (powers != nullptr) && (pubkeys != nullptr) == (byte_lengths != nullptr));
    if (n == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t slots = (size_t)n * N;
    int rc = ws_begin(ctx, ws_size(sizeof(bsx_commit_in) * (size_t)n) + ws_size(sizeof(bsx_commit_sig_in) * slots) +
                               ws_size(slots * BSX_VAL_IN_BYTES) + ws_size(slots * 32) + ws_size(slots * 8) + ws_size(slots * 4) + ws_size(4 * (size_t)n));
    if (rc) return rc;
    auto *d_cm = ws_take<bsx_commit_in>(ctx, n);
    auto *d_sg = ws_take<bsx_commit_sig_in>(ctx, slots);
    uint8_t *d_val = ws_take<uint8_t>(ctx, slots * BSX_VAL_IN_BYTES), *d_pk = ws_take<uint8_t>(ctx, slots * 32);
    uint64_t *d_pw = ws_take<uint64_t>(ctx, slots);
    uint32_t *d_bl = ws_take<uint32_t>(ctx, slots), *d_fail = ws_take<uint32_t>(ctx, n);
    cudaStream_t st = ctx->stream;
    BSX_CUDA(ctx, cudaMemcpyAsync(d_cm, commits, sizeof(bsx_commit_in) * (size_t)n, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_sg, sigs, sizeof(bsx_commit_sig_in) * slots, cudaMemcpyHostToDevice, st));
    rc = bsx_validator_records_dev(ctx, st, n, N, d_cm, d_sg, validators ? d_val : nullptr, pubkeys ? d_pk : nullptr, pubkeys ? d_pw : nullptr,
                                   pubkeys ? d_bl : nullptr, d_fail);
    if (rc) return rc;
    if (validators) BSX_CUDA(ctx, cudaMemcpyAsync(validators, d_val, slots * BSX_VAL_IN_BYTES, cudaMemcpyDeviceToHost, st));
    if (pubkeys) {
        BSX_CUDA(ctx, cudaMemcpyAsync(pubkeys, d_pk, slots * 32, cudaMemcpyDeviceToHost, st));
        BSX_CUDA(ctx, cudaMemcpyAsync(powers, d_pw, slots * 8, cudaMemcpyDeviceToHost, st));
        BSX_CUDA(ctx, cudaMemcpyAsync(byte_lengths, d_bl, slots * 4, cudaMemcpyDeviceToHost, st));
    }
    BSX_CUDA(ctx, cudaMemcpyAsync(fail, d_fail, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaStreamSynchronize(st));
    return BSX_OK;
}

extern "C" int bsx_present_on_trusted(bsx_ctx *ctx, uint32_t n, uint32_t N, const bsx_commit_sig_in *target_sigs, const uint32_t *n_target,
                                      const bsx_commit_sig_in *trusted_sigs, const uint32_t *n_trusted, uint8_t *validators, uint32_t *fail) {
    BSX_REQUIRE(ctx, ctx && target_sigs && n_target && trusted_sigs && n_trusted && validators && fail && N >= 1);
    if (n == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t slots = (size_t)n * N, s_sg = sizeof(bsx_commit_sig_in) * slots;
    int rc = ws_begin(ctx, 2 * ws_size(s_sg) + 3 * ws_size(4 * (size_t)n) + ws_size(slots * BSX_VAL_IN_BYTES));
    if (rc) return rc;
    auto *d_tg = ws_take<bsx_commit_sig_in>(ctx, slots), *d_tr = ws_take<bsx_commit_sig_in>(ctx, slots);
    uint32_t *d_nt = ws_take<uint32_t>(ctx, n), *d_ns = ws_take<uint32_t>(ctx, n), *d_fail = ws_take<uint32_t>(ctx, n);
    uint8_t *d_val = ws_take<uint8_t>(ctx, slots * BSX_VAL_IN_BYTES);
    cudaStream_t st = ctx->stream;
    BSX_CUDA(ctx, cudaMemcpyAsync(d_tg, target_sigs, s_sg, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_tr, trusted_sigs, s_sg, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_nt, n_target, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_ns, n_trusted, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_val, validators, slots * BSX_VAL_IN_BYTES, cudaMemcpyHostToDevice, st));
    rc = bsx_present_on_trusted_dev(ctx, st, n, N, d_tg, d_nt, d_tr, d_ns, d_val, d_fail);
    if (rc) return rc;
    BSX_CUDA(ctx, cudaMemcpyAsync(validators, d_val, slots * BSX_VAL_IN_BYTES, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(fail, d_fail, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaStreamSynchronize(st));
    return BSX_OK;
}

// SHA-256 execution trace on the device: the per-round table a STARK over the Curta SHA-256 accelerator
// commits to, written column-major (one polynomial per column) from the HashInputData this library already produces.
// Reference: `HashStark::prove` fills its trace row by row on one CPU thread -- TraceWriter::write_row_instructions over
// num_rows = 64 x chunks (PX/frontend/hash/curta/stark.rs:107-133; widths NUM_FREE_COLUMNS = 418 / EXTENDED_COLUMNS = 912
// at PX/frontend/hash/sha/sha256/curta.rs:32-33).  The column assignment itself lives in starkyx@ad8eb4ba (UN-VENDORED:
// SHABuilder / UintInstruction register allocation), so this layout is OURS and PARITY IS UNPINNED: it carries the values
// a byte-oriented SHA-256 AIR constrains -- every 32-bit word as 4 little-endian byte limbs like starkyx's U32Register,
// every rotation / xor / and / modular sum of a round and of the schedule step as its own word, the carries, the row
// flags -- 176 of the reference's 418 free columns (the rest are lookup multiplicities and bus bookkeeping that depend on
// the un-vendored AIR; the 912 extended columns are challenge-dependent accumulators written after the first commitment).
// Sizing: header_range_1024 map job = 1246 chunks -> 79 744 rows -> 2^17 rows x 176 columns x 8 B = 185 MB;
// at the reference's 418 + 912 columns it is 1.4 GB per map circuit -- an HBM write stream.
//
// One warp per chunk: lane 0 runs the 64 rounds once and leaves the schedule and the 65 working states in shared memory;
// then lane l expands rows l and l + 32, so that for every column a warp stores 32 consecutive rows (256 bytes).
#include "common.cuh"
#include "sha256.cuh"
#include "sha512.cuh"

namespace bsx {

#define TR_COLS BSX_SHA256_TRACE_COLS


struct TraceRow {
    uint64_t *base;      // column 0 of this row
    size_t stride;       // rows per column
    __device__ __forceinline__ void byte4(int col, uint32_t v) const {
#pragma unroll
        for (int k = 0; k < 4; k++) __stcs(base + (size_t)(col + k) * stride, (uint64_t)((v >> (8 * k)) & 0xff));
    }
    __device__ __forceinline__ void put(int col, uint64_t v) const { __stcs(base + (size_t)col * stride, v); }
};

__constant__ uint32_t TR_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
__constant__ uint32_t TR_IV[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};

// one round t on state s (a..h), schedule word w: returns the new state in s
__device__ __forceinline__ void tr_round(uint32_t s[8], uint32_t w, uint32_t k) {
    const uint32_t S1 = rotr32(s[4], 6) ^ rotr32(s[4], 11) ^ rotr32(s[4], 25), ch = (s[4] & s[5]) ^ (~s[4] & s[6]);
    const uint32_t t1 = s[7] + S1 + ch + k + w;
    const uint32_t S0 = rotr32(s[0], 2) ^ rotr32(s[0], 13) ^ rotr32(s[0], 22), mj = (s[0] & s[1]) ^ (s[0] & s[2]) ^ (s[1] & s[2]);
    const uint32_t t2 = S0 + mj;
    s[7] = s[6]; s[6] = s[5]; s[5] = s[4]; s[4] = s[3] + t1; s[3] = s[2]; s[2] = s[1]; s[1] = s[0]; s[0] = t1 + t2;
}

// columns of one row (see include/bsx.h BSX_SHA256_TRACE_COLS for the table)
__device__ __forceinline__ void tr_expand_row(const TraceRow &o, int t, const uint32_t *W, const uint32_t *st, bool end_bit, bool digest_bit) {
    const uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7], w = W[t];
    o.byte4(0, w);
#pragma unroll
    for (int k = 0; k < 8; k++) o.byte4(4 + 4 * k, st[k]);
    const uint32_t r6 = rotr32(e, 6), r11 = rotr32(e, 11), r25 = rotr32(e, 25), S1 = r6 ^ r11 ^ r25;
    o.byte4(36, r6); o.byte4(40, r11); o.byte4(44, r25); o.byte4(48, S1);
    const uint32_t ef = e & f, neg = ~e & g, ch = ef ^ neg;
    o.byte4(52, ef); o.byte4(56, neg); o.byte4(60, ch);
    const uint32_t r2 = rotr32(a, 2), r13 = rotr32(a, 13), r22 = rotr32(a, 22), S0 = r2 ^ r13 ^ r22;
    o.byte4(64, r2); o.byte4(68, r13); o.byte4(72, r22); o.byte4(76, S0);
    const uint32_t ab = a & b, ac = a & c, bc = b & c, mj = ab ^ ac ^ bc;
    o.byte4(80, ab); o.byte4(84, ac); o.byte4(88, bc); o.byte4(92, mj);
    const uint64_t t1w = (uint64_t)h + S1 + ch + TR_K[t] + w;
    const uint32_t t1 = (uint32_t)t1w;
    o.byte4(96, t1); o.put(100, t1w >> 32);
    const uint64_t t2w = (uint64_t)S0 + mj;
    const uint32_t t2 = (uint32_t)t2w;
    o.byte4(101, t2); o.put(105, t2w >> 32);
    const uint64_t aw = (uint64_t)t1 + t2, ew = (uint64_t)d + t1;
    o.byte4(106, (uint32_t)aw); o.put(110, aw >> 32);
    o.byte4(111, (uint32_t)ew); o.put(115, ew >> 32);
    // schedule step producing w_{t+16} (rows 0..47; zero on the last 16 rows of a chunk)
    if (t < 48) {
        const uint32_t w1 = W[t + 1], w14 = W[t + 14], w9 = W[t + 9];
        const uint32_t q7 = rotr32(w1, 7), q18 = rotr32(w1, 18), q3 = w1 >> 3, s0 = q7 ^ q18 ^ q3;
        const uint32_t q17 = rotr32(w14, 17), q19 = rotr32(w14, 19), q10 = w14 >> 10, s1 = q17 ^ q19 ^ q10;
        o.byte4(116, w1); o.byte4(120, q7); o.byte4(124, q18); o.byte4(128, q3); o.byte4(132, s0);
        o.byte4(136, w14); o.byte4(140, q17); o.byte4(144, q19); o.byte4(148, q10); o.byte4(152, s1);
        const uint64_t ww = (uint64_t)s1 + w9 + s0 + w;
        o.byte4(156, w9); o.byte4(160, (uint32_t)ww); o.put(164, ww >> 32);
    } else {
#pragma unroll 1
        for (int col = 116; col <= 164; col++) o.put(col, 0);
    }
    o.put(165, t == 0); o.put(166, t == 63); o.put(167, end_bit && t == 63); o.put(168, digest_bit && t == 63);
#pragma unroll
    for (int k = 0; k < 6; k++) o.put(169 + k, (t >> k) & 1);
    o.put(175, TR_K[t]);
}

// blockIdx.y = circuit: its chunks / flags start chunk_stride chunks further, its table TR_COLS * n_rows elements further
__global__ void __launch_bounds__(128) sha256_trace_kernel(const uint32_t *__restrict__ chunks, const uint8_t *__restrict__ end_bits,
                                                           const uint8_t *__restrict__ digest_bits, uint32_t n_chunks, size_t chunk_stride,
                                                           size_t n_rows, uint64_t *__restrict__ trace) {
    __shared__ uint32_t sW[4][64], sS[4][64][8];
    const uint32_t wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t blk = blockIdx.x * 4 + wid;
    if (blk >= n_chunks) return;
    chunks += (size_t)blockIdx.y * chunk_stride * 16;
    end_bits += (size_t)blockIdx.y * chunk_stride;
    digest_bits += (size_t)blockIdx.y * chunk_stride;
    trace += (size_t)blockIdx.y * TR_COLS * n_rows;
    if (lane == 0) {
        // entering state: IV at the first chunk of a request, else the chaining value of the chunks before it
        uint32_t s0 = blk;
        while (s0 > 0 && !end_bits[s0 - 1]) s0--;
        uint32_t hst[8];
#pragma unroll
        for (int k = 0; k < 8; k++) hst[k] = TR_IV[k];
        for (uint32_t q = s0; q < blk; q++) {
            uint32_t w[16];
#pragma unroll
            for (int k = 0; k < 16; k++) w[k] = chunks[(size_t)q * 16 + k];
            sha256_compress(hst, w);
        }
        uint32_t *W = sW[wid];
        for (int k = 0; k < 16; k++) W[k] = chunks[(size_t)blk * 16 + k];
        for (int t = 16; t < 64; t++) {
            const uint32_t w1 = W[t - 15], w14 = W[t - 2];
            W[t] = (rotr32(w14, 17) ^ rotr32(w14, 19) ^ (w14 >> 10)) + W[t - 7] + (rotr32(w1, 7) ^ rotr32(w1, 18) ^ (w1 >> 3)) + W[t - 16];
        }
        uint32_t s[8];
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = hst[k];
        for (int t = 0; t < 64; t++) {
#pragma unroll
            for (int k = 0; k < 8; k++) sS[wid][t][k] = s[k];
            tr_round(s, W[t], TR_K[t]);
        }
    }
    __syncwarp();
    const bool eb = end_bits[blk] != 0, db = digest_bits[blk] != 0;
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
        const int t = (int)lane + 32 * half;
        TraceRow o{trace + (size_t)blk * 64 + t, n_rows};
        tr_expand_row(o, t, sW[wid], sS[wid][t], eb, db);
    }
}

// ---- SHA-512 (the EdDSA accelerator: SHA512(R ‖ A ‖ M), PX/frontend/hash/sha/sha512/curta.rs): 80 rows per 128-byte chunk,
// 64-bit words as 8 little-endian byte limbs; same construction as above ----
struct TraceRow64 {
    uint64_t *base;
    size_t stride;
    __device__ __forceinline__ void byte8(int col, uint64_t v) const {
#pragma unroll
        for (int k = 0; k < 8; k++) __stcs(base + (size_t)(col + k) * stride, (v >> (8 * k)) & 0xff);
    }
    __device__ __forceinline__ void put(int col, uint64_t v) const { __stcs(base + (size_t)col * stride, v); }
};

__device__ __forceinline__ void tr512_round(uint64_t s[8], uint64_t w, uint64_t k) {
    const uint64_t S1 = rotr64(s[4], 14) ^ rotr64(s[4], 18) ^ rotr64(s[4], 41), ch = (s[4] & s[5]) ^ (~s[4] & s[6]);
    const uint64_t t1 = s[7] + S1 + ch + k + w;
    const uint64_t S0 = rotr64(s[0], 28) ^ rotr64(s[0], 34) ^ rotr64(s[0], 39), mj = (s[0] & s[1]) ^ (s[0] & s[2]) ^ (s[1] & s[2]);
    const uint64_t t2 = S0 + mj;
    s[7] = s[6]; s[6] = s[5]; s[5] = s[4]; s[4] = s[3] + t1; s[3] = s[2]; s[2] = s[1]; s[1] = s[0]; s[0] = t1 + t2;
}

__device__ __forceinline__ void tr512_expand_row(const TraceRow64 &o, int t, const uint64_t *W, const uint64_t *st, bool end_bit, bool digest_bit) {
    typedef unsigned __int128 u128;
    const uint64_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7], w = W[t];
    o.byte8(0, w);
#pragma unroll
    for (int k = 0; k < 8; k++) o.byte8(8 + 8 * k, st[k]);
    const uint64_t r14 = rotr64(e, 14), r18 = rotr64(e, 18), r41 = rotr64(e, 41), S1 = r14 ^ r18 ^ r41;
    o.byte8(72, r14); o.byte8(80, r18); o.byte8(88, r41); o.byte8(96, S1);
    const uint64_t ef = e & f, neg = ~e & g, ch = ef ^ neg;
    o.byte8(104, ef); o.byte8(112, neg); o.byte8(120, ch);
    const uint64_t r28 = rotr64(a, 28), r34 = rotr64(a, 34), r39 = rotr64(a, 39), S0 = r28 ^ r34 ^ r39;
    o.byte8(128, r28); o.byte8(136, r34); o.byte8(144, r39); o.byte8(152, S0);
    const uint64_t ab = a & b, ac = a & c, bc = b & c, mj = ab ^ ac ^ bc;
    o.byte8(160, ab); o.byte8(168, ac); o.byte8(176, bc); o.byte8(184, mj);
    const u128 t1w = (u128)h + S1 + ch + K512[t] + w;
    const uint64_t t1 = (uint64_t)t1w;
    o.byte8(192, t1); o.put(200, (uint64_t)(t1w >> 64));
    const u128 t2w = (u128)S0 + mj;
    const uint64_t t2 = (uint64_t)t2w;
    o.byte8(201, t2); o.put(209, (uint64_t)(t2w >> 64));
    const u128 aw = (u128)t1 + t2, ew = (u128)d + t1;
    o.byte8(210, (uint64_t)aw); o.put(218, (uint64_t)(aw >> 64));
    o.byte8(219, (uint64_t)ew); o.put(227, (uint64_t)(ew >> 64));
    if (t < 64) {
        const uint64_t w1 = W[t + 1], w14 = W[t + 14], w9 = W[t + 9];
        const uint64_t q1 = rotr64(w1, 1), q8 = rotr64(w1, 8), q7 = w1 >> 7, s0 = q1 ^ q8 ^ q7;
        const uint64_t q19 = rotr64(w14, 19), q61 = rotr64(w14, 61), q6 = w14 >> 6, s1 = q19 ^ q61 ^ q6;
        o.byte8(228, w1); o.byte8(236, q1); o.byte8(244, q8); o.byte8(252, q7); o.byte8(260, s0);
        o.byte8(268, w14); o.byte8(276, q19); o.byte8(284, q61); o.byte8(292, q6); o.byte8(300, s1);
        const u128 ww = (u128)s1 + w9 + s0 + w;
        o.byte8(308, w9); o.byte8(316, (uint64_t)ww); o.put(324, (uint64_t)(ww >> 64));
    } else {
#pragma unroll 1
        for (int col = 228; col <= 324; col++) o.put(col, 0);
    }
    o.put(325, t == 0); o.put(326, t == 79); o.put(327, end_bit && t == 79); o.put(328, digest_bit && t == 79);
#pragma unroll
    for (int k = 0; k < 7; k++) o.put(329 + k, (t >> k) & 1);
    o.put(336, (uint32_t)K512[t]); o.put(337, K512[t] >> 32);
}

__global__ void __launch_bounds__(64) sha512_trace_kernel(const uint64_t *__restrict__ chunks, const uint8_t *__restrict__ end_bits,
                                                          const uint8_t *__restrict__ digest_bits, uint32_t n_chunks, size_t n_rows,
                                                          uint64_t *__restrict__ trace) {
    __shared__ uint64_t sW[2][80], sS[2][80][8];
    const uint32_t wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t blk = blockIdx.x * 2 + wid;
    if (blk >= n_chunks) return;
    if (lane == 0) {
        uint32_t s0 = blk;
        while (s0 > 0 && !end_bits[s0 - 1]) s0--;
        uint64_t hst[8];
        sha512_init(hst);
        for (uint32_t q = s0; q < blk; q++) {
            uint64_t w[16];
#pragma unroll
            for (int k = 0; k < 16; k++) w[k] = chunks[(size_t)q * 16 + k];
            sha512_compress(hst, w);
        }
        uint64_t *W = sW[wid];
        for (int k = 0; k < 16; k++) W[k] = chunks[(size_t)blk * 16 + k];
        for (int t = 16; t < 80; t++) {
            const uint64_t w1 = W[t - 15], w14 = W[t - 2];
            W[t] = (rotr64(w14, 19) ^ rotr64(w14, 61) ^ (w14 >> 6)) + W[t - 7] + (rotr64(w1, 1) ^ rotr64(w1, 8) ^ (w1 >> 7)) + W[t - 16];
        }
        uint64_t s[8];
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = hst[k];
        for (int t = 0; t < 80; t++) {
#pragma unroll
            for (int k = 0; k < 8; k++) sS[wid][t][k] = s[k];
            tr512_round(s, W[t], K512[t]);
        }
    }
    __syncwarp();
    const bool eb = end_bits[blk] != 0, db = digest_bits[blk] != 0;
#pragma unroll 1
    for (int part = 0; part < 3; part++) {
        const int t = (int)lane + 32 * part;
        if (t >= 80) break;
        TraceRow64 o{trace + (size_t)blk * 80 + t, n_rows};
        tr512_expand_row(o, t, sW[wid], sS[wid][t], eb, db);
    }
}

}  // namespace bsx

using namespace bsx;

// SHA-256 trace of n_chunks padded chunks (bsx_hash_input_data layout): trace = BSX_SHA256_TRACE_COLS columns of
// 2^log_rows rows each (column-major, u64 field elements), rows beyond 64 * n_chunks zero.
// _batch_: n_circuits accelerators of the same shape in ONE launch (the map circuits of a range have the same request
// schedule): circuit c reads its chunks / flags at c * chunk_stride chunks and writes its own table after the others'.  One
// circuit is 312 CTAs -- two per SM, all ramp-up and tail; 32 of them in one grid keep the machine full.
extern "C" int bsx_sha256_trace_batch_dev(bsx_ctx *ctx, void *stream, const uint32_t *padded_chunks, const uint8_t *end_bits,
                                          const uint8_t *digest_bits, uint32_t n_chunks, uint32_t n_circuits, size_t chunk_stride,
                                          uint32_t log_rows, uint64_t *trace) {
    BSX_REQUIRE(ctx, ctx && padded_chunks && end_bits && digest_bits && trace && log_rows <= 30 && n_circuits <= 65535);
    const size_t n_rows = (size_t)1 << log_rows;
    BSX_REQUIRE(ctx, (size_t)n_chunks * 64 <= n_rows && (n_circuits <= 1 || chunk_stride >= n_chunks));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t used = (size_t)n_chunks * 64;
    if (used < n_rows && n_circuits)
        BSX_CUDA(ctx, cudaMemset2DAsync(trace + used, n_rows * sizeof(uint64_t), 0, (n_rows - used) * sizeof(uint64_t), (size_t)TR_COLS * n_circuits, st));
    if (n_chunks == 0 || n_circuits == 0) return BSX_OK;
    sha256_trace_kernel<<<dim3((n_chunks + 3) / 4, n_circuits), 128, 0, st>>>(padded_chunks, end_bits, digest_bits, n_chunks, chunk_stride, n_rows, trace);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

extern "C" int bsx_sha256_trace_dev(bsx_ctx *ctx, void *stream, const uint32_t *padded_chunks, const uint8_t *end_bits,
                                    const uint8_t *digest_bits, uint32_t n_chunks, uint32_t log_rows, uint64_t *trace) {
    return bsx_sha256_trace_batch_dev(ctx, stream, padded_chunks, end_bits, digest_bits, n_chunks, 1, n_chunks, log_rows, trace);
}

// SHA-512 trace: BSX_SHA512_TRACE_COLS columns of 2^log_rows rows, 80 rows per 128-byte chunk (padded_chunks = u64 words,
// bsx_hash_input_data with sha512 = 1)
extern "C" int bsx_sha512_trace_dev(bsx_ctx *ctx, void *stream, const uint64_t *padded_chunks, const uint8_t *end_bits,
                                    const uint8_t *digest_bits, uint32_t n_chunks, uint32_t log_rows, uint64_t *trace) {
    BSX_REQUIRE(ctx, ctx && padded_chunks && end_bits && digest_bits && trace && log_rows <= 30);
    const size_t n_rows = (size_t)1 << log_rows;
    BSX_REQUIRE(ctx, (size_t)n_chunks * 80 <= n_rows);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t used = (size_t)n_chunks * 80;
    if (used < n_rows)
        BSX_CUDA(ctx, cudaMemset2DAsync(trace + used, n_rows * sizeof(uint64_t), 0, (n_rows - used) * sizeof(uint64_t), BSX_SHA512_TRACE_COLS, st));
    if (n_chunks == 0) return BSX_OK;
    sha512_trace_kernel<<<(n_chunks + 1) / 2, 64, 0, st>>>(padded_chunks, end_bits, digest_bits, n_chunks, n_rows, trace);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

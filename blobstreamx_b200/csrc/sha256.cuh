// SHA-256 device primitives for sm_100a.
//
// Replaces the per-request CPU hint  HashDigestHint<SHA256>::hint -> SHA256::hash
// (PX/frontend/hash/curta/digest_hint.rs:30-38, PX/frontend/hash/sha/sha256/curta.rs:94-102).
// All state lives in registers; the 64 round constants are folded into immediates by full
// unrolling; rotations are funnel shifts (SHF), Ch/Maj/xor3 are single LOP3s.
#pragma once
#include <stdint.h>

#include "sha256_tail_table.cuh"

namespace bsx {

__device__ __forceinline__ uint32_t rotr32(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t ch32(uint32_t e, uint32_t f, uint32_t g) {
    uint32_t r;  // (e & f) | (~e & g)  == 0xCA
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(e), "r"(f), "r"(g));
    return r;
}
__device__ __forceinline__ uint32_t maj32(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;  // majority == 0xE8
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

__device__ __forceinline__ void sha256_init(uint32_t st[8]) {
    st[0] = 0x6a09e667u; st[1] = 0xbb67ae85u; st[2] = 0x3c6ef372u; st[3] = 0xa54ff53au;
    st[4] = 0x510e527fu; st[5] = 0x9b05688cu; st[6] = 0x1f83d9abu; st[7] = 0x5be0cd19u;
}

// One compression.  w[16] is the big-endian-decoded chunk and is clobbered (rolling schedule).
__device__ __forceinline__ void sha256_rounds(uint32_t st[8], uint32_t w[16]) {
    constexpr uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
        0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
        0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
        0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
        0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
        0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
        0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            uint32_t s0 = xor3(rotr32(w15, 7), rotr32(w15, 18), w15 >> 3);
            uint32_t s1 = xor3(rotr32(w2, 17), rotr32(w2, 19), w2 >> 10);
            w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
        }
        uint32_t t1 = h + xor3(rotr32(e, 6), rotr32(e, 11), rotr32(e, 25)) + ch32(e, f, g) + K[i] + w[i & 15];
        uint32_t t2 = xor3(rotr32(a, 2), rotr32(a, 13), rotr32(a, 22)) + maj32(a, b, c);
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// The same compression as a loop: rounds 0..15 unrolled, then three passes of 16 schedule steps + 16 rounds (after 16
// rounds both the a..h rotation and the rolling schedule index are back where they started, so a pass is a loop body
// with no register shuffling).  ~10 KB of SASS instead of 22 KB; the round constants of the looped part come from
// constant memory, indexed by the pass.
__device__ __constant__ uint32_t BSX_SHA256_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
// Pipe steering (r02i).  The ALU pipe (SHF / LOP3 / IADD3: one warp instruction per 2 cycles per sub-partition) is what
// binds every SHA-256 kernel here (ncu: 92 % busy, `math_pipe_throttle` the top stall) while the FMA pipe idles at 7 %.
// An addition written as  a * ONE + b  with ONE read from constant memory cannot be folded back into IADD3 by ptxas
// and issues as IMAD on the FMA pipe.  BSX_SHA_FMA_ADDS is a mask of the addition groups moved there:
//   1: the T1 chain   2: e = d + T1   4: T2 and a = T1 + T2   8: the message schedule
// Measured on B200 (profiles/r02j_ab_sha.txt): 2048-leaf trees 2.60 -> 2.10 ms per 4096 with mask 5 (2.48 with 13), map stage
// 2.02 -> 1.87 ms with 13 (1.92 with 5), input shaping 0.955 -> 0.936; each translation unit picks its mask before the include.
#ifndef BSX_SHA_FMA_ADDS
#define BSX_SHA_FMA_ADDS 5
#endif
__device__ __constant__ uint32_t BSX_SHA_ONE = 1;
__device__ __forceinline__ uint32_t add_fma(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(BSX_SHA_ONE), "r"(b));
    return r;
}
template <int BIT>
__device__ __forceinline__ uint32_t sha_add(uint32_t a, uint32_t b) {
    if constexpr ((BSX_SHA_FMA_ADDS & BIT) != 0) return add_fma(a, b);
    else return a + b;
}
#define BSX_SHA_ROUND(wk)                                                                              \
    {                                                                                                  \
        const uint32_t t1 = sha_add<1>(sha_add<1>(sha_add<1>(h, (wk)), ch32(e, f, g)),                   \
                                       xor3(rotr32(e, 6), rotr32(e, 11), rotr32(e, 25)));               \
        const uint32_t t2 = sha_add<4>(xor3(rotr32(a, 2), rotr32(a, 13), rotr32(a, 22)), maj32(a, b, c)); \
        h = g; g = f; f = e; e = sha_add<2>(d, t1); d = c; c = b; b = a; a = sha_add<4>(t1, t2);       \
    }
__device__ __forceinline__ void sha256_rounds_looped(uint32_t st[8], uint32_t w[16]) {
    constexpr uint32_t K0[16] = {0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
                                 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174};
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int i = 0; i < 16; i++) BSX_SHA_ROUND(K0[i] + w[i])
#pragma unroll 1
    for (int p = 1; p < 4; p++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            const uint32_t s0 = xor3(rotr32(w15, 7), rotr32(w15, 18), w15 >> 3);
            const uint32_t s1 = xor3(rotr32(w2, 17), rotr32(w2, 19), w2 >> 10);
            w[i] = sha_add<8>(sha_add<8>(sha_add<8>(w[i], w[(i + 9) & 15]), s0), s1);
            BSX_SHA_ROUND(sha_add<8>(BSX_SHA256_K[16 * p + i], w[i]))
        }
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// The 64 unrolled rounds are ~1500 instructions (24 KB).  Inlined at every hash site a kernel grows
// to hundreds of KB and stalls on instruction fetch (r01a profile: 52 % of samples `no_instructions`),
// so the compression is ONE real function per module; state and block travel by value in registers.
//   BSX_SHA_VARIANT 0: inline everywhere (the r01a build)   1: one generic function
//                   2: generic + a function specialised for the constant tail block of 65-byte messages
//                   3: generic + a tail function whose schedule comes from a 256-row table (below)
//                   4: 3 with both functions looped (16 rounds unrolled + loop): 15 KB of SASS instead of 37 KB --
//                      with two fully unrolled functions the proofs kernel lost its gain to instruction-cache misses
//                      (ncu r01m: no_instruction 5.3 stall cycles per issue against 0.09)
#ifndef BSX_SHA_VARIANT
#define BSX_SHA_VARIANT 4  // measured on B200 (profiles/r01b): 0 -> 1.68 ms, 1 -> 1.13 ms, 2 -> 1.15 ms per 8192 map jobs; 3: r01m
#endif
struct sha256_state { uint32_t s[8]; };
struct sha256_block { uint32_t w[16]; };
#if BSX_SHA_VARIANT >= 1
static __device__ __noinline__ sha256_state sha256_compress_fn(sha256_state st, sha256_block blk) {
#if BSX_SHA_VARIANT >= 4
    sha256_rounds_looped(st.s, blk.w);
#else
    sha256_rounds(st.s, blk.w);
#endif
    return st;
}
__device__ __forceinline__ void sha256_compress(uint32_t st[8], uint32_t w[16]) {
    sha256_state s; sha256_block b;
#pragma unroll
    for (int i = 0; i < 8; i++) s.s[i] = st[i];
#pragma unroll
    for (int i = 0; i < 16; i++) b.w[i] = w[i];
    s = sha256_compress_fn(s, b);
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = s.s[i];
}
#else
__device__ __forceinline__ void sha256_compress(uint32_t st[8], uint32_t w[16]) { sha256_rounds(st, w); }
#endif
// second block of a 65-byte message (inner nodes, data-root tuples): byte 64, 0x80, zeros, bit length 520
#if BSX_SHA_VARIANT >= 3
// Only schedule word 0 of this block depends on the data, so W[16..63] + K[16..63] is one of 256 precomputed rows
// (sha256_tail_table.cuh, 48 KB, L1-resident): the 48 schedule steps (~480 ALU-pipe instructions of ~1390) become
// twelve 16-byte loads on the otherwise idle load pipe.  18 of the 39 compressions per header in the map stage are
// such tail blocks.
static __device__ __noinline__ sha256_state sha256_tail65_fn(sha256_state st, uint32_t last_byte) {
    constexpr uint32_t K0[16] = {0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
                                 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174};
    uint32_t a = st.s[0], b = st.s[1], c = st.s[2], d = st.s[3], e = st.s[4], f = st.s[5], g = st.s[6], h = st.s[7];
    BSX_SHA_ROUND(K0[0] + ((last_byte << 24) | 0x00800000u))
#pragma unroll
    for (int i = 1; i < 15; i++) BSX_SHA_ROUND(K0[i])
    BSX_SHA_ROUND(K0[15] + 65u * 8u)
    const uint4 *row = BSX_SHA_TAIL_WK + 12 * last_byte;
#if BSX_SHA_VARIANT >= 4
#pragma unroll 1
    for (int q = 0; q < 12; q += 2) {          // 8 rounds per pass: the a..h rotation is back in place
        const uint4 wk = __ldg(row + q), wl = __ldg(row + q + 1);
        BSX_SHA_ROUND(wk.x) BSX_SHA_ROUND(wk.y) BSX_SHA_ROUND(wk.z) BSX_SHA_ROUND(wk.w)
        BSX_SHA_ROUND(wl.x) BSX_SHA_ROUND(wl.y) BSX_SHA_ROUND(wl.z) BSX_SHA_ROUND(wl.w)
    }
#else
#pragma unroll
    for (int q = 0; q < 12; q++) {
        const uint4 wk = __ldg(row + q);
        BSX_SHA_ROUND(wk.x) BSX_SHA_ROUND(wk.y) BSX_SHA_ROUND(wk.z) BSX_SHA_ROUND(wk.w)
    }
#endif
    st.s[0] += a; st.s[1] += b; st.s[2] += c; st.s[3] += d; st.s[4] += e; st.s[5] += f; st.s[6] += g; st.s[7] += h;
    return st;
}
__device__ __forceinline__ void sha256_tail65(uint32_t st[8], uint32_t last_byte) {
    sha256_state s;
#pragma unroll
    for (int i = 0; i < 8; i++) s.s[i] = st[i];
    s = sha256_tail65_fn(s, last_byte);
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = s.s[i];
}
#elif BSX_SHA_VARIANT >= 2
static __device__ __noinline__ sha256_state sha256_tail65_fn(sha256_state st, uint32_t last_byte) {
    uint32_t w[16];
    w[0] = (last_byte << 24) | 0x00800000u;
#pragma unroll
    for (int i = 1; i < 15; i++) w[i] = 0;
    w[15] = 65 * 8;
    sha256_rounds(st.s, w);
    return st;
}
__device__ __forceinline__ void sha256_tail65(uint32_t st[8], uint32_t last_byte) {
    sha256_state s;
#pragma unroll
    for (int i = 0; i < 8; i++) s.s[i] = st[i];
    s = sha256_tail65_fn(s, last_byte);
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = s.s[i];
}
#else
__device__ __forceinline__ void sha256_tail65(uint32_t st[8], uint32_t last_byte) {
    uint32_t w[16];
    w[0] = (last_byte << 24) | 0x00800000u;
#pragma unroll
    for (int i = 1; i < 15; i++) w[i] = 0;
    w[15] = 65 * 8;
    sha256_compress(st, w);
}
#endif

// Tendermint inner node  sha256(0x01 ‖ l ‖ r)  on big-endian state words (65 bytes = 2 chunks).
// PX/frontend/merkle/tendermint.rs:108-122
__device__ __forceinline__ void tm_inner_hash(const uint32_t l[8], const uint32_t r[8], uint32_t out[8]) {
    uint32_t w[16];
    w[0] = 0x01000000u | (l[0] >> 8);
#pragma unroll
    for (int i = 1; i < 8; i++) w[i] = __funnelshift_r(l[i], l[i - 1], 8);
    w[8] = __funnelshift_r(r[0], l[7], 8);
#pragma unroll
    for (int i = 1; i < 8; i++) w[8 + i] = __funnelshift_r(r[i], r[i - 1], 8);
    sha256_init(out);
    sha256_compress(out, w);
    sha256_tail65(out, r[7] & 0xffu);
}

// sha256 of `len` bytes given through a byte getter (any address space); generic path used for
// leaves and variable-length requests.  get(i) must be valid for i < len.
template <typename Get>
__device__ __forceinline__ void sha256_bytes(Get get, uint32_t len, uint32_t out[8]) {
    sha256_init(out);
    uint32_t nblk = (len + 9 + 63) >> 6;
    for (uint32_t b = 0; b < nblk; b++) {
        uint32_t w[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t idx = b * 64 + k * 4 + j;
                uint32_t byte = idx < len ? (uint32_t)get(idx) : (idx == len ? 0x80u : 0u);
                v = (v << 8) | byte;
            }
            w[k] = v;
        }
        if (b == nblk - 1) { w[14] = 0; w[15] = len << 3; }  // len < 2^29
        sha256_compress(out, w);
    }
}

// Leaf node sha256(0x00 ‖ leaf) with leaf bytes behind a getter.  tendermint.rs:95-106
template <typename Get>
__device__ __forceinline__ void tm_leaf_hash(Get get, uint32_t leaf_len, uint32_t out[8]) {
    sha256_bytes([&](uint32_t i) -> uint8_t { return i == 0 ? (uint8_t)0 : get(i - 1); }, leaf_len + 1, out);
}

__device__ __forceinline__ void load_digest_be(const uint8_t* p, uint32_t d[8]) {  // p 4-byte aligned
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = bswap32(q[i]);
}
__device__ __forceinline__ void store_digest_be(uint8_t* p, const uint32_t d[8]) {  // p 16-byte aligned
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(bswap32(d[0]), bswap32(d[1]), bswap32(d[2]), bswap32(d[3]));
    q[1] = make_uint4(bswap32(d[4]), bswap32(d[5]), bswap32(d[6]), bswap32(d[7]));
}
__device__ __forceinline__ bool digest_eq(const uint32_t a[8], const uint32_t b[8]) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x |= a[i] ^ b[i];
    return x == 0;
}

}  // namespace bsx

/* TEST INFRASTRUCTURE -- CPU restatement of the SHA-256 execution trace.  The reference fills its trace
 * with starkyx's TraceWriter (PX/frontend/hash/curta/stark.rs:107-133); the column assignment is starkyx's (un-vendored),
 * so the layout here is the one documented in include/bsx.h (BSX_SHA256_TRACE_COLS) and PARITY IS UNPINNED against the
 * reference; tests/test_oracle_trace.py pins it by recomputing every digest from the columns. */
#include "bsx_oracle.h"

#include <string.h>

static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
static const uint32_t IV256[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};

static uint32_t ror(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

#define COLS 176
static void put4(uint64_t *row0, size_t stride, int col, uint32_t v) {
    for (int k = 0; k < 4; k++) row0[(size_t)(col + k) * stride] = (v >> (8 * k)) & 0xff;
}

void orc_sha256_trace(const uint32_t *chunks, const uint8_t *end_bits, const uint8_t *digest_bits, uint32_t n_chunks,
                      uint32_t log_rows, uint64_t *trace) {
    const size_t n_rows = (size_t)1 << log_rows;
    memset(trace, 0, sizeof(uint64_t) * COLS * n_rows);
    uint32_t hst[8];
    memcpy(hst, IV256, sizeof hst);
    for (uint32_t b = 0; b < n_chunks; b++) {
        uint32_t W[64], s[8];
        for (int k = 0; k < 16; k++) W[k] = chunks[(size_t)b * 16 + k];
        for (int t = 16; t < 64; t++)
            W[t] = (ror(W[t - 2], 17) ^ ror(W[t - 2], 19) ^ (W[t - 2] >> 10)) + W[t - 7] +
                   (ror(W[t - 15], 7) ^ ror(W[t - 15], 18) ^ (W[t - 15] >> 3)) + W[t - 16];
        memcpy(s, hst, sizeof s);
        for (int t = 0; t < 64; t++) {
            uint64_t *o = trace + (size_t)b * 64 + t;
            const uint32_t a = s[0], bb = s[1], c = s[2], d = s[3], e = s[4], f = s[5], g = s[6], h = s[7], w = W[t];
            put4(o, n_rows, 0, w);
            for (int k = 0; k < 8; k++) put4(o, n_rows, 4 + 4 * k, s[k]);
            const uint32_t r6 = ror(e, 6), r11 = ror(e, 11), r25 = ror(e, 25), S1 = r6 ^ r11 ^ r25;
            put4(o, n_rows, 36, r6); put4(o, n_rows, 40, r11); put4(o, n_rows, 44, r25); put4(o, n_rows, 48, S1);
            const uint32_t ef = e & f, ng = ~e & g, ch = ef ^ ng;
            put4(o, n_rows, 52, ef); put4(o, n_rows, 56, ng); put4(o, n_rows, 60, ch);
            const uint32_t r2 = ror(a, 2), r13 = ror(a, 13), r22 = ror(a, 22), S0 = r2 ^ r13 ^ r22;
            put4(o, n_rows, 64, r2); put4(o, n_rows, 68, r13); put4(o, n_rows, 72, r22); put4(o, n_rows, 76, S0);
            const uint32_t ab = a & bb, ac = a & c, bc = bb & c, mj = ab ^ ac ^ bc;
            put4(o, n_rows, 80, ab); put4(o, n_rows, 84, ac); put4(o, n_rows, 88, bc); put4(o, n_rows, 92, mj);
            const uint64_t t1w = (uint64_t)h + S1 + ch + K256[t] + w, t2w = (uint64_t)S0 + mj;
            const uint32_t t1 = (uint32_t)t1w, t2 = (uint32_t)t2w;
            put4(o, n_rows, 96, t1); o[(size_t)100 * n_rows] = t1w >> 32;
            put4(o, n_rows, 101, t2); o[(size_t)105 * n_rows] = t2w >> 32;
            const uint64_t aw = (uint64_t)t1 + t2, ew = (uint64_t)d + t1;
            put4(o, n_rows, 106, (uint32_t)aw); o[(size_t)110 * n_rows] = aw >> 32;
            put4(o, n_rows, 111, (uint32_t)ew); o[(size_t)115 * n_rows] = ew >> 32;
            if (t < 48) {
                const uint32_t w1 = W[t + 1], w14 = W[t + 14], w9 = W[t + 9];
                const uint32_t q7 = ror(w1, 7), q18 = ror(w1, 18), q3 = w1 >> 3, s0 = q7 ^ q18 ^ q3;
                const uint32_t q17 = ror(w14, 17), q19 = ror(w14, 19), q10 = w14 >> 10, s1 = q17 ^ q19 ^ q10;
                put4(o, n_rows, 116, w1); put4(o, n_rows, 120, q7); put4(o, n_rows, 124, q18); put4(o, n_rows, 128, q3); put4(o, n_rows, 132, s0);
                put4(o, n_rows, 136, w14); put4(o, n_rows, 140, q17); put4(o, n_rows, 144, q19); put4(o, n_rows, 148, q10); put4(o, n_rows, 152, s1);
                const uint64_t ww = (uint64_t)s1 + w9 + s0 + w;
                put4(o, n_rows, 156, w9); put4(o, n_rows, 160, (uint32_t)ww); o[(size_t)164 * n_rows] = ww >> 32;
            }
            o[(size_t)165 * n_rows] = t == 0; o[(size_t)166 * n_rows] = t == 63;
            o[(size_t)167 * n_rows] = end_bits[b] && t == 63; o[(size_t)168 * n_rows] = digest_bits[b] && t == 63;
            for (int k = 0; k < 6; k++) o[(size_t)(169 + k) * n_rows] = (t >> k) & 1;
            o[(size_t)175 * n_rows] = K256[t];
            s[7] = g; s[6] = f; s[5] = e; s[4] = (uint32_t)ew; s[3] = c; s[2] = bb; s[1] = a; s[0] = (uint32_t)aw;
        }
        for (int k = 0; k < 8; k++) hst[k] += s[k];
        if (end_bits[b]) memcpy(hst, IV256, sizeof hst);
    }
}

/* ---- SHA-512 (the EdDSA accelerator): 80 rows per 128-byte chunk, 64-bit words as 8 byte limbs (include/bsx.h,
 * BSX_SHA512_TRACE_COLS = 338) ---- */
static const uint64_t K512[80] = {
    0x428a2f98d728ae22ULL, 0x7137449123ef65cdULL, 0xb5c0fbcfec4d3b2fULL, 0xe9b5dba58189dbbcULL, 0x3956c25bf348b538ULL, 0x59f111f1b605d019ULL,
    0x923f82a4af194f9bULL, 0xab1c5ed5da6d8118ULL, 0xd807aa98a3030242ULL, 0x12835b0145706fbeULL, 0x243185be4ee4b28cULL, 0x550c7dc3d5ffb4e2ULL,
    0x72be5d74f27b896fULL, 0x80deb1fe3b1696b1ULL, 0x9bdc06a725c71235ULL, 0xc19bf174cf692694ULL, 0xe49b69c19ef14ad2ULL, 0xefbe4786384f25e3ULL,
    0x0fc19dc68b8cd5b5ULL, 0x240ca1cc77ac9c65ULL, 0x2de92c6f592b0275ULL, 0x4a7484aa6ea6e483ULL, 0x5cb0a9dcbd41fbd4ULL, 0x76f988da831153b5ULL,
    0x983e5152ee66dfabULL, 0xa831c66d2db43210ULL, 0xb00327c898fb213fULL, 0xbf597fc7beef0ee4ULL, 0xc6e00bf33da88fc2ULL, 0xd5a79147930aa725ULL,
    0x06ca6351e003826fULL, 0x142929670a0e6e70ULL, 0x27b70a8546d22ffcULL, 0x2e1b21385c26c926ULL, 0x4d2c6dfc5ac42aedULL, 0x53380d139d95b3dfULL,
    0x650a73548baf63deULL, 0x766a0abb3c77b2a8ULL, 0x81c2c92e47edaee6ULL, 0x92722c851482353bULL, 0xa2bfe8a14cf10364ULL, 0xa81a664bbc423001ULL,
    0xc24b8b70d0f89791ULL, 0xc76c51a30654be30ULL, 0xd192e819d6ef5218ULL, 0xd69906245565a910ULL, 0xf40e35855771202aULL, 0x106aa07032bbd1b8ULL,
    0x19a4c116b8d2d0c8ULL, 0x1e376c085141ab53ULL, 0x2748774cdf8eeb99ULL, 0x34b0bcb5e19b48a8ULL, 0x391c0cb3c5c95a63ULL, 0x4ed8aa4ae3418acbULL,
    0x5b9cca4f7763e373ULL, 0x682e6ff3d6b2b8a3ULL, 0x748f82ee5defb2fcULL, 0x78a5636f43172f60ULL, 0x84c87814a1f0ab72ULL, 0x8cc702081a6439ecULL,
    0x90befffa23631e28ULL, 0xa4506cebde82bde9ULL, 0xbef9a3f7b2c67915ULL, 0xc67178f2e372532bULL, 0xca273eceea26619cULL, 0xd186b8c721c0c207ULL,
    0xeada7dd6cde0eb1eULL, 0xf57d4f7fee6ed178ULL, 0x06f067aa72176fbaULL, 0x0a637dc5a2c898a6ULL, 0x113f9804bef90daeULL, 0x1b710b35131c471bULL,
    0x28db77f523047d84ULL, 0x32caab7b40c72493ULL, 0x3c9ebe0a15c9bebcULL, 0x431d67c49c100d4cULL, 0x4cc5d4becb3e42b6ULL, 0x597f299cfc657e2aULL,
    0x5fcb6fab3ad6faecULL, 0x6c44198c4a475817ULL};
static const uint64_t IV512[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                  0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
static uint64_t ror64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
static void put8(uint64_t *row0, size_t stride, int col, uint64_t v) {
    for (int k = 0; k < 8; k++) row0[(size_t)(col + k) * stride] = (v >> (8 * k)) & 0xff;
}
typedef unsigned __int128 u128t;

void orc_sha512_trace(const uint64_t *chunks, const uint8_t *end_bits, const uint8_t *digest_bits, uint32_t n_chunks,
                      uint32_t log_rows, uint64_t *trace) {
    const size_t n = (size_t)1 << log_rows;
    memset(trace, 0, sizeof(uint64_t) * 338 * n);
    uint64_t hst[8];
    memcpy(hst, IV512, sizeof hst);
    for (uint32_t b = 0; b < n_chunks; b++) {
        uint64_t W[80], s[8];
        for (int k = 0; k < 16; k++) W[k] = chunks[(size_t)b * 16 + k];
        for (int t = 16; t < 80; t++)
            W[t] = (ror64(W[t - 2], 19) ^ ror64(W[t - 2], 61) ^ (W[t - 2] >> 6)) + W[t - 7] +
                   (ror64(W[t - 15], 1) ^ ror64(W[t - 15], 8) ^ (W[t - 15] >> 7)) + W[t - 16];
        memcpy(s, hst, sizeof s);
        for (int t = 0; t < 80; t++) {
            uint64_t *o = trace + (size_t)b * 80 + t;
            const uint64_t a = s[0], bb = s[1], c = s[2], d = s[3], e = s[4], f = s[5], g = s[6], h = s[7], w = W[t];
            put8(o, n, 0, w);
            for (int k = 0; k < 8; k++) put8(o, n, 8 + 8 * k, s[k]);
            const uint64_t r14 = ror64(e, 14), r18 = ror64(e, 18), r41 = ror64(e, 41), S1 = r14 ^ r18 ^ r41;
            put8(o, n, 72, r14); put8(o, n, 80, r18); put8(o, n, 88, r41); put8(o, n, 96, S1);
            const uint64_t ef = e & f, ng = ~e & g, ch = ef ^ ng;
            put8(o, n, 104, ef); put8(o, n, 112, ng); put8(o, n, 120, ch);
            const uint64_t r28 = ror64(a, 28), r34 = ror64(a, 34), r39 = ror64(a, 39), S0 = r28 ^ r34 ^ r39;
            put8(o, n, 128, r28); put8(o, n, 136, r34); put8(o, n, 144, r39); put8(o, n, 152, S0);
            const uint64_t ab = a & bb, ac = a & c, bc = bb & c, mj = ab ^ ac ^ bc;
            put8(o, n, 160, ab); put8(o, n, 168, ac); put8(o, n, 176, bc); put8(o, n, 184, mj);
            const u128t t1w = (u128t)h + S1 + ch + K512[t] + w, t2w = (u128t)S0 + mj;
            const uint64_t t1 = (uint64_t)t1w, t2 = (uint64_t)t2w;
            put8(o, n, 192, t1); o[(size_t)200 * n] = (uint64_t)(t1w >> 64);
            put8(o, n, 201, t2); o[(size_t)209 * n] = (uint64_t)(t2w >> 64);
            const u128t aw = (u128t)t1 + t2, ew = (u128t)d + t1;
            put8(o, n, 210, (uint64_t)aw); o[(size_t)218 * n] = (uint64_t)(aw >> 64);
            put8(o, n, 219, (uint64_t)ew); o[(size_t)227 * n] = (uint64_t)(ew >> 64);
            if (t < 64) {
                const uint64_t w1 = W[t + 1], w14 = W[t + 14], w9 = W[t + 9];
                const uint64_t q1 = ror64(w1, 1), q8 = ror64(w1, 8), q7 = w1 >> 7, s0 = q1 ^ q8 ^ q7;
                const uint64_t q19 = ror64(w14, 19), q61 = ror64(w14, 61), q6 = w14 >> 6, s1 = q19 ^ q61 ^ q6;
                put8(o, n, 228, w1); put8(o, n, 236, q1); put8(o, n, 244, q8); put8(o, n, 252, q7); put8(o, n, 260, s0);
                put8(o, n, 268, w14); put8(o, n, 276, q19); put8(o, n, 284, q61); put8(o, n, 292, q6); put8(o, n, 300, s1);
                const u128t ww = (u128t)s1 + w9 + s0 + w;
                put8(o, n, 308, w9); put8(o, n, 316, (uint64_t)ww); o[(size_t)324 * n] = (uint64_t)(ww >> 64);
            }
            o[(size_t)325 * n] = t == 0; o[(size_t)326 * n] = t == 79;
            o[(size_t)327 * n] = end_bits[b] && t == 79; o[(size_t)328 * n] = digest_bits[b] && t == 79;
            for (int k = 0; k < 7; k++) o[(size_t)(329 + k) * n] = (t >> k) & 1;
            o[(size_t)336 * n] = (uint32_t)K512[t]; o[(size_t)337 * n] = K512[t] >> 32;
            s[7] = g; s[6] = f; s[5] = e; s[4] = (uint64_t)ew; s[3] = c; s[2] = bb; s[1] = a; s[0] = (uint64_t)aw;
        }
        for (int k = 0; k < 8; k++) hst[k] += s[k];
        if (end_bits[b]) memcpy(hst, IV512, sizeof hst);
    }
}

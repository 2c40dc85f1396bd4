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

// Ed25519 witness arithmetic for sm_100a: every value the plonky2x EC hints produce for one signature.
//   EcOpResultHint::{ScalarMul, Decompress, Add}   PX/frontend/ecc/curve25519/curta/result_hint.rs:21-50
//   verification schedule                           PX/frontend/ecc/curve25519/ed25519/eddsa.rs:161-203
//   h = LE512(SHA512(R‖A‖M)) div/rem l              PX/frontend/uint/num/biguint/mod.rs:451-488
//
// Design (one thread per signature, FMA-pipe bound):
//   * field elements are 10 signed limbs of 26/25 bits; products accumulate in 64 bits (IMAD.WIDE);
//     limb bounds follow the classic scheme: mul/sq accept |limb| <= 1.65*2^26 (even) / 2^25 (odd) and
//     return |limb| <= 1.01*2^25 / 2^24; add/sub take mul outputs and feed only mul inputs.
//   * fe_mul / fe_sq / fe_sqn are real calls (by-value structs travel in registers), so the whole
//     kernel stays a few thousand instructions and lives in the instruction cache -- fully inlined
//     it would be >1 MB of SASS.
//   * s*G: 32 unsigned 8-bit windows over a precomputed affine table ((y+x, y-x, 2dxy) entries, 1 MB, built
//     once per context on the device); h*A: signed 4-bit windows over an 8-entry cached table (1A..8A) in
//     local memory, 4 doublings + 1 addition per window, extended coordinates (the quad-lane kernel keeps
//     its own unsigned 16-entry table, one component per lane); all three projective results share ONE inversion.
//   * h, div: Barrett division of the 512-bit digest by l with a 264-bit reciprocal.
// Affine results in canonical form are unique, so parity with the reference's BigUint affine
// arithmetic (starkyx, un-vendored) is exact; `decompress` returns the EVEN root.
//
// The same source compiles for the host (plain g++) in tests/host_check: that build is TEST
// infrastructure that checks this arithmetic against the oracle on a machine without a GPU.
#pragma once
#include <stdint.h>

#include "fe51d.cuh"

#if defined(__CUDACC__)
#define BSX_HD __host__ __device__ __forceinline__
#define BSX_CALL static __host__ __device__ __noinline__
#else
#define BSX_HD static inline
#define BSX_CALL static __attribute__((noinline))
#endif

namespace bsx {
namespace ed {

struct fe { int32_t v[10]; };

// ---------------------------------------------------------------------------------------------
// field arithmetic mod p = 2^255 - 19
// ---------------------------------------------------------------------------------------------
BSX_HD fe fe_zero() { fe r; for (int i = 0; i < 10; i++) r.v[i] = 0; return r; }
BSX_HD fe fe_one() { fe r = fe_zero(); r.v[0] = 1; return r; }
BSX_HD fe fe_add(const fe &a, const fe &b) { fe r; for (int i = 0; i < 10; i++) r.v[i] = a.v[i] + b.v[i]; return r; }
BSX_HD fe fe_sub(const fe &a, const fe &b) { fe r; for (int i = 0; i < 10; i++) r.v[i] = a.v[i] - b.v[i]; return r; }
BSX_HD fe fe_neg(const fe &a) { fe r; for (int i = 0; i < 10; i++) r.v[i] = -a.v[i]; return r; }
BSX_HD fe fe_select(bool c, const fe &a, const fe &b) { fe r; for (int i = 0; i < 10; i++) r.v[i] = c ? a.v[i] : b.v[i]; return r; }

// balanced carry chain on 64-bit limbs -> |even| <= 2^25, |odd| <= 2^24 (+ a few units)
BSX_HD fe fe_carry64(int64_t h[10]) {
    int64_t c;
#define BSX_FE_CARRY(i, bits, nxt, mul)                      \
    c = (h[i] + ((int64_t)1 << ((bits)-1))) >> (bits);       \
    h[nxt] += c * (mul);                                     \
    h[i] -= c << (bits);
    BSX_FE_CARRY(0, 26, 1, 1) BSX_FE_CARRY(4, 26, 5, 1)
    BSX_FE_CARRY(1, 25, 2, 1) BSX_FE_CARRY(5, 25, 6, 1)
    BSX_FE_CARRY(2, 26, 3, 1) BSX_FE_CARRY(6, 26, 7, 1)
    BSX_FE_CARRY(3, 25, 4, 1) BSX_FE_CARRY(7, 25, 8, 1)
    BSX_FE_CARRY(4, 26, 5, 1) BSX_FE_CARRY(8, 26, 9, 1)
    BSX_FE_CARRY(9, 25, 0, 19)
    BSX_FE_CARRY(0, 26, 1, 1)
#undef BSX_FE_CARRY
    fe r;
    for (int i = 0; i < 10; i++) r.v[i] = (int32_t)h[i];
    return r;
}

// re-balance an add/sub result (same value mod p, limbs back within the mul-output bounds)
BSX_HD fe fe_reduce(const fe &f) {
    int64_t h[10];
    for (int i = 0; i < 10; i++) h[i] = f.v[i];
    return fe_carry64(h);
}

// One parallel carry round in 32-bit arithmetic: same value mod p, every limb back in [-38, 2^26 + 19] (even) /
// [-2, 2^25 + 1] (odd) for inputs up to 3 "units".  Used on the one factor of a 3-unit x 3-unit product that
// becomes the g-side of fe_mul (see its bounds).  ~30 ALU-pipe operations, no dependent chain.
BSX_HD fe fe_tighten(const fe &f) {
    fe r;
    int32_t c[10];
#pragma unroll
    for (int i = 0; i < 10; i++) c[i] = f.v[i] >> ((i & 1) ? 25 : 26);
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const int32_t lowpart = f.v[i] & ((i & 1) ? 0x1ffffff : 0x3ffffff);
        r.v[i] = lowpart + (i == 0 ? 19 * c[9] : c[i - 1]);
    }
    return r;
}

#ifndef BSX_FE_SCHOOLBOOK
// h = f*g.  The limbs are read as five pairs (a_k + 2^26 b_k) 2^(51k); each pair product costs THREE
// 32x32->64 multiplications (Karatsuba on the pair: a c, b d, (a+b)(c+d)) instead of four, so the product is
// 75 IMAD.WIDE instead of 100 -- the fmaheavy pipe IMAD.WIDE runs on is what bounds this kernel.
//   LL_n = sum a_k c_m, HH_n = sum b_k d_m, SS_n = sum (a_k+b_k)(c_m+d_m)   over k+m = n (mod 5), x19 on wrap
//   h[2n+1] = SS_n - LL_n - HH_n;  h[2n] = LL_n + 2 HH_{n-1};  h[0] = LL_0 + 38 HH_4
// Bounds: g at most 2 units (|g_even| <= 2.02*2^25, |g_odd| <= 2.02*2^24: one add/sub of carried values, or a
// fe_tighten output) so that 19 (c_m + d_m) < 2^31; f at most 3 units.  Sums stay below 2^61.
BSX_HD fe fe_mul_inl(const fe &f, const fe &g) {
    int32_t fs[5], gs[5], c19[5], d19[5], gs19[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        fs[k] = f.v[2 * k] + f.v[2 * k + 1];
        gs[k] = g.v[2 * k] + g.v[2 * k + 1];
        c19[k] = 19 * g.v[2 * k];
        d19[k] = 19 * g.v[2 * k + 1];
        gs19[k] = c19[k] + d19[k];
    }
    int64_t LL[5], HH[5], SS[5];
#pragma unroll
    for (int n = 0; n < 5; n++) {
        int64_t ll = 0, hh = 0, ss = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const int m = (n - k + 5) % 5;
            const bool wrap = (k + m) >= 5;
            ll += (int64_t)f.v[2 * k] * (wrap ? c19[m] : g.v[2 * m]);
            hh += (int64_t)f.v[2 * k + 1] * (wrap ? d19[m] : g.v[2 * m + 1]);
            ss += (int64_t)fs[k] * (wrap ? gs19[m] : gs[m]);
        }
        LL[n] = ll; HH[n] = hh; SS[n] = ss;
    }
    int64_t h[10];
#pragma unroll
    for (int n = 0; n < 5; n++) {
        h[2 * n + 1] = SS[n] - LL[n] - HH[n];
        h[2 * n] = LL[n] + (n == 0 ? 38 * HH[4] : 2 * HH[n - 1]);
    }
    return fe_carry64(h);
}

#else
BSX_HD fe fe_mul_inl(const fe &f, const fe &g) {
    int32_t g19[10], f2[10];
#pragma unroll
    for (int i = 0; i < 10; i++) { g19[i] = 19 * g.v[i]; f2[i] = 2 * f.v[i]; }
    int64_t h[10];
#pragma unroll
    for (int k = 0; k < 10; k++) {
        int64_t acc = 0;
#pragma unroll
        for (int i = 0; i < 10; i++) {
            const int j = (k - i + 10) % 10;
            const bool wrap = (i + j) >= 10, dbl = (i & 1) && (j & 1);
            acc += (int64_t)(dbl ? f2[i] : f.v[i]) * (wrap ? g19[j] : g.v[j]);
        }
        h[k] = acc;
    }
    return fe_carry64(h);
}

#endif
BSX_CALL fe fe_mul(const fe f, const fe g) { return fe_mul_inl(f, g); }
// INL selects the inlined body (hot loops of the scalar multiplications: the independent multiplications of one point
// operation are then scheduled together, and no registers are shuffled into a call) or the shared function (everything
// else: keeps the kernel small)
template <bool INL> BSX_HD fe fe_mul_x(const fe &f, const fe &g) { if (INL) return fe_mul_inl(f, g); return fe_mul(f, g); }

// h = f*f (times 2 when `twice`), using the symmetry of the product: 55 IMAD.WIDE.  (The pair-Karatsuba form of the
// squaring is 45 IMAD.WIDE but measured slower on B200 at every occupancy -- 376 vs 341 cycles per squaring per SM
// sub-partition at 8 warps, 438 vs 413 at 2 -- because its extra additions land on the same pipe; profiles/r01j_ubench_femul.txt.)
template <bool TWICE>
BSX_HD fe fe_sq_impl(const fe &f) {
    int32_t f2[10], f19[10], f38[10];
#pragma unroll
    for (int i = 0; i < 10; i++) { f2[i] = 2 * f.v[i]; f19[i] = 19 * f.v[i]; f38[i] = (i & 1) ? 38 * f.v[i] : 0; }   // 38 f_i fits (and is used) only for odd i
    int64_t h[10];
#pragma unroll
    for (int k = 0; k < 10; k++) {
        int64_t acc = 0;
#pragma unroll
        for (int i = 0; i < 10; i++) {
            const int j = (k - i + 10) % 10;
            if (i > j) continue;
            const bool wrap = (i + j) >= 10, dbl = (i & 1) && (j & 1);
            // term = f_i f_j * (dbl?2:1) * (wrap?19:1) * (i<j?2:1)
            int32_t a, b;
            if (i == j) { a = f.v[i]; b = wrap ? (dbl ? f38[i] : f19[i]) : (dbl ? f2[i] : f.v[i]); }
            else if (wrap && (j & 1)) { a = dbl ? f2[i] : f.v[i]; b = f38[j]; }   // 38 f_j fits only for odd j
            else if (wrap) { a = f2[i]; b = f19[j]; }
            else { a = dbl ? f2[i] : f.v[i]; b = f2[j]; }
            acc += (int64_t)a * b;
        }
        h[k] = TWICE ? acc * 2 : acc;
    }
    return fe_carry64(h);
}
BSX_CALL fe fe_sq(const fe f) { return fe_sq_impl<false>(f); }
BSX_CALL fe fe_sq2(const fe f) { return fe_sq_impl<true>(f); }
template <bool INL> BSX_HD fe fe_sq_x(const fe &f) { if (INL) return fe_sq_impl<false>(f); return fe_sq(f); }
template <bool INL> BSX_HD fe fe_sq2_x(const fe &f) { if (INL) return fe_sq_impl<true>(f); return fe_sq2(f); }
// n >= 1 successive squarings (one call for the long chains of inversion / square roots)
BSX_CALL fe fe_sqn(fe f, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) f = fe_sq_impl<false>(f);
    return f;
}

// 32 little-endian bytes (bit 255 ignored) -> balanced limbs
BSX_HD fe fe_frombytes(const uint8_t *s) {
    uint64_t w[4];
    for (int i = 0; i < 4; i++) {
        uint64_t x = 0;
        for (int j = 7; j >= 0; j--) x = (x << 8) | s[8 * i + j];
        w[i] = x;
    }
    w[3] &= 0x7fffffffffffffffULL;
    int64_t h[10];
    const int off[10] = {0, 26, 51, 77, 102, 128, 153, 179, 204, 230};
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const int o = off[i], wi = o >> 6, sh = o & 63, bits = (i & 1) ? 25 : 26;
        uint64_t x = w[wi] >> sh;
        if (sh + bits > 64 && wi < 3) x |= w[wi + 1] << (64 - sh);
        h[i] = (int64_t)(x & (((uint64_t)1 << bits) - 1));
    }
    return fe_carry64(h);
}

// canonical little-endian bytes in [0, p); input within the mul-output bounds
BSX_HD void fe_tobytes(uint8_t *s, const fe &f) {
    int32_t h[10];
    for (int i = 0; i < 10; i++) h[i] = f.v[i];
    int32_t q = (19 * h[9] + (1 << 24)) >> 25;
#pragma unroll
    for (int i = 0; i < 10; i++) q = (h[i] + q) >> ((i & 1) ? 25 : 26);
    h[0] += 19 * q;
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const int bits = (i & 1) ? 25 : 26;
        int32_t c = h[i] >> bits;
        if (i < 9) h[i + 1] += c;
        h[i] -= c << bits;
    }
    uint64_t w[4] = {0, 0, 0, 0};
    const int off[10] = {0, 26, 51, 77, 102, 128, 153, 179, 204, 230};
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const int o = off[i], wi = o >> 6, sh = o & 63;
        w[wi] |= (uint64_t)(uint32_t)h[i] << sh;
        if (sh + 26 > 64 && wi < 3) w[wi + 1] |= (uint64_t)(uint32_t)h[i] >> (64 - sh);
    }
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) s[8 * i + j] = (uint8_t)(w[i] >> (8 * j));
}

BSX_HD bool fe_iszero_bytes(const uint8_t *s) {
    uint8_t x = 0;
    for (int i = 0; i < 32; i++) x |= s[i];
    return x == 0;
}
BSX_HD bool bytes_eq32(const uint8_t *a, const uint8_t *b) {
    uint8_t x = 0;
    for (int i = 0; i < 32; i++) x |= a[i] ^ b[i];
    return x == 0;
}

// z^(2^252 - 3) = z^((p-5)/8)
BSX_HD fe fe_pow22523(const fe &z) {
    fe t0 = fe_sq(z);                 // 2
    fe t1 = fe_sqn(t0, 2);            // 8
    t1 = fe_mul(z, t1);               // 9
    t0 = fe_mul(t0, t1);              // 11
    t0 = fe_sq(t0);                   // 22
    t0 = fe_mul(t1, t0);              // 31 = 2^5-1
    t1 = fe_sqn(t0, 5);
    t0 = fe_mul(t1, t0);              // 2^10-1
    t1 = fe_sqn(t0, 10);
    t1 = fe_mul(t1, t0);              // 2^20-1
    fe t2 = fe_sqn(t1, 20);
    t1 = fe_mul(t2, t1);              // 2^40-1
    t1 = fe_sqn(t1, 10);
    t0 = fe_mul(t1, t0);              // 2^50-1
    t1 = fe_sqn(t0, 50);
    t1 = fe_mul(t1, t0);              // 2^100-1
    t2 = fe_sqn(t1, 100);
    t1 = fe_mul(t2, t1);              // 2^200-1
    t1 = fe_sqn(t1, 50);
    t0 = fe_mul(t1, t0);              // 2^250-1
    t0 = fe_sqn(t0, 2);               // 2^252-4
    return fe_mul(t0, z);             // 2^252-3
}

// z^(p-2)
BSX_HD fe fe_invert(const fe &z) {
    fe t0 = fe_sq(z);                 // 2
    fe t1 = fe_sqn(t0, 2);            // 8
    t1 = fe_mul(z, t1);               // 9
    t0 = fe_mul(t0, t1);              // 11
    fe t2 = fe_sq(t0);                // 22
    t1 = fe_mul(t1, t2);              // 31
    t2 = fe_sqn(t1, 5);
    t1 = fe_mul(t2, t1);              // 2^10-1
    t2 = fe_sqn(t1, 10);
    t2 = fe_mul(t2, t1);              // 2^20-1
    fe t3 = fe_sqn(t2, 20);
    t2 = fe_mul(t3, t2);              // 2^40-1
    t2 = fe_sqn(t2, 10);
    t1 = fe_mul(t2, t1);              // 2^50-1
    t2 = fe_sqn(t1, 50);
    t2 = fe_mul(t2, t1);              // 2^100-1
    t3 = fe_sqn(t2, 100);
    t2 = fe_mul(t3, t2);              // 2^200-1
    t2 = fe_sqn(t2, 50);
    t1 = fe_mul(t2, t1);              // 2^250-1
    t1 = fe_sqn(t1, 5);               // 2^255-32
    return fe_mul(t1, t0);            // 2^255-21
}

BSX_HD fe fe_const(const int32_t c[10]) { fe r; for (int i = 0; i < 10; i++) r.v[i] = c[i]; return r; }
#define BSX_FE_D {56195235, 13857412, 51736253, 6949390, 114729, 24766616, 60832955, 30306712, 48412415, 21499315}
#define BSX_FE_2D {45281625, 27714825, 36363642, 13898781, 229458, 15978800, 54557047, 27058993, 29715967, 9444199}
#define BSX_FE_SQRTM1 {34513072, 25610706, 9377949, 3500415, 12389472, 33281959, 41962654, 31548777, 326685, 11406482}
#define BSX_FE_GX {52811034, 25909283, 16144682, 17082669, 27570973, 30858332, 40966398, 8378388, 20764389, 8758491}
#define BSX_FE_GY {40265304, 26843545, 13421772, 20132659, 26843545, 6710886, 53687091, 13421772, 40265318, 26843545}

// ---------------------------------------------------------------------------------------------
// points: extended twisted Edwards coordinates, a = -1
// ---------------------------------------------------------------------------------------------
struct ge_p3 { fe X, Y, Z, T; };           // x = X/Z, y = Y/Z, xy = T/Z
struct ge_p1p1 { fe X, Y, Z, T; };         // completed: x = X/Z, y = Y/T
struct ge_cached { fe YpX, YmX, Z, T2d; };
struct ge_niels { fe ypx, ymx, xy2d; };    // affine precomputed (Z = 1)

BSX_HD ge_p3 ge_identity() { ge_p3 r; r.X = fe_zero(); r.Y = fe_one(); r.Z = fe_one(); r.T = fe_zero(); return r; }
BSX_HD ge_p3 ge_from_affine(const fe &x, const fe &y) { ge_p3 r; r.X = x; r.Y = y; r.Z = fe_one(); r.T = fe_mul(x, y); return r; }
BSX_HD ge_cached ge_to_cached(const ge_p3 &p) {
    const int32_t d2[10] = BSX_FE_2D;
    ge_cached c; c.YpX = fe_add(p.Y, p.X); c.YmX = fe_sub(p.Y, p.X); c.Z = p.Z; c.T2d = fe_mul(p.T, fe_const(d2));
    return c;
}
// completed -> extended (4M); with_t=false skips T (3M) when the next operation is a doubling
// Operand order follows fe_mul's bounds: X and T of a completed point are 3-unit values (a - (yy + xx),
// 2zz - (yy - xx), 2zz +- c), Y is 2 units, Z is 2 (doubling) or 3 (addition); the g-side is Y or the tightened T.
template <bool INL = false>
BSX_HD ge_p3 ge_p1p1_to_p3(const ge_p1p1 &p, bool with_t) {
    const fe tt = fe_tighten(p.T);
    ge_p3 r; r.X = fe_mul_x<INL>(p.X, tt); r.Y = fe_mul_x<INL>(p.Z, p.Y); r.Z = fe_mul_x<INL>(p.Z, tt);
    r.T = with_t ? fe_mul_x<INL>(p.X, p.Y) : fe_zero();
    return r;
}
// doubling (uses X, Y, Z only): 3S + 1 S2
template <bool INL = false>
BSX_HD ge_p1p1 ge_dbl(const ge_p3 &p) {
    ge_p1p1 r;
    fe xx = fe_sq_x<INL>(p.X), yy = fe_sq_x<INL>(p.Y), zz2 = fe_sq2_x<INL>(p.Z);
    fe a = fe_sq_x<INL>(fe_add(p.X, p.Y));
    r.Y = fe_add(yy, xx); r.Z = fe_sub(yy, xx); r.X = fe_sub(a, r.Y); r.T = fe_sub(zz2, r.Z);
    return r;
}
// p + q, q cached: 4M
template <bool INL = false>
BSX_HD ge_p1p1 ge_add_cached(const ge_p3 &p, const ge_cached &q) {
    ge_p1p1 r;
    fe a = fe_mul_x<INL>(fe_add(p.Y, p.X), q.YpX), b = fe_mul_x<INL>(fe_sub(p.Y, p.X), q.YmX);
    fe c = fe_mul_x<INL>(q.T2d, p.T), zz = fe_mul_x<INL>(p.Z, q.Z);
    fe d = fe_add(zz, zz);
    r.X = fe_sub(a, b); r.Y = fe_add(a, b); r.Z = fe_add(d, c); r.T = fe_sub(d, c);
    return r;
}
// p + q, q affine precomputed: 3M
template <bool INL = false>
BSX_HD ge_p1p1 ge_add_niels(const ge_p3 &p, const ge_niels &q) {
    ge_p1p1 r;
    fe a = fe_mul_x<INL>(fe_add(p.Y, p.X), q.ypx), b = fe_mul_x<INL>(fe_sub(p.Y, p.X), q.ymx);
    fe c = fe_mul_x<INL>(q.xy2d, p.T);
    fe d = fe_add(p.Z, p.Z);
    r.X = fe_sub(a, b); r.Y = fe_add(a, b); r.Z = fe_add(d, c); r.T = fe_sub(d, c);
    return r;
}

// ---------------------------------------------------------------------------------------------
// decompress (starkyx chip::ec::edwards::ed25519::decompress, call site result_hint.rs:40):
// returns the affine point and `root` = the EVEN square root of (y^2-1)/(d y^2+1); x = root when
// the sign bit is 0, p - root otherwise.  ok=false when the ratio is not a square (the reference
// panics there); the outputs are then the identity and root = 0 (same convention in the oracle).
// ---------------------------------------------------------------------------------------------
BSX_HD bool ge_decompress(const uint8_t *in, fe &x, fe &y, uint8_t x_bytes[32], uint8_t y_bytes[32], uint8_t root_bytes[32]) {
    const int32_t dc[10] = BSX_FE_D, sm1[10] = BSX_FE_SQRTM1;
    const bool sign = (in[31] >> 7) != 0;
    y = fe_frombytes(in);
    fe yy = fe_sq(y);
    fe u = fe_sub(yy, fe_one());
    fe v = fe_add(fe_mul(yy, fe_const(dc)), fe_one());
    fe v3 = fe_mul(fe_sq(v), v);
    fe uv7 = fe_mul(fe_mul(fe_sq(v3), v), u);
    fe r = fe_mul(fe_mul(fe_pow22523(uv7), v3), u);   // u v^3 (u v^7)^((p-5)/8)
    fe vxx = fe_mul(fe_sq(r), v);
    uint8_t a[32], b[32];
    fe_tobytes(a, fe_reduce(fe_sub(vxx, u)));  // v r^2 - u
    bool ok = fe_iszero_bytes(a);
    if (!ok) {
        fe_tobytes(b, fe_reduce(fe_add(vxx, u)));  // v r^2 + u
        if (fe_iszero_bytes(b)) { r = fe_mul(r, fe_const(sm1)); ok = true; }
    }
    if (!ok) {
        x = fe_zero(); y = fe_one();
        for (int i = 0; i < 32; i++) { x_bytes[i] = 0; y_bytes[i] = 0; root_bytes[i] = 0; }
        y_bytes[0] = 1;
        return false;
    }
    fe_tobytes(root_bytes, r);
    if (root_bytes[0] & 1) { r = fe_neg(r); fe_tobytes(root_bytes, r); }
    x = sign ? fe_neg(r) : r;
    if (sign) fe_tobytes(x_bytes, x); else for (int i = 0; i < 32; i++) x_bytes[i] = root_bytes[i];
    fe_tobytes(y_bytes, y);
    return true;
}

// ---------------------------------------------------------------------------------------------
// scalar multiplication
// ---------------------------------------------------------------------------------------------
#define BSX_ED_BASE_WINDOWS 32
#define BSX_ED_BASE_ENTRIES 255
// table[w*255 + (d-1)] = d * 256^w * G   (d = 1..255), affine precomputed form: 8-bit windows, so s*G is at most
// 32 mixed additions (7M each) and no doubling.  8160 entries x 128 B = 1 MB, resident in L2; each lane gathers
// its own entry with eight 16-byte loads.
struct alignas(16) ge_niels_slot { int32_t v[32]; };   // ypx[10] ymx[10] xy2d[10] pad[2]
BSX_HD ge_niels ge_niels_load(const ge_niels_slot *slot) {
    ge_niels q;
#if defined(__CUDA_ARCH__)
    int32_t v[32];
    const int4 *p = reinterpret_cast<const int4 *>(slot);
#pragma unroll
    for (int i = 0; i < 8; i++) { const int4 t = __ldg(p + i); v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w; }
#else
    const int32_t *v = slot->v;
#endif
#pragma unroll
    for (int i = 0; i < 10; i++) { q.ypx.v[i] = v[i]; q.ymx.v[i] = v[10 + i]; q.xy2d.v[i] = v[20 + i]; }
    return q;
}
BSX_HD void ge_niels_store(ge_niels_slot *slot, const ge_niels &q) {
    for (int i = 0; i < 10; i++) { slot->v[i] = q.ypx.v[i]; slot->v[10 + i] = q.ymx.v[i]; slot->v[20 + i] = q.xy2d.v[i]; }
    slot->v[30] = 0; slot->v[31] = 0;
}
template <bool INL = false>
BSX_HD ge_p3 ge_scalarmult_base(const uint8_t s[32], const ge_niels_slot *table) {
    ge_p3 acc = ge_identity();
#pragma unroll 1
    for (int w = 0; w < BSX_ED_BASE_WINDOWS; w++) {
        const uint32_t dgt = s[w];
        if (dgt) {
            const ge_niels q = ge_niels_load(table + w * BSX_ED_BASE_ENTRIES + (dgt - 1));
            acc = ge_p1p1_to_p3<INL>(ge_add_niels<INL>(acc, q), true);
        }
    }
    return acc;
}

// -q for an addend in cached form: swap Y+X and Y-X, negate 2dT
BSX_HD ge_cached ge_cached_cneg(const ge_cached &q, bool neg) {
    ge_cached r;
    r.YpX = fe_select(neg, q.YmX, q.YpX); r.YmX = fe_select(neg, q.YpX, q.YmX); r.Z = q.Z;
    r.T2d = fe_select(neg, fe_neg(q.T2d), q.T2d);
    return r;
}

// scalar * P for an arbitrary 256-bit scalar (EcOpResultHint::ScalarMul takes the U256 unreduced).
// Signed radix-16 digits e_w in [-8, 8) (w = 0..63, carry-out e_64 in {0, 1}): the table holds 1P..8P only -- 7
// additions to build instead of 14 and 1.25 KB of local memory per thread instead of 2.4 KB (the table is read from
// L1/L2 on every window; halving it matters once four CTAs share an SM's L1).  Doublings start at the first non-zero digit.
template <bool INL = false>
BSX_HD ge_p3 ge_scalarmult(const uint8_t s[32], const ge_p3 &P) {
    ge_cached tab[8];                        // tab[d-1] = d*P
    {
        ge_p3 cur = P;
        tab[0] = ge_to_cached(cur);
#pragma unroll 1
        for (int d = 2; d <= 8; d++) {
            cur = ge_p1p1_to_p3(ge_add_cached(cur, tab[0]), true);
            tab[d - 1] = ge_to_cached(cur);
        }
    }
    int8_t e[65];
    {
        int carry = 0;
#pragma unroll 1
        for (int w = 0; w < 64; w++) {
            int d = (int)((s[w >> 1] >> ((w & 1) * 4)) & 15) + carry;
            carry = d >= 8;
            e[w] = (int8_t)(d - 16 * carry);
        }
        e[64] = (int8_t)carry;
    }
    ge_p3 acc = ge_identity();
    bool started = false;
#pragma unroll 1
    for (int w = 64; w >= 0; w--) {
        if (started) {
#pragma unroll 1
            for (int k = 0; k < 4; k++) acc = ge_p1p1_to_p3<INL>(ge_dbl<INL>(acc), k == 3);
        }
        const int d = e[w];
        // T is consumed only by additions: the 4th doubling of a window produces it, an addition
        // drops it again (the next step is a doubling) except in the last window.
        if (d) {
            const ge_cached q = ge_cached_cneg(tab[(d < 0 ? -d : d) - 1], d < 0);
            acc = ge_p1p1_to_p3<INL>(ge_add_cached<INL>(acc, q), w == 0);
            started = true;
        }
    }
    return acc;
}

// one entry of the s*G table: d * 256^w * G in affine precomputed form (run once per context)
BSX_HD ge_niels ge_base_table_entry(int w, int d) {
    const int32_t gx[10] = BSX_FE_GX, gy[10] = BSX_FE_GY, d2[10] = BSX_FE_2D;
    uint8_t s[32];
    for (int i = 0; i < 32; i++) s[i] = 0;
    s[w] = (uint8_t)d;
    // canonical constants are one-sided 2-unit values; balance them so that Y + X stays a legal g operand
    ge_p3 r = ge_scalarmult(s, ge_from_affine(fe_reduce(fe_const(gx)), fe_reduce(fe_const(gy))));
    fe zi = fe_invert(r.Z);
    fe x = fe_mul(r.X, zi), y = fe_mul(r.Y, zi);
    ge_niels n;
    n.ypx = fe_reduce(fe_add(y, x)); n.ymx = fe_reduce(fe_sub(y, x)); n.xy2d = fe_mul(fe_mul(x, y), fe_const(d2));
    return n;
}

// ---------------------------------------------------------------------------------------------
// 512-bit digest div/rem l  (l = 2^252 + 27742317777372353535851937790883648493)
// ---------------------------------------------------------------------------------------------
BSX_HD void sc_divrem_l(const uint8_t digest[64], uint8_t rem[32], uint8_t div[40]) {
    const uint32_t L[9] = {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0u, 0u, 0u, 0x10000000u, 0u};
    const uint32_t MU[9] = {0xa2c131b3u, 0xd9ce5a30u, 0x86329a7eu, 0x106215d0u, 0xfffffeb2u, 0xffffffffu,
                            0xffffffffu, 0xffffffffu, 0xffu};  // floor(2^516 / l)
    uint32_t x[16];
    for (int i = 0; i < 16; i++)
        x[i] = (uint32_t)digest[4 * i] | ((uint32_t)digest[4 * i + 1] << 8) | ((uint32_t)digest[4 * i + 2] << 16) |
               ((uint32_t)digest[4 * i + 3] << 24);
    // x1 = x >> 248 (9 limbs)
    uint32_t x1[9];
    for (int i = 0; i < 9; i++) {
        uint32_t lo = x[7 + i] >> 24, hi = (7 + i + 1 < 16) ? (x[8 + i] << 8) : 0u;
        x1[i] = lo | hi;
    }
    // t = x1 * MU (18 limbs); q = t >> 268
    uint32_t t[18];
    for (int i = 0; i < 18; i++) t[i] = 0;
    for (int i = 0; i < 9; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < 9; j++) {
            uint64_t m = (uint64_t)x1[i] * MU[j] + t[i + j] + carry;
            t[i + j] = (uint32_t)m;
            carry = m >> 32;
        }
        t[i + 9] = (uint32_t)carry;
    }
    uint32_t q[9];
    for (int i = 0; i < 9; i++) {  // shift right by 268 = 8 limbs + 12 bits
        uint32_t lo = t[8 + i] >> 12, hi = (9 + i < 18) ? (t[9 + i] << 20) : 0u;
        q[i] = lo | hi;
    }
    // r = x - q*l  (mod 2^288)
    uint32_t ql[9];
    for (int i = 0; i < 9; i++) ql[i] = 0;
    for (int i = 0; i < 9; i++) {
        uint64_t carry = 0;
        for (int j = 0; i + j < 9; j++) {
            uint64_t m = (uint64_t)q[i] * L[j] + ql[i + j] + carry;
            ql[i + j] = (uint32_t)m;
            carry = m >> 32;
        }
    }
    uint32_t r[9];
    {
        uint64_t borrow = 0;
        for (int i = 0; i < 9; i++) {
            uint64_t dd = (uint64_t)x[i] - ql[i] - borrow;
            r[i] = (uint32_t)dd;
            borrow = (dd >> 32) & 1;
        }
    }
    // at most one correction (q_hat in {q-1, q})
    for (int it = 0; it < 2; it++) {
        bool ge = true;
        for (int i = 8; i >= 0; i--) {
            if (r[i] != L[i]) { ge = r[i] > L[i]; break; }
        }
        if (!ge) break;
        uint64_t borrow = 0;
        for (int i = 0; i < 9; i++) {
            uint64_t dd = (uint64_t)r[i] - L[i] - borrow;
            r[i] = (uint32_t)dd;
            borrow = (dd >> 32) & 1;
        }
        for (int i = 0; i < 9; i++) { if (++q[i] != 0) break; }
    }
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) rem[4 * i + j] = (uint8_t)(r[i] >> (8 * j));
    for (int i = 0; i < 9; i++) for (int j = 0; j < 4; j++) div[4 * i + j] = (uint8_t)(q[i] >> (8 * j));
    for (int j = 36; j < 40; j++) div[j] = 0;
}

BSX_HD bool sc_lt_l(const uint8_t s[32]) {
    const uint8_t Lb[32] = {0xed, 0xd3, 0xf5, 0x5c, 0x1a, 0x63, 0x12, 0x58, 0xd6, 0x9c, 0xf7, 0xa2, 0xde, 0xf9, 0xde, 0x14,
                            0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0x10};
    for (int i = 31; i >= 0; i--) {
        if (s[i] != Lb[i]) return s[i] < Lb[i];
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// one signature: the record of include/bsx.h (BSX_SIG_OUT_BYTES) from pk, sig and the SHA-512 digest
//   [0..64) digest  [64..96) h  [96..136) div  [136..200) sG  [200..264) A  [264..296) A_root
//   [296..360) hA  [360..424) Rp  [424..456) R_root  [456..520) Rp+hA  [520..524) flags
// flags: 1 s<l, 2 A decompressed, 4 R decompressed, 8 sG == Rp+hA
// ---------------------------------------------------------------------------------------------
// INL: inline the field arithmetic of the scalar-multiplication loops (+10 % when the kernel has the SMs to itself:
// independent multiplications of a point operation are scheduled together and nothing is shuffled into calls; next to
// the SHA-256 kernels of bsx_header_range the larger loop bodies cost more than they save -- measured r01o).
template <bool INL = false>
BSX_HD void ed25519_witness_core(const uint8_t pk[32], const uint8_t sig[64], const uint8_t digest[64],
                                 const ge_niels_slot *base_table, uint8_t *out) {
    for (int i = 0; i < 64; i++) out[i] = digest[i];
    sc_divrem_l(digest, out + 64, out + 96);
    uint32_t flags = sc_lt_l(sig + 32) ? 1u : 0u;
    // A, Rp
    fe ax, ay, rx, ry;
    if (ge_decompress(pk, ax, ay, out + 200, out + 232, out + 264)) flags |= 2u;
    if (ge_decompress(sig, rx, ry, out + 360, out + 392, out + 424)) flags |= 4u;
    // sG, hA, Rp + hA in projective form
    ge_p3 sg = ge_scalarmult_base<INL>(sig + 32, base_table);
    ge_p3 ha = ge_scalarmult<INL>(out + 64, ge_from_affine(ax, ay));
    ge_p3 sum = ge_p1p1_to_p3(ge_add_cached(ha, ge_to_cached(ge_from_affine(rx, ry))), false);
    // one inversion for the three Z's
    fe z12 = fe_mul(sg.Z, ha.Z);
    fe inv = fe_invert(fe_mul(z12, sum.Z));
    fe isum = fe_mul(inv, z12);
    fe i12 = fe_mul(inv, sum.Z);
    fe isg = fe_mul(i12, ha.Z), iha = fe_mul(i12, sg.Z);
    fe_tobytes(out + 136, fe_mul(sg.X, isg)); fe_tobytes(out + 168, fe_mul(sg.Y, isg));
    fe_tobytes(out + 296, fe_mul(ha.X, iha)); fe_tobytes(out + 328, fe_mul(ha.Y, iha));
    fe_tobytes(out + 456, fe_mul(sum.X, isum)); fe_tobytes(out + 488, fe_mul(sum.Y, isum));
    if (bytes_eq32(out + 136, out + 456) && bytes_eq32(out + 168, out + 488)) flags |= 8u;
    out[520] = (uint8_t)flags; out[521] = 0; out[522] = 0; out[523] = 0;
    for (int i = 524; i < 576; i++) out[i] = 0;
}

}  // namespace ed

// ---------------------------------------------------------------------------------------------
// The same group arithmetic over the FP64-pipe field (fe51d.cuh): 5 x 51-bit limbs in doubles.
// Unit bookkeeping ("u" = a carried value, |limb| <= 2^50 + 2^14; a product needs |f_i g_j| < 2^103, i.e. units
// multiplying to at most 7): every multiplication / squaring returns 1u; sums are noted where they arise.
// ---------------------------------------------------------------------------------------------
namespace edd {
using ed::fe;

#define BSX_FED_D {929955233495222.0, 466365720129213.0, -589740348686295.0, -217950738957124.0, -809005158844672.0}
#define BSX_FED_2D {-391889346694823.0, 932731440258427.0, 1072319116312658.0, -435901477914249.0, 633789495995904.0}
#define BSX_FED_SQRTM1 {-533094393274192.0, 234908883556510.0, -18285341111200.0, -134597186663265.0, 765476049583134.0}
BSX_HD fed fed_const(const double c[5]) { fed r; for (int i = 0; i < 5; i++) r.v[i] = c[i]; return r; }

// Record bytes are produced in thread-local arrays and leave as 8-byte words (every field of the 576-byte record starts at a
// multiple of 8): r02p's capture of the batch kernel had the read-backs of bytes it had just stored to global memory
// (`long_scoreboard`) and its byte-wide stores (`lg_throttle`) as the top stalls.  dst must be 8-byte aligned.
template <int N>
BSX_HD void put_bytes(uint8_t *dst, const uint8_t *src) {
#if defined(__CUDA_ARCH__)
    uint2 *q = reinterpret_cast<uint2 *>(dst);
#pragma unroll
    for (int k = 0; k < N / 8; k++) {
        const uint8_t *b = src + 8 * k;
        q[k] = make_uint2((uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24),
                          (uint32_t)b[4] | ((uint32_t)b[5] << 8) | ((uint32_t)b[6] << 16) | ((uint32_t)b[7] << 24));
    }
#else
    for (int k = 0; k < N; k++) dst[k] = src[k];
#endif
}
BSX_HD void put_flags_and_padding(uint8_t *out, uint32_t flags) {   // [520, 576): flags byte, zeros
#if defined(__CUDA_ARCH__)
    uint2 *q = reinterpret_cast<uint2 *>(out + 520);
    q[0] = make_uint2(flags & 0xffu, 0u);
#pragma unroll
    for (int k = 1; k < 7; k++) q[k] = make_uint2(0u, 0u);
#else
    out[520] = (uint8_t)flags;
    for (int i = 521; i < 576; i++) out[i] = 0;
#endif
}

// 10 x 25.5-bit integer limbs <-> 5 x 51-bit double limbs (limb pair 2k, 2k+1 = bits 51k .. 51k+50)
BSX_HD fed fed_from_fe(const fe &f) {
    int64_t E[5];
#pragma unroll
    for (int k = 0; k < 5; k++) E[k] = (int64_t)f.v[2 * k] + ((int64_t)f.v[2 * k + 1] << 26);
    return fed_carry(E);
}
BSX_HD fe fe_from_fed(const fed &f) {
    int64_t h[10];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const int64_t l = fed_limb_i64(f.v[k]);
        h[2 * k] = l & 0x3ffffff;
        h[2 * k + 1] = l >> 26;
    }
    return ed::fe_carry64(h);
}

struct ged_p3 { fed X, Y, Z, T; };
struct ged_p1p1 { fed X, Y, Z, T; };
struct ged_cached { fed YpX, YmX, Z, T2d; };   // YpX, YmX: 2u;  Z, T2d: 1u
struct ged_niels { fed ypx, ymx, xy2d; };      // 1u each (converted table entries)

BSX_HD ged_p3 ged_identity() { ged_p3 r; r.X = fed_zero(); r.Y = fed_one(); r.Z = fed_one(); r.T = fed_zero(); return r; }
BSX_HD ged_p3 ged_from_affine(const fed &x, const fed &y) { ged_p3 r; r.X = x; r.Y = y; r.Z = fed_one(); r.T = fed_mul(x, y); return r; }
BSX_HD ged_cached ged_to_cached(const ged_p3 &p) {
    const double d2[5] = BSX_FED_2D;
    ged_cached c; c.YpX = fed_add(p.Y, p.X); c.YmX = fed_sub(p.Y, p.X); c.Z = p.Z; c.T2d = fed_mul(p.T, fed_const(d2));
    return c;
}
// completed -> extended.  Units of the completed point: after a doubling X 3u, Y 2u, Z 2u, T 1u; after an addition all 2u
// (T 1u for a table addition) -- every product below stays within 6.
template <bool INL = false>
BSX_HD ged_p3 ged_p1p1_to_p3(const ged_p1p1 &p, bool with_t) {
    ged_p3 r; r.X = fed_mul_x<INL>(p.X, p.T); r.Y = fed_mul_x<INL>(p.Z, p.Y); r.Z = fed_mul_x<INL>(p.Z, p.T);
    r.T = with_t ? fed_mul_x<INL>(p.X, p.Y) : fed_zero();
    return r;
}
// doubling (X, Y, Z of a 1u point): T' = 2 ZZ - YY + XX is combined on the folded 64-bit columns BEFORE the carry, so it
// comes out carried (1u) instead of as a 3-unit difference -- X' = A - YY - XX is 3u, and 3u x 3u would exceed the product bound.
template <bool INL = false>
BSX_HD ged_p1p1 ged_dbl(const ged_p3 &p) {
    ged_p1p1 r;
    const fed_raw exx = fed_sq_raw_x<INL>(p.X), eyy = fed_sq_raw_x<INL>(p.Y), ezz = fed_sq_raw_x<INL>(p.Z);
    const fed a = fed_sq_x<INL>(fed_add(p.X, p.Y));                    // (2u)^2
    int64_t et[5];
#pragma unroll
    for (int k = 0; k < 5; k++) et[k] = 2 * ezz.E[k] - eyy.E[k] + exx.E[k];   // < 4 * 2^59.7
    const fed xx = fed_carry(exx.E), yy = fed_carry(eyy.E);
    r.T = fed_carry(et);
    r.Y = fed_add(yy, xx); r.Z = fed_sub(yy, xx); r.X = fed_sub(a, r.Y);
    return r;
}
// p + q, q cached: 2 Z1 Z2 comes doubled out of the multiplication (1u), so Z' and T' are 2u
template <bool INL = false>
BSX_HD ged_p1p1 ged_add_cached(const ged_p3 &p, const ged_cached &q) {
    ged_p1p1 r;
    const fed a = fed_mul_x<INL>(fed_add(p.Y, p.X), q.YpX), b = fed_mul_x<INL>(fed_sub(p.Y, p.X), q.YmX);
    const fed c = fed_mul_x<INL>(q.T2d, p.T), d = fed_mul2_x<INL>(p.Z, q.Z);
    r.X = fed_sub(a, b); r.Y = fed_add(a, b); r.Z = fed_add(d, c); r.T = fed_sub(d, c);
    return r;
}
// p + q, q affine precomputed: d = 2 Z1 is 2u, so T' = d - c is re-carried (Z' stays 3u: 3u x 2u and 3u x 1u are fine)
template <bool INL = false>
BSX_HD ged_p1p1 ged_add_niels(const ged_p3 &p, const ged_niels &q) {
    ged_p1p1 r;
    const fed a = fed_mul_x<INL>(fed_add(p.Y, p.X), q.ypx), b = fed_mul_x<INL>(fed_sub(p.Y, p.X), q.ymx);
    const fed c = fed_mul_x<INL>(q.xy2d, p.T);
    const fed d = fed_add(p.Z, p.Z);
    r.X = fed_sub(a, b); r.Y = fed_add(a, b); r.Z = fed_add(d, c); r.T = fed_reduce(fed_sub(d, c));
    return r;
}

// decompress: as ed::ge_decompress, the square-root chain on the FP64 pipe
BSX_HD bool ged_decompress(const uint8_t *in, fed &x, fed &y, uint8_t x_bytes[32], uint8_t y_bytes[32], uint8_t root_bytes[32]) {
    const double dc[5] = BSX_FED_D, sm1[5] = BSX_FED_SQRTM1;
    const bool sign = (in[31] >> 7) != 0;
    const fe yi = ed::fe_frombytes(in);
    y = fed_from_fe(yi);
    const fed yy = fed_sq(y);
    const fed u = fed_sub(yy, fed_one());                       // 1u + 1
    const fed v = fed_add(fed_mul(yy, fed_const(dc)), fed_one());
    const fed v3 = fed_mul(fed_sq(v), v);
    const fed uv7 = fed_mul(fed_mul(fed_sq(v3), v), u);
    fed r = fed_mul(fed_mul(fed_pow22523(uv7), v3), u);   // u v^3 (u v^7)^((p-5)/8)
    const fed vxx = fed_mul(fed_sq(r), v);
    uint8_t a[32], b[32];
    ed::fe_tobytes(a, fe_from_fed(fed_sub(vxx, u)));      // v r^2 - u
    bool ok = ed::fe_iszero_bytes(a);
    if (!ok) {
        ed::fe_tobytes(b, fe_from_fed(fed_add(vxx, u)));  // v r^2 + u
        if (ed::fe_iszero_bytes(b)) { r = fed_mul(r, fed_const(sm1)); ok = true; }
    }
    if (!ok) {
        x = fed_zero(); y = fed_one();
        for (int i = 0; i < 32; i++) { x_bytes[i] = 0; y_bytes[i] = 0; root_bytes[i] = 0; }
        y_bytes[0] = 1;
        return false;
    }
    ed::fe_tobytes(root_bytes, fe_from_fed(r));
    if (root_bytes[0] & 1) { r = fed_neg(r); ed::fe_tobytes(root_bytes, fe_from_fed(r)); }
    x = sign ? fed_neg(r) : r;
    if (sign) ed::fe_tobytes(x_bytes, fe_from_fed(x)); else for (int i = 0; i < 32; i++) x_bytes[i] = root_bytes[i];
    ed::fe_tobytes(y_bytes, yi);
    return true;
}

// table entry (integer limbs, ed::ge_niels_slot) -> doubles: limb k = v[2k] + 2^26 v[2k+1], exact (|.| < 2^51)
BSX_HD ged_niels ged_niels_load(const ed::ge_niels_slot *slot) {
    const ed::ge_niels q = ed::ge_niels_load(slot);
    ged_niels r;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        r.ypx.v[k] = (double)q.ypx.v[2 * k] + 67108864.0 * (double)q.ypx.v[2 * k + 1];
        r.ymx.v[k] = (double)q.ymx.v[2 * k] + 67108864.0 * (double)q.ymx.v[2 * k + 1];
        r.xy2d.v[k] = (double)q.xy2d.v[2 * k] + 67108864.0 * (double)q.xy2d.v[2 * k + 1];
    }
    return r;
}
template <bool INL = false>
BSX_HD ged_p3 ged_scalarmult_base(const uint8_t s[32], const ed::ge_niels_slot *table) {
    ged_p3 acc = ged_identity();
#pragma unroll 1
    for (int w = 0; w < BSX_ED_BASE_WINDOWS; w++) {
        const uint32_t dgt = s[w];
        if (dgt) {
            const ged_niels q = ged_niels_load(table + w * BSX_ED_BASE_ENTRIES + (dgt - 1));
            acc = ged_p1p1_to_p3<INL>(ged_add_niels<INL>(acc, q), true);
        }
    }
    return acc;
}
BSX_HD ged_cached ged_cached_cneg(const ged_cached &q, bool neg) {
    ged_cached r;
    r.YpX = fed_select(neg, q.YmX, q.YpX); r.YmX = fed_select(neg, q.YpX, q.YmX); r.Z = q.Z;
    r.T2d = fed_select(neg, fed_neg(q.T2d), q.T2d);
    return r;
}
// scalar * P: signed radix-16 digits over the table 1P..8P, as ed::ge_scalarmult
template <bool INL = false>
BSX_HD ged_p3 ged_scalarmult(const uint8_t s[32], const ged_p3 &P) {
    ged_cached tab[8];
    {
        ged_p3 cur = P;
        tab[0] = ged_to_cached(cur);
#pragma unroll 1
        for (int d = 2; d <= 8; d++) {
            cur = ged_p1p1_to_p3(ged_add_cached(cur, tab[0]), true);
            tab[d - 1] = ged_to_cached(cur);
        }
    }
    int8_t e[65];
    {
        int carry = 0;
#pragma unroll 1
        for (int w = 0; w < 64; w++) {
            int d = (int)((s[w >> 1] >> ((w & 1) * 4)) & 15) + carry;
            carry = d >= 8;
            e[w] = (int8_t)(d - 16 * carry);
        }
        e[64] = (int8_t)carry;
    }
    ged_p3 acc = ged_identity();
    bool started = false;
#pragma unroll 1
    for (int w = 64; w >= 0; w--) {
        if (started) {
#pragma unroll 1
            for (int k = 0; k < 4; k++) acc = ged_p1p1_to_p3<INL>(ged_dbl<INL>(acc), k == 3);
        }
        const int d = e[w];
        if (d) {
            const ged_cached q = ged_cached_cneg(tab[(d < 0 ? -d : d) - 1], d < 0);
            acc = ged_p1p1_to_p3<INL>(ged_add_cached<INL>(acc, q), w == 0);
            started = true;
        }
    }
    return acc;
}

// one signature on the FP64 pipe: same record as ed::ed25519_witness_core.
// r02j: R is not decompressed (a square-root chain of ~265 field operations) when the signature verifies.  R' = sG - hA is
// formed from the two scalar multiplications and shares their inversion; if its canonical encoding IS the signature's R
// (y bytes equal, sign bit = parity of x'), then decompress(R) = R' with root = the even one of x', p - x', and
// R + hA = sG as points, so every byte of the record follows without the chain.  Anything else -- a signature that does not
// verify, a non-canonical y, an R off the curve -- takes the general path below (decompression, addition, a second
// inversion), which is the r02b code.  Same bytes either way (tests/test_ed25519_host_check.py, the GPU stress inputs).
// everything after the two scalar multiplications, in two halves around the inversion so that a thread that handles
// several signatures can share ONE inversion between them (Montgomery's trick):
//   ed25519_witness_points: R' = sG - hA and z = Z_sG Z_hA Z_R'
//   ed25519_witness_finish: given 1/z -- affine bytes, R by the shortcut or the general path, R + hA, flags
struct ed_pending { fed sgX, sgY, sgZ, haX, haY, haZ, haT, rpX, rpY, rpZ, z; };
BSX_HD ed_pending ed25519_witness_points(const ged_p3 &sg, const ged_p3 &ha) {
    const ged_p3 rp = ged_p1p1_to_p3(ged_add_cached(sg, ged_cached_cneg(ged_to_cached(ha), true)), false);
    ed_pending p;
    p.sgX = sg.X; p.sgY = sg.Y; p.sgZ = sg.Z; p.haX = ha.X; p.haY = ha.Y; p.haZ = ha.Z; p.haT = ha.T;
    p.rpX = rp.X; p.rpY = rp.Y; p.rpZ = rp.Z;
    p.z = fed_mul(fed_mul(sg.Z, ha.Z), rp.Z);
    return p;
}
BSX_HD void ed25519_witness_finish(const ed_pending &p, const fed &inv, const uint8_t sig[64], uint32_t flags, uint8_t *out) {
    const fed irp = fed_mul(inv, fed_mul(p.sgZ, p.haZ));
    const fed i12 = fed_mul(inv, p.rpZ);
    const fed isg = fed_mul(i12, p.haZ), iha = fed_mul(i12, p.sgZ);
    uint8_t sgb[64], t[64], rb[96];      // sG (x, y); scratch; R (x, y, root)
    ed::fe_tobytes(sgb, fe_from_fed(fed_mul(p.sgX, isg))); ed::fe_tobytes(sgb + 32, fe_from_fed(fed_mul(p.sgY, isg)));
    put_bytes<64>(out + 136, sgb);
    ed::fe_tobytes(t, fe_from_fed(fed_mul(p.haX, iha))); ed::fe_tobytes(t + 32, fe_from_fed(fed_mul(p.haY, iha)));
    put_bytes<64>(out + 296, t);
    const fed rpx = fed_mul(p.rpX, irp);
    ed::fe_tobytes(rb, fe_from_fed(rpx)); ed::fe_tobytes(rb + 32, fe_from_fed(fed_mul(p.rpY, irp)));
    const bool sign = (sig[31] >> 7) != 0;
    bool same = ((rb[0] & 1) != 0) == sign && rb[32 + 31] == (sig[31] & 0x7f);
    for (int i = 0; i < 31; i++) same = same && rb[32 + i] == sig[i];
    if (same) {
        flags |= 4u | 8u;
        if (sign) ed::fe_tobytes(rb + 64, fe_from_fed(fed_neg(rpx)));             // x' odd: the even root is p - x'
        else for (int i = 0; i < 32; i++) rb[64 + i] = rb[i];
        put_bytes<96>(out + 360, rb);
        put_bytes<64>(out + 456, sgb);                                              // R + hA = sG
    } else {
        fed rx, ry;
        if (ged_decompress(sig, rx, ry, rb, rb + 32, rb + 64)) flags |= 4u;
        put_bytes<96>(out + 360, rb);
        ged_p3 ha; ha.X = p.haX; ha.Y = p.haY; ha.Z = p.haZ; ha.T = p.haT;
        const ged_p3 sum = ged_p1p1_to_p3(ged_add_cached(ha, ged_to_cached(ged_from_affine(rx, ry))), false);
        const fed isum = fed_invert(sum.Z);
        ed::fe_tobytes(t, fe_from_fed(fed_mul(sum.X, isum))); ed::fe_tobytes(t + 32, fe_from_fed(fed_mul(sum.Y, isum)));
        put_bytes<64>(out + 456, t);
        if (ed::bytes_eq32(sgb, t) && ed::bytes_eq32(sgb + 32, t + 32)) flags |= 8u;
    }
    put_flags_and_padding(out, flags);
}
BSX_HD void ed25519_witness_tail(const ged_p3 &sg, const ged_p3 &ha, const uint8_t sig[64], uint32_t flags, uint8_t *out) {
    const ed_pending p = ed25519_witness_points(sg, ha);
    ed25519_witness_finish(p, fed_invert(p.z), sig, flags, out);
}

template <bool INL = false>
BSX_HD void ed25519_witness_core(const uint8_t pk[32], const uint8_t sig[64], const uint8_t digest[64],
                                 const ed::ge_niels_slot *base_table, uint8_t *out) {
    uint8_t h[32], div[40], ab[96];
    put_bytes<64>(out, digest);
    ed::sc_divrem_l(digest, h, div);
    put_bytes<32>(out + 64, h); put_bytes<40>(out + 96, div);
    uint32_t flags = ed::sc_lt_l(sig + 32) ? 1u : 0u;
    fed ax, ay;
    if (ged_decompress(pk, ax, ay, ab, ab + 32, ab + 64)) flags |= 2u;
    put_bytes<96>(out + 200, ab);
    const ged_p3 sg = ged_scalarmult_base<INL>(sig + 32, base_table);
    const ged_p3 ha = ged_scalarmult<INL>(h, ged_from_affine(ax, ay));
    ed25519_witness_tail(sg, ha, sig, flags, out);
}

// ---- per-key window tables (r02k) ----
// A batch that repeats public keys (one validator set signs every range of a batch) pays the 252 doublings of h*A and the
// decompression of A once per KEY instead of once per signature: for every distinct key the windows 16^w A (w = 0..63) and
// their multiples 1..2^(b-1) are tabulated in the projective addend form (b = BSX_ED_KEY_BITS = 6: 43 windows of 32 entries),
// and h*A becomes one table addition per window over the signed radix-2^b digits of h (h < l < 2^253: no carry out of the top).  The result is the same
// group element, hence the same affine bytes.  Layouts (doubles / bytes):
//   key record  BSX_ED_KEYREC_BYTES: x[32] y[32] root[32] ok[1] of decompress(A) -- exactly the bytes of the signature record
//   bases       [key][w]    one extended point (X, Y, Z, T) = 20 doubles
//   table       [key][w][d] addend form of (d+1) 2^(b w) A (YpX, YmX, Z, T2d) = 20 doubles, d < 2^(b-1)
#define BSX_ED_KEYREC_BYTES 128
#ifndef BSX_ED_KEY_BITS
#define BSX_ED_KEY_BITS 6                                            // window width: signed digits in [-2^(b-1), 2^(b-1)]
#endif
#define BSX_ED_KEY_WINDOWS ((253 + BSX_ED_KEY_BITS - 1) / BSX_ED_KEY_BITS)   // h < l < 2^253
#define BSX_ED_KEY_ENTRIES (1 << (BSX_ED_KEY_BITS - 1))             // multiples 1 .. 2^(b-1) of a window base
BSX_HD void ged_store20(double *dst, const fed &a, const fed &b, const fed &c, const fed &d) {
#pragma unroll
    for (int k = 0; k < 5; k++) { dst[k] = a.v[k]; dst[5 + k] = b.v[k]; dst[10 + k] = c.v[k]; dst[15 + k] = d.v[k]; }
}
BSX_HD void ged_load20(const double *src, fed &a, fed &b, fed &c, fed &d) {
    double t[20];
#if defined(__CUDA_ARCH__)
    const double2 *q = reinterpret_cast<const double2 *>(src);   // 160-byte entries of a 256-byte-aligned allocation
#pragma unroll
    for (int k = 0; k < 10; k++) { const double2 v = __ldg(q + k); t[2 * k] = v.x; t[2 * k + 1] = v.y; }
#else
    for (int k = 0; k < 20; k++) t[k] = src[k];
#endif
#pragma unroll
    for (int k = 0; k < 5; k++) { a.v[k] = t[k]; b.v[k] = t[5 + k]; c.v[k] = t[10 + k]; d.v[k] = t[15 + k]; }
}
// one key: decompress, then the 64 window bases (252 dependent doublings -- the part that is shared by every signature of the key)
BSX_HD void ed25519_key_bases(const uint8_t pk[32], uint8_t *rec, double *bases) {
    fed ax, ay;
    rec[96] = ged_decompress(pk, ax, ay, rec, rec + 32, rec + 64) ? 1 : 0;
    ged_p3 p = ged_from_affine(ax, ay);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int w = 0; w < BSX_ED_KEY_WINDOWS; w++) {
        ged_store20(bases + 20 * w, p.X, p.Y, p.Z, p.T);
        if (w + 1 < BSX_ED_KEY_WINDOWS) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int k = 0; k < BSX_ED_KEY_BITS; k++) p = ged_p1p1_to_p3(ged_dbl(p), k == BSX_ED_KEY_BITS - 1);
        }
    }
}
// one (key, window): the addend forms of 1..8 times the window base
BSX_HD void ed25519_key_window(const double *base, double *tab) {
    ged_p3 cur;
    ged_load20(base, cur.X, cur.Y, cur.Z, cur.T);
    const ged_cached first = ged_to_cached(cur);
    ged_store20(tab, first.YpX, first.YmX, first.Z, first.T2d);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int d = 2; d <= BSX_ED_KEY_ENTRIES; d++) {
        cur = ged_p1p1_to_p3(ged_add_cached(cur, first), true);
        const ged_cached c = ged_to_cached(cur);
        ged_store20(tab + 20 * (d - 1), c.YpX, c.YmX, c.Z, c.T2d);
    }
}
// h * A from the key's table: signed radix-16 digits as in ged_scalarmult, one table addition per non-zero digit
template <bool INL = false>
BSX_HD ged_p3 ged_scalarmult_keyed(const uint8_t s[32], const double *tab) {
    ged_p3 acc = ged_identity();
    int carry = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int w = 0; w < BSX_ED_KEY_WINDOWS; w++) {
        const int bit = w * BSX_ED_KEY_BITS, byte = bit >> 3;
        const uint32_t v = (uint32_t)s[byte] | ((uint32_t)(byte + 1 < 32 ? s[byte + 1] : 0) << 8);
        int d = (int)((v >> (bit & 7)) & ((1u << BSX_ED_KEY_BITS) - 1)) + carry;
        carry = d > BSX_ED_KEY_ENTRIES;
        d -= (1 << BSX_ED_KEY_BITS) * carry;
        if (d) {
            ged_cached q;
            ged_load20(tab + 20 * (BSX_ED_KEY_ENTRIES * w + (d < 0 ? -d : d) - 1), q.YpX, q.YmX, q.Z, q.T2d);
            acc = ged_p1p1_to_p3<INL>(ged_add_cached<INL>(acc, ged_cached_cneg(q, d < 0)), true);
        }
    }
    return acc;   // no carry out of the top window: the remainder mod l is < 2^253 and the windows cover >= 253 bits with room
}
// the record of one signature whose key has been tabulated: same bytes as ed25519_witness_core.  First half (up to the
// inversion) and the whole of it for one signature per thread.
template <bool INL = false>
BSX_HD ed_pending ed25519_witness_keyed_points(const uint8_t sig[64], const uint8_t digest[64], const ed::ge_niels_slot *base_table,
                                               const uint8_t *key_rec, const double *key_tab, uint8_t *out, uint32_t &flags) {
    uint8_t h[32], div[40];
    put_bytes<64>(out, digest);
    ed::sc_divrem_l(digest, h, div);
    put_bytes<32>(out + 64, h); put_bytes<40>(out + 96, div);
    flags = ed::sc_lt_l(sig + 32) ? 1u : 0u;
#if defined(__CUDA_ARCH__)
    {   // x, y, root of decompress(A): 96 bytes of the key record, word for word
        const uint2 *src = reinterpret_cast<const uint2 *>(key_rec);
        uint2 *dst = reinterpret_cast<uint2 *>(out + 200);
#pragma unroll
        for (int k = 0; k < 12; k++) dst[k] = __ldg(src + k);
    }
#else
    for (int i = 0; i < 96; i++) out[200 + i] = key_rec[i];
#endif
    if (key_rec[96]) flags |= 2u;
    const ged_p3 sg = ged_scalarmult_base<INL>(sig + 32, base_table);
    const ged_p3 ha = ged_scalarmult_keyed<INL>(h, key_tab);
    return ed25519_witness_points(sg, ha);
}
template <bool INL = false>
BSX_HD void ed25519_witness_core_keyed(const uint8_t sig[64], const uint8_t digest[64], const ed::ge_niels_slot *base_table,
                                       const uint8_t *key_rec, const double *key_tab, uint8_t *out) {
    uint32_t flags;
    const ed_pending p = ed25519_witness_keyed_points<INL>(sig, digest, base_table, key_rec, key_tab, out, flags);
    ed25519_witness_finish(p, fed_invert(p.z), sig, flags, out);
}

}  // namespace edd
}  // namespace bsx

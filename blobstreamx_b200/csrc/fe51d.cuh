// Field arithmetic mod p = 2^255 - 19 on the FP64 pipe of B200 (sm_100a): 5 balanced limbs of 51 bits held in doubles.
//
// Why: the 10 x 25.5-bit integer form (ed25519.cuh) is bound by IMAD.WIDE, which issues once per 4 cycles per SM
// sub-partition (profiles/r01j_ubench_femul.txt); B200 (unlike B300) has a full-rate FP64 pipe -- one DFMA per 2 cycles
// per sub-partition, a pipe of its own (profiles/r02a_ubench_fp64.txt).  A 51x51-bit product is split exactly by two
// round-toward-zero FMAs against a large constant (the fixed exponent turns the mantissa into an integer):
//     p_hi = fma_rz(a, b, C1)            C1 = 1.5 * 2^104      bits(p_hi) = bits(C1) + floor(a b / 2^52)
//     p_lo = fma_rz(a, b, C2 - p_hi)     C2 = C1 + 2^52        bits(p_lo) = bits(2^52) + (a b mod 2^52)
// valid for |a b| < 2^103.  The bit patterns are summed per column as 64-bit integers (the constants come off once per
// column), folded with 19 and carried in ONE parallel round (carries are < 2^10 against 51-bit limbs), then turned back
// into doubles with the 1.5 * 2^52 magic constant.  25 products x (2 DFMA + 1 DADD) per multiplication, 15 per squaring.
//
// Bounds: "carried" = |limb| <= 2^50 + 2^14.  fe_mul / fe_sq accept |f_i| * |g_j| < 2^103, i.e. the sum/difference of two
// carried values on each side, or three against two; every function here returns carried limbs.  Cross terms of a
// squaring are doubled in the integer domain, so fe_sq takes the same 2-unit inputs.
#pragma once
#include <stdint.h>
#include <string.h>
#if !defined(__CUDA_ARCH__)
#include <math.h>
#endif

#if defined(__CUDACC__)
#define BSX_D_HD __host__ __device__ __forceinline__
#define BSX_D_CALL static __host__ __device__ __noinline__
#else
#define BSX_D_HD static inline
#define BSX_D_CALL static __attribute__((noinline))
#endif

namespace bsx {
namespace edd {

struct fed { double v[5]; };

BSX_D_HD int64_t d2bits(double x) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(x);
#else
    int64_t r; memcpy(&r, &x, 8); return r;
#endif
}
BSX_D_HD double bits2d(int64_t x) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(x);
#else
    double r; memcpy(&r, &x, 8); return r;
#endif
}
// a*b + c rounded toward zero (host build: the caller runs under fesetround(FE_TOWARDZERO))
BSX_D_HD double fma_rz(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rz(a, b, c);
#else
    return fma(a, b, c);
#endif
}
// exact by construction (difference of two doubles of one binade / sum below 2^53): any rounding mode
BSX_D_HD double dsub(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
BSX_D_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

#define BSX_D_C1_BITS 0x4678000000000000LL   /* 1.5 * 2^104 */
#define BSX_D_C2_BITS 0x4678000000000001LL   /* 1.5 * 2^104 + 2^52 */
#define BSX_D_LO_BITS 0x4330000000000000LL   /* 2^52 */
#define BSX_D_MAGIC_BITS 0x4338000000000000LL /* 1.5 * 2^52 */

struct prod { uint64_t hi, lo; };   // bit patterns, constants still inside (summed modulo 2^64: unsigned)
BSX_D_HD prod dmul(double a, double b) {
    const double c1 = bits2d(BSX_D_C1_BITS), c2 = bits2d(BSX_D_C2_BITS);
    const double ph = fma_rz(a, b, c1);
    const double pl = fma_rz(a, b, dsub(c2, ph));
    prod r; r.hi = (uint64_t)d2bits(ph); r.lo = (uint64_t)d2bits(pl);
    return r;
}

BSX_D_HD fed fed_zero() { fed r; for (int i = 0; i < 5; i++) r.v[i] = 0.0; return r; }
BSX_D_HD fed fed_one() { fed r = fed_zero(); r.v[0] = 1.0; return r; }
BSX_D_HD fed fed_add(const fed &a, const fed &b) { fed r; for (int i = 0; i < 5; i++) r.v[i] = dadd(a.v[i], b.v[i]); return r; }
BSX_D_HD fed fed_sub(const fed &a, const fed &b) { fed r; for (int i = 0; i < 5; i++) r.v[i] = dsub(a.v[i], b.v[i]); return r; }
BSX_D_HD fed fed_neg(const fed &a) { fed r; for (int i = 0; i < 5; i++) r.v[i] = -a.v[i]; return r; }
BSX_D_HD fed fed_select(bool c, const fed &a, const fed &b) { fed r; for (int i = 0; i < 5; i++) r.v[i] = c ? a.v[i] : b.v[i]; return r; }

// E[k] (64-bit, weight 2^(51k), |E| < 2^62) -> carried doubles.  One parallel round: c_k = round(E_k / 2^51),
// limb_k = E_k - c_k 2^51 + c_{k-1} (19 c_4 into limb 0).
BSX_D_HD fed fed_carry(const int64_t E[5]) {
    int64_t c[5], l[5];
#pragma unroll
    for (int k = 0; k < 5; k++) { c[k] = (E[k] + ((int64_t)1 << 50)) >> 51; l[k] = E[k] - c[k] * ((int64_t)1 << 51); }
    fed r;
    const double magic = bits2d(BSX_D_MAGIC_BITS);
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const int64_t t = l[k] + (k == 0 ? 19 * c[4] : c[k - 1]);
        r.v[k] = dsub(bits2d(t + BSX_D_MAGIC_BITS), magic);
    }
    return r;
}

// column sums of the products -> E[5]
//   H[k], L[k]: sums of hi / lo bit patterns of column k (k = 0..8) with nh[k] terms each
BSX_D_HD void fed_fold(const uint64_t H[9], const uint64_t L[9], const int nterms[9], int64_t E[5]) {
    int64_t hs[9], ls[9];
#pragma unroll
    for (int k = 0; k < 9; k++) {
        hs[k] = (int64_t)(H[k] - (uint64_t)nterms[k] * (uint64_t)BSX_D_C1_BITS);
        ls[k] = (int64_t)(L[k] - (uint64_t)nterms[k] * (uint64_t)BSX_D_LO_BITS);
    }
    // value = sum_k (ls[k] + 2^52 hs[k]) 2^(51k);  2^52 2^(51k) = 2 * 2^(51(k+1));  2^(51*5) = 19
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const int64_t a = ls[k] + (k < 4 ? 19 * ls[k + 5] : 0);
        const int64_t b = (k == 0 ? 19 * hs[4] : hs[k - 1] + 19 * hs[k + 4]);
        E[k] = a + 2 * b;
    }
}

struct fed_raw { int64_t E[5]; };   // folded columns before the carry: weight 2^(51k), |E_k| < 2^59.7

BSX_D_HD fed_raw fed_mul_raw(const fed &f, const fed &g) {
    uint64_t H[9], L[9];
    const int nt[9] = {1, 2, 3, 4, 5, 4, 3, 2, 1};
#pragma unroll
    for (int k = 0; k < 9; k++) { H[k] = 0; L[k] = 0; }
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const prod p = dmul(f.v[i], g.v[j]);
            H[i + j] += p.hi; L[i + j] += p.lo;
        }
    fed_raw r;
    fed_fold(H, L, nt, r.E);
    return r;
}

// f*f: 15 products, cross terms doubled as integers (so that a 2-unit input stays within the product bound)
BSX_D_HD fed_raw fed_sq_raw(const fed &f) {
    uint64_t Hd[9], Ld[9], Hc[9], Lc[9];
#pragma unroll
    for (int k = 0; k < 9; k++) { Hd[k] = 0; Ld[k] = 0; Hc[k] = 0; Lc[k] = 0; }
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = i; j < 5; j++) {
            const prod p = dmul(f.v[i], f.v[j]);
            if (i == j) { Hd[i + j] += p.hi; Ld[i + j] += p.lo; }
            else { Hc[i + j] += p.hi; Lc[i + j] += p.lo; }
        }
    const int nd[9] = {1, 0, 1, 0, 1, 0, 1, 0, 1};
    const int nc[9] = {0, 1, 1, 2, 2, 2, 1, 1, 0};
    uint64_t H[9], L[9];
#pragma unroll
    for (int k = 0; k < 9; k++) {
        // the cross sums lose their constants before the doubling, so that one fold serves both kinds
        const uint64_t hc = Hc[k] - (uint64_t)nc[k] * (uint64_t)BSX_D_C1_BITS, lc = Lc[k] - (uint64_t)nc[k] * (uint64_t)BSX_D_LO_BITS;
        H[k] = Hd[k] + 2 * hc; L[k] = Ld[k] + 2 * lc;
    }
    fed_raw r;
    fed_fold(H, L, nd, r.E);
    return r;
}

BSX_D_HD fed fed_mul_inl(const fed &f, const fed &g) { const fed_raw r = fed_mul_raw(f, g); return fed_carry(r.E); }
BSX_D_HD fed fed_mul2_inl(const fed &f, const fed &g) {
    fed_raw r = fed_mul_raw(f, g);
#pragma unroll
    for (int k = 0; k < 5; k++) r.E[k] *= 2;
    return fed_carry(r.E);
}
template <bool TWICE>
BSX_D_HD fed fed_sq_impl(const fed &f) {
    fed_raw r = fed_sq_raw(f);
    if (TWICE) {
#pragma unroll
        for (int k = 0; k < 5; k++) r.E[k] *= 2;
    }
    return fed_carry(r.E);
}

BSX_D_CALL fed fed_mul(const fed f, const fed g) { return fed_mul_inl(f, g); }
BSX_D_CALL fed fed_mul2(const fed f, const fed g) { return fed_mul2_inl(f, g); }
BSX_D_CALL fed fed_sq(const fed f) { return fed_sq_impl<false>(f); }
BSX_D_CALL fed fed_sq2(const fed f) { return fed_sq_impl<true>(f); }
BSX_D_CALL fed_raw fed_sq_raw_call(const fed f) { return fed_sq_raw(f); }
template <bool INL> BSX_D_HD fed fed_mul_x(const fed &f, const fed &g) { if (INL) return fed_mul_inl(f, g); return fed_mul(f, g); }
template <bool INL> BSX_D_HD fed fed_mul2_x(const fed &f, const fed &g) { if (INL) return fed_mul2_inl(f, g); return fed_mul2(f, g); }
template <bool INL> BSX_D_HD fed fed_sq_x(const fed &f) { if (INL) return fed_sq_impl<false>(f); return fed_sq(f); }
template <bool INL> BSX_D_HD fed_raw fed_sq_raw_x(const fed &f) { if (INL) return fed_sq_raw(f); return fed_sq_raw_call(f); }
// n >= 1 successive squarings (the long chains of the inversion and the square roots)
BSX_D_CALL fed fed_sqn(fed f, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) f = fed_sq_impl<false>(f);
    return f;
}

// integer-valued limbs below 2^62 in magnitude -> carried (same value mod p)
BSX_D_HD fed fed_from_i64(const int64_t l[5]) { return fed_carry(l); }
// exact integer value of a limb (|v| < 2^62)
BSX_D_HD int64_t fed_limb_i64(double v) {
#if defined(__CUDA_ARCH__)
    return __double2ll_rn(v);
#else
    return (int64_t)v;
#endif
}
// re-carry a sum/difference of carried values (same value mod p)
BSX_D_HD fed fed_reduce(const fed &f) {
    int64_t E[5];
#pragma unroll
    for (int k = 0; k < 5; k++) E[k] = fed_limb_i64(f.v[k]);
    return fed_carry(E);
}

// z^(2^252 - 3) = z^((p-5)/8)
BSX_D_HD fed fed_pow22523(const fed &z) {
    fed t0 = fed_sq(z);
    fed t1 = fed_sqn(t0, 2);
    t1 = fed_mul(z, t1);
    t0 = fed_mul(t0, t1);
    t0 = fed_sq(t0);
    t0 = fed_mul(t1, t0);
    t1 = fed_sqn(t0, 5);
    t0 = fed_mul(t1, t0);
    t1 = fed_sqn(t0, 10);
    t1 = fed_mul(t1, t0);
    fed t2 = fed_sqn(t1, 20);
    t1 = fed_mul(t2, t1);
    t1 = fed_sqn(t1, 10);
    t0 = fed_mul(t1, t0);
    t1 = fed_sqn(t0, 50);
    t1 = fed_mul(t1, t0);
    t2 = fed_sqn(t1, 100);
    t1 = fed_mul(t2, t1);
    t1 = fed_sqn(t1, 50);
    t0 = fed_mul(t1, t0);
    t0 = fed_sqn(t0, 2);
    return fed_mul(t0, z);
}

// z^(p-2)
BSX_D_HD fed fed_invert(const fed &z) {
    fed t0 = fed_sq(z);
    fed t1 = fed_sqn(t0, 2);
    t1 = fed_mul(z, t1);
    t0 = fed_mul(t0, t1);
    fed t2 = fed_sq(t0);
    t1 = fed_mul(t1, t2);
    t2 = fed_sqn(t1, 5);
    t1 = fed_mul(t2, t1);
    t2 = fed_sqn(t1, 10);
    t2 = fed_mul(t2, t1);
    fed t3 = fed_sqn(t2, 20);
    t2 = fed_mul(t3, t2);
    t2 = fed_sqn(t2, 10);
    t1 = fed_mul(t2, t1);
    t2 = fed_sqn(t1, 50);
    t2 = fed_mul(t2, t1);
    t3 = fed_sqn(t2, 100);
    t2 = fed_mul(t3, t2);
    t2 = fed_sqn(t2, 50);
    t1 = fed_mul(t2, t1);
    t1 = fed_sqn(t1, 5);
    return fed_mul(t1, t0);
}

}  // namespace edd
}  // namespace bsx

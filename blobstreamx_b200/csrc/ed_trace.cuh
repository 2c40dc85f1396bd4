// Ed25519 scalar-multiplication trace: the per-multiplication chain and the per-row expansion (see k_trace_ed.cu for
// the layout and the reference citations).  Host + device like ed25519.cuh, so tests/host_check can run the very code
// the kernels run against oracle/ed_trace.py on a machine without a GPU.
#pragma once
#include "ed25519.cuh"

namespace bsx {
namespace edt {

using namespace ed;

#define EDT_OP 92                 // result 16, carry 16, witness_low 30, witness_high 30
#define EDT_OFFSET (1 << 22)
#define EDT_CHAIN_WORDS 80        // per row: sum and dbl, each X Y Z + the running product before it (4 x 10 limbs)
#define EDT_AFF_WORDS 32          // per row: sum.x sum.y dbl.x dbl.y, 8 little-endian words each

#if defined(__CUDA_ARCH__)
#define EDT_ST(ptr, val) __stcs((ptr), (val))
#define EDT_CLZ(x) __clz(x)
#else
#define EDT_ST(ptr, val) (*(ptr) = (val))
#define EDT_CLZ(x) __builtin_clz(x)
#endif

BSX_HD void st_fe4(int32_t *dst, const fe &a, const fe &b, const fe &c, const fe &d) {
    int32_t w[40];
#pragma unroll
    for (int i = 0; i < 10; i++) { w[i] = a.v[i]; w[10 + i] = b.v[i]; w[20 + i] = c.v[i]; w[30 + i] = d.v[i]; }
#if defined(__CUDA_ARCH__)
    int4 *o = reinterpret_cast<int4 *>(dst);
#pragma unroll
    for (int i = 0; i < 10; i++) o[i] = make_int4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
#else
    for (int i = 0; i < 40; i++) dst[i] = w[i];
#endif
}
BSX_HD void ld_fe4(const int32_t *src, fe &a, fe &b, fe &c, fe &d) {
    int32_t w[40];
#if defined(__CUDA_ARCH__)
    const int4 *in = reinterpret_cast<const int4 *>(src);
#pragma unroll
    for (int i = 0; i < 10; i++) { const int4 q = in[i]; w[4 * i] = q.x; w[4 * i + 1] = q.y; w[4 * i + 2] = q.z; w[4 * i + 3] = q.w; }
#else
    for (int i = 0; i < 40; i++) w[i] = src[i];
#endif
#pragma unroll
    for (int i = 0; i < 10; i++) { a.v[i] = w[i]; b.v[i] = w[10 + i]; c.v[i] = w[20 + i]; d.v[i] = w[30 + i]; }
}
BSX_HD uint32_t ld_le32(const uint8_t *s) { return (uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16) | ((uint32_t)s[3] << 24); }
BSX_HD void st_affine(uint32_t *dst, const fe &x, const fe &y) {
    uint8_t b[64];
    fe_tobytes(b, x); fe_tobytes(b + 32, y);
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = ld_le32(b + 4 * i);
#if defined(__CUDA_ARCH__)
    uint4 *o = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int i = 0; i < 4; i++) o[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
#else
    for (int i = 0; i < 16; i++) dst[i] = w[i];
#endif
}

// One multiplication k * P.  ch: 256 * EDT_CHAIN_WORDS words of working values; af: 256 * EDT_AFF_WORDS words, the affine
// sum and dbl of every step; result (may be null): 64 bytes k * P.
BSX_HD void edt_chain_core(const uint8_t *scalar, const uint8_t *point, int32_t *ch, uint32_t *af, uint8_t *result) {
    uint32_t k[8];
#pragma unroll
    for (int i = 0; i < 8; i++) k[i] = ld_le32(scalar + 4 * i);
    const fe px = fe_frombytes(point), py = fe_frombytes(point + 32);
    ge_p3 temp = ge_from_affine(px, py), acc = ge_identity();
    fe run = fe_one();
#pragma unroll 1
    for (int j = 0; j < 256; j++) {
        const ge_p3 sum = ge_p1p1_to_p3(ge_add_cached(acc, ge_to_cached(temp)), true);
        const ge_p3 dbl = ge_p1p1_to_p3(ge_dbl(temp), true);
        st_fe4(ch + j * EDT_CHAIN_WORDS, sum.X, sum.Y, sum.Z, run);
        run = fe_mul(run, sum.Z);
        st_fe4(ch + j * EDT_CHAIN_WORDS + 40, dbl.X, dbl.Y, dbl.Z, run);
        run = fe_mul(run, dbl.Z);
        const bool bit = (k[j >> 5] >> (j & 31)) & 1;
        acc.X = fe_select(bit, sum.X, acc.X); acc.Y = fe_select(bit, sum.Y, acc.Y);
        acc.Z = fe_select(bit, sum.Z, acc.Z); acc.T = fe_select(bit, sum.T, acc.T);
        temp = dbl;
    }
    fe inv = fe_invert(run);
#pragma unroll 1
    for (int j = 255; j >= 0; j--) {
        fe X, Y, Z, before;
        ld_fe4(ch + j * EDT_CHAIN_WORDS + 40, X, Y, Z, before);
        fe zi = fe_mul(inv, before);
        inv = fe_mul(inv, Z);
        st_affine(af + j * EDT_AFF_WORDS + 16, fe_mul(X, zi), fe_mul(Y, zi));
        ld_fe4(ch + j * EDT_CHAIN_WORDS, X, Y, Z, before);
        zi = fe_mul(inv, before);
        inv = fe_mul(inv, Z);
        st_affine(af + j * EDT_AFF_WORDS, fe_mul(X, zi), fe_mul(Y, zi));
    }
    if (result) {
        int top = -1;
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (k[i]) top = 32 * i + 31 - EDT_CLZ(k[i]);
        for (int i = 0; i < 16; i++) {
            const uint32_t w = top < 0 ? (i == 8 ? 1u : 0u) : af[top * EDT_AFF_WORDS + i];
            for (int b = 0; b < 4; b++) result[4 * i + b] = (uint8_t)(w >> (8 * b));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// row expansion: integer arithmetic on 16-bit limbs
// ---------------------------------------------------------------------------------------------------------------------
// V[i + j] += sign * a[i] * b[j]
template <int SIGN>
BSX_CALL void edt_pmac(int64_t *V, const uint32_t *a, const uint32_t *b) {
    int64_t acc[31];
    uint32_t x[16], y[16];
#pragma unroll
    for (int i = 0; i < 31; i++) acc[i] = V[i];
#pragma unroll
    for (int i = 0; i < 16; i++) { x[i] = a[i]; y[i] = b[i]; }
#pragma unroll
    for (int i = 0; i < 16; i++)
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int64_t p = (int64_t)((uint64_t)x[i] * y[j]);
            acc[i + j] += SIGN > 0 ? p : -p;
        }
#pragma unroll
    for (int i = 0; i < 31; i++) V[i] = acc[i];
}

// N = V(2^16) >= 0 (below 2^512) -> quotient q (8 words, < 2^256) and remainder r (8 words, < p) by p = 2^255 - 19.
// N = q p + r  <=>  N + 19 q = q 2^255 + r: iterate q <- (N + 19 q) >> 255 from q = 0.  The first pass gives N >> 255, at
// most 39 below the quotient; the second lands on it or one below (one below when r < 19 (q* - q), in particular for every
// exact division); the third pass only forms T = N + 19 q, whose top part tells which, and whose low 255 bits are r or r + p - 2^255.
BSX_CALL void edt_divmod_p(const int64_t *V, uint32_t *q, uint32_t *r) {
    uint32_t N[16];
    int64_t c = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        int64_t t = V[2 * k] + c;
        const uint32_t lo = (uint32_t)t & 0xffff;
        c = t >> 16;
        t = (2 * k + 1 < 31 ? V[2 * k + 1] : 0) + c;
        N[k] = lo | (((uint32_t)t & 0xffff) << 16);
        c = t >> 16;
    }
    uint32_t qq[9], T[17];
#pragma unroll
    for (int i = 0; i < 9; i++) qq[i] = 0;
#pragma unroll 1
    for (int pass = 0; pass < 3; pass++) {
        uint64_t cy = 0;
#pragma unroll
        for (int i = 0; i < 17; i++) {
            const uint64_t t = (i < 16 ? (uint64_t)N[i] : 0) + (i < 9 ? 19ull * qq[i] : 0) + cy;
            T[i] = (uint32_t)t;
            cy = t >> 32;
        }
        if (pass < 2) {
#pragma unroll
            for (int i = 0; i < 9; i++) qq[i] = (T[i + 7] >> 31) | (i + 8 < 17 ? (T[i + 8] << 1) : 0);
        }
    }
    bool bump = false;                                    // T >> 255 != q: the quotient is q + 1
#pragma unroll
    for (int i = 0; i < 9; i++) bump |= (((T[i + 7] >> 31) | (i + 8 < 17 ? (T[i + 8] << 1) : 0)) != qq[i]);
    uint32_t lo[8], lo19[8];
#pragma unroll
    for (int i = 0; i < 8; i++) lo[i] = T[i];
    lo[7] &= 0x7fffffffu;
    uint64_t cy = 19;
#pragma unroll
    for (int i = 0; i < 8; i++) { const uint64_t t = (uint64_t)lo[i] + cy; lo19[i] = (uint32_t)t; cy = t >> 32; }
    bump |= (lo19[7] >> 31) != 0;                         // low255(T) >= p
    lo19[7] &= 0x7fffffffu;
    uint32_t inc = bump ? 1u : 0u;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r[i] = bump ? lo19[i] : lo[i];
        const uint64_t t = (uint64_t)qq[i] + inc;
        q[i] = (uint32_t)t;
        inc = (uint32_t)(t >> 32);
    }
}

BSX_HD void edt_limbs(uint32_t *l, const uint32_t *w) {
#pragma unroll
    for (int i = 0; i < 8; i++) { l[2 * i] = w[i] & 0xffff; l[2 * i + 1] = w[i] >> 16; }
}

// One field operation.  V = the left-hand-side polynomial (for a division, b res +- (res - a): the result is part of it).
// Writes result / carry / witness columns at col (column stride = n_rows) and returns the result limbs in res (mul-type).
template <bool DEN>
BSX_CALL void edt_emit(int64_t *V, uint32_t *res, uint64_t *col, size_t n_rows) {
    const uint32_t P16[16] = {0xffed, 0xffff, 0xffff, 0xffff, 0xffff, 0xffff, 0xffff, 0xffff,
                              0xffff, 0xffff, 0xffff, 0xffff, 0xffff, 0xffff, 0xffff, 0x7fff};
    uint32_t q[8], r[8], cl[16];
    edt_divmod_p(V, q, r);
    edt_limbs(cl, q);
    if (!DEN) {
        edt_limbs(res, r);
#pragma unroll
        for (int k = 0; k < 16; k++) V[k] -= res[k];
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {
        EDT_ST(col + (size_t)k * n_rows, (uint64_t)res[k]);
        EDT_ST(col + (size_t)(16 + k) * n_rows, (uint64_t)cl[k]);
    }
    edt_pmac<-1>(V, cl, P16);
    int64_t prev = 0;
#pragma unroll
    for (int k = 0; k < 30; k++) {
        prev = (prev - V[k]) >> 16;                       // exact
        const uint32_t sh = (uint32_t)(prev + EDT_OFFSET);
        EDT_ST(col + (size_t)(32 + k) * n_rows, (uint64_t)(sh & 0xffff));
        EDT_ST(col + (size_t)(62 + k) * n_rows, (uint64_t)(sh >> 16));
    }
}

BSX_HD void edt_zero(int64_t *V) {
#pragma unroll
    for (int i = 0; i < 31; i++) V[i] = 0;
}

// the eight operations of (x1, y1) + (x2, y2) = (x3, y3), x3 / y3 known (the chain's affine values)
BSX_CALL void edt_add_emit(const uint32_t *x1, const uint32_t *y1, const uint32_t *x2, const uint32_t *y2, uint32_t *x3, uint32_t *y3,
                           uint64_t *col, size_t n_rows) {
    // d = -121665 / 121666 mod p
    const uint32_t D16[16] = {0x78a3, 0x1359, 0x4dca, 0x75eb, 0xd8ab, 0x4141, 0x0a4d, 0x0070,
                              0xe898, 0x7779, 0x4079, 0x8cc7, 0xfe73, 0x2b6f, 0x6cee, 0x5203};
    int64_t V[31];
    uint32_t xn[16], yn[16], m1[16], m2[16], f[16], df[16];
    const size_t op = (size_t)EDT_OP * n_rows;
    edt_zero(V); edt_pmac<1>(V, x1, y2); edt_pmac<1>(V, x2, y1); edt_emit<false>(V, xn, col, n_rows);
    edt_zero(V); edt_pmac<1>(V, y1, y2); edt_pmac<1>(V, x1, x2); edt_emit<false>(V, yn, col + op, n_rows);
    edt_zero(V); edt_pmac<1>(V, x1, y1); edt_emit<false>(V, m1, col + 2 * op, n_rows);
    edt_zero(V); edt_pmac<1>(V, x2, y2); edt_emit<false>(V, m2, col + 3 * op, n_rows);
    edt_zero(V); edt_pmac<1>(V, m1, m2); edt_emit<false>(V, f, col + 4 * op, n_rows);
    edt_zero(V); edt_pmac<1>(V, D16, f); edt_emit<false>(V, df, col + 5 * op, n_rows);
    edt_zero(V); edt_pmac<1>(V, df, x3);
#pragma unroll
    for (int k = 0; k < 16; k++) V[k] += (int64_t)x3[k] - (int64_t)xn[k];
    edt_emit<true>(V, x3, col + 6 * op, n_rows);
    edt_zero(V); edt_pmac<1>(V, df, y3);
#pragma unroll
    for (int k = 0; k < 16; k++) V[k] += (int64_t)yn[k] - (int64_t)y3[k];
    edt_emit<true>(V, y3, col + 7 * op, n_rows);
}

BSX_HD void edt_ld_point(uint32_t *x, uint32_t *y, const uint32_t *src) {
    uint32_t w[16];
#if defined(__CUDA_ARCH__)
    const uint4 *in = reinterpret_cast<const uint4 *>(src);
#pragma unroll
    for (int i = 0; i < 4; i++) { const uint4 v = in[i]; w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
#else
    for (int i = 0; i < 16; i++) w[i] = src[i];
#endif
    edt_limbs(x, w); edt_limbs(y, w + 8);
}
BSX_HD void edt_identity(uint32_t *x, uint32_t *y) {
#pragma unroll
    for (int i = 0; i < 16; i++) { x[i] = 0; y[i] = 0; }
    y[0] = 1;
}

// Row j of one multiplication (real) or a padding row: scalar / point / af belong to the multiplication (unused for padding);
// col = &trace[row], column stride n_rows.
BSX_HD void edt_row_core(bool real, uint32_t j, const uint8_t *scalar, const uint8_t *point, const uint32_t *af, uint64_t *col, size_t n_rows) {
    uint32_t tx[16], ty[16], ax[16], ay[16], sx[16], sy[16], dx[16], dy[16];
    uint32_t bit = 0;
    edt_identity(tx, ty); edt_identity(ax, ay); edt_identity(sx, sy); edt_identity(dx, dy);
    if (real) {
        uint32_t k[8];
#pragma unroll
        for (int i = 0; i < 8; i++) k[i] = ld_le32(scalar + 4 * i);
        bit = (k[j >> 5] >> (j & 31)) & 1;
        int top = -1;                                        // last set bit below j: acc = the sum of that row
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t w = k[i];
            if ((uint32_t)(32 * i) >= j) w = 0;
            else if (j - 32 * i < 32) w &= (1u << (j - 32 * i)) - 1;
            if (w) top = 32 * i + 31 - EDT_CLZ(w);
        }
        edt_ld_point(sx, sy, af + j * EDT_AFF_WORDS);
        edt_ld_point(dx, dy, af + j * EDT_AFF_WORDS + 16);
        if (j) edt_ld_point(tx, ty, af + (j - 1) * EDT_AFF_WORDS + 16);
        else {
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 16; i++) w[i] = ld_le32(point + 4 * i);
            edt_limbs(tx, w); edt_limbs(ty, w + 8);
        }
        if (top >= 0) edt_ld_point(ax, ay, af + top * EDT_AFF_WORDS);
    }
    EDT_ST(col, (uint64_t)bit);
    EDT_ST(col + n_rows, (uint64_t)(real ? 1 : 0));
    EDT_ST(col + 2 * n_rows, (uint64_t)(j == 0));
    EDT_ST(col + 3 * n_rows, (uint64_t)(j == 255));
#pragma unroll
    for (int i = 0; i < 16; i++) {
        EDT_ST(col + (size_t)(4 + i) * n_rows, (uint64_t)tx[i]);
        EDT_ST(col + (size_t)(20 + i) * n_rows, (uint64_t)ty[i]);
        EDT_ST(col + (size_t)(36 + i) * n_rows, (uint64_t)ax[i]);
        EDT_ST(col + (size_t)(52 + i) * n_rows, (uint64_t)ay[i]);
    }
    edt_add_emit(ax, ay, tx, ty, sx, sy, col + (size_t)68 * n_rows, n_rows);
    edt_add_emit(tx, ty, tx, ty, dx, dy, col + (size_t)(68 + 8 * EDT_OP) * n_rows, n_rows);
}

}  // namespace edt
}  // namespace bsx

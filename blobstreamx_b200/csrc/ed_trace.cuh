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
#define EDT_CHAIN_WORDS 64        // per row: sum and dbl in extended coordinates, X Y Z (3 x 10 limbs) padded to 32 words each
#define EDT_AFF_WORDS 32          // per row: sum.x sum.y dbl.x dbl.y, 8 little-endian words each

// Row kernel output path.  Default: every thread stores its own 8 bytes (a warp = 256 B of a column).  EDT_STAGE = 1 (A/B
// build, measured and NOT kept): a CTA's 128 rows of an operation's 92 columns are staged in shared memory ([column][row])
// and leave as TMA bulk stores of 1 KB per column (cp.async.bulk shared -> global), the tile in two halves with their own
// issuing threads so that a half is only waited for one operation later -- 2.98 ms against 2.43 ms for 2^20 rows: the four
// barriers per operation put the CTA's warps in lockstep and the kernel is not short of store bandwidth but of independent
// work (without any global store the staged form takes 1.47 ms, profiles/r04k_ed_trace_stage.txt).
#ifndef EDT_STAGE
#define EDT_STAGE 0
#endif
#define EDT_TILE 128              // rows per CTA of the row kernel = its threads
#if defined(__CUDA_ARCH__)
#if EDT_STAGE
#define EDT_ST(ptr, val) (*(ptr) = (val))
#else
#define EDT_ST(ptr, val) __stcs((ptr), (val))
#endif
#define EDT_CLZ(x) __clz(x)
#else
#define EDT_ST(ptr, val) (*(ptr) = (val))
#define EDT_CLZ(x) __builtin_clz(x)
#endif

// Staged output (EDT_STAGE): the tile [92 columns][128 rows] has two halves with their own issuing threads, so that a half
// is only waited for when it is written again, one whole operation after its bulk stores were issued:
//   A = columns 0..31 (result, carry), issued by threads 64..95;  B = columns 32..91 (the witness halves), by threads 0..59.
#if defined(__CUDA_ARCH__) && EDT_STAGE
#define EDT_BULK(gdst, ssrc) \
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), \
                 "r"((uint32_t)(EDT_TILE * 8)) : "memory")
#define EDT_BULK_COMMIT() asm volatile("cp.async.bulk.commit_group;" ::: "memory")
#define EDT_BULK_WAIT() asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory")
#define EDT_IS_A (threadIdx.x >= 64 && threadIdx.x < 96)
#define EDT_IS_B (threadIdx.x < 60)
#define EDT_PRE(is)  do { if (is) EDT_BULK_WAIT(); __syncthreads(); } while (0)
#define EDT_POST(is, c0, stage, gcol, grows)                                                        \
    do {                                                                                            \
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                                \
        __syncthreads();                                                                            \
        if (is) {                                                                                   \
            const uint32_t c_ = (c0) + (threadIdx.x & 63);                                          \
            if (EDT_STAGE != 2) EDT_BULK((gcol) + (size_t)c_ * (grows), (stage) + (size_t)c_ * EDT_TILE); \
            EDT_BULK_COMMIT();                                                                      \
        }                                                                                           \
    } while (0)
#else
#define EDT_PRE(is) do { } while (0)
#define EDT_POST(is, c0, stage, gcol, grows) do { } while (0)
#endif

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

BSX_HD void st_fe3(int32_t *dst, const fe &a, const fe &b, const fe &c) {
    int32_t w[32];
#pragma unroll
    for (int i = 0; i < 10; i++) { w[i] = a.v[i]; w[10 + i] = b.v[i]; w[20 + i] = c.v[i]; }
    w[30] = 0; w[31] = 0;
#if defined(__CUDA_ARCH__)
    int4 *o = reinterpret_cast<int4 *>(dst);
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = make_int4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
#else
    for (int i = 0; i < 32; i++) dst[i] = w[i];
#endif
}
BSX_HD void ld_fe3(const int32_t *src, fe &a, fe &b, fe &c) {
    int32_t w[32];
#if defined(__CUDA_ARCH__)
    const int4 *in = reinterpret_cast<const int4 *>(src);
#pragma unroll
    for (int i = 0; i < 8; i++) { const int4 q = in[i]; w[4 * i] = q.x; w[4 * i + 1] = q.y; w[4 * i + 2] = q.z; w[4 * i + 3] = q.w; }
#else
    for (int i = 0; i < 32; i++) w[i] = src[i];
#endif
#pragma unroll
    for (int i = 0; i < 10; i++) { a.v[i] = w[i]; b.v[i] = w[10 + i]; c.v[i] = w[20 + i]; }
}

// Forward pass of one multiplication k * P in extended coordinates: at step j, sum = acc + temp and dbl = 2 temp (both,
// whatever the bit: the row witnesses both) go to ch[j * EDT_CHAIN_WORDS] (sum: X Y Z at words 0..29, dbl: at 32..61).
template <bool INL>
BSX_HD void edt_forward_core(const uint8_t *scalar, const uint8_t *point, int32_t *ch) {
    uint32_t k[8];
#pragma unroll
    for (int i = 0; i < 8; i++) k[i] = ld_le32(scalar + 4 * i);
    const fe px = fe_frombytes(point), py = fe_frombytes(point + 32);
    ge_p3 temp = ge_from_affine(px, py), acc = ge_identity();
#pragma unroll 1
    for (int j = 0; j < 256; j++) {
        const ge_p3 sum = ge_p1p1_to_p3<INL>(ge_add_cached<INL>(acc, ge_to_cached(temp)), true);
        const ge_p3 dbl = ge_p1p1_to_p3<INL>(ge_dbl<INL>(temp), true);
        st_fe3(ch + j * EDT_CHAIN_WORDS, sum.X, sum.Y, sum.Z);
        st_fe3(ch + j * EDT_CHAIN_WORDS + 32, dbl.X, dbl.Y, dbl.Z);
        const bool bit = (k[j >> 5] >> (j & 31)) & 1;
        acc.X = fe_select(bit, sum.X, acc.X); acc.Y = fe_select(bit, sum.Y, acc.Y);
        acc.Z = fe_select(bit, sum.Z, acc.Z); acc.T = fe_select(bit, sum.T, acc.T);
        temp = dbl;
    }
}

// EDT_GROUP consecutive steps of one multiplication -> affine: the 2 * EDT_GROUP Z's are inverted with ONE field inversion
// (prefix products), so a multiplication's 512 inversions cost 16 (one per group, groups in parallel).
// ch / af point at the group's first step.
#define EDT_GROUP 16
BSX_HD void edt_affine_core(const int32_t *ch, uint32_t *af) {
    fe pre[2 * EDT_GROUP];
    fe run = fe_one();
#pragma unroll 1
    for (int i = 0; i < 2 * EDT_GROUP; i++) {
        fe X, Y, Z;
        ld_fe3(ch + i * 32, X, Y, Z);
        pre[i] = run;
        run = fe_mul(run, Z);
    }
    fe inv = fe_invert(run);
#pragma unroll 1
    for (int i = 2 * EDT_GROUP - 1; i >= 0; i--) {
        fe X, Y, Z;
        ld_fe3(ch + i * 32, X, Y, Z);
        const fe zi = fe_mul(inv, pre[i]);
        inv = fe_mul(inv, Z);
        st_affine(af + i * 16, fe_mul(X, zi), fe_mul(Y, zi));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// row expansion: integer arithmetic on 16-bit limbs
// ---------------------------------------------------------------------------------------------------------------------
BSX_HD void edt_limbs(uint32_t *l, const uint32_t *w) {
#pragma unroll
    for (int i = 0; i < 8; i++) { l[2 * i] = w[i] & 0xffff; l[2 * i + 1] = w[i] >> 16; }
}

// One witnessed field operation, everything in registers from the limb products to the stores.
//   KIND 0  mul      lhs = a1 b1                    res <- lhs mod p
//   KIND 1  inner    lhs = a1 b1 + a2 b2            res <- lhs mod p
//   KIND 2  den (+)  lhs = a1 res + res - a2 = carry p   (res = a2 / (1 + a1), given)
//   KIND 3  den (-)  lhs = a1 res + a2 - res = carry p   (res = a2 / (1 - a1), given)
// Columns at col (n_rows = distance between columns there: the table's, or the staging tile's): result[16] carry[16]
// witness_low[30] witness_high[30], where
// lhs(x) - result(x) [KIND 0, 1] - carry(x) p(x) = (x - 2^16) w(x) and the stored witness is w_k + EDT_OFFSET.
//
// Quotient by p = 2^255 - 19:  N = q p + r  <=>  N + 19 q = q 2^255 + r.  Iterate q <- (N + 19 q) >> 255 from q = 0: the first
// pass gives N >> 255, at most 39 below the quotient; the second lands on it or one below (one below when r < 19 (q* - q), in
// particular for every exact division); the third pass only forms T = N + 19 q, whose top part tells which, and whose low
// 255 bits are r or r + p - 2^255.
// EDT_UNIFIED (default): ONE copy of the operation's code with the kind as a run-time (warp-uniform) argument -- the four
// template instances are 7.5 k instructions = 120 KB of code for the row kernel, and ncu shows 0.9 instruction-fetch stall
// cycles per issue; a doubling's x1 y2 + x2 y1 (same operands twice) is one product, doubled.
#ifndef EDT_UNIFIED
#define EDT_UNIFIED 1
#endif
#if EDT_UNIFIED
BSX_CALL void edt_op_rt(int KIND, const uint32_t *a1, const uint32_t *b1, const uint32_t *a2, const uint32_t *b2, uint32_t *res, uint64_t *col,
                        size_t n_rows, uint64_t *gcol, size_t grows) {
    int64_t V[31];
#pragma unroll
    for (int i = 0; i < 31; i++) V[i] = 0;
    const bool twice = KIND == 1 && a1 == a2 && b1 == b2;        // the doubling's inner product: 2 (x y)
    const int n_prod = (KIND == 1 && !twice) ? 2 : 1;
#pragma unroll 1
    for (int p = 0; p < n_prod; p++) {
        const uint32_t *xa = p ? a2 : a1, *yb = p ? b2 : (KIND >= 2 ? res : b1);
        uint32_t x[16], y[16];
#pragma unroll
        for (int i = 0; i < 16; i++) { x[i] = xa[i]; y[i] = yb[i]; }
#pragma unroll
        for (int i = 0; i < 16; i++)
#pragma unroll
            for (int j = 0; j < 16; j++) V[i + j] += (int64_t)((uint64_t)x[i] * y[j]);
    }
    if (twice) {
#pragma unroll
        for (int i = 0; i < 31; i++) V[i] += V[i];
    }
    if (KIND >= 2) {                                             // res (1 + a1) - a2  resp.  a2 - res (1 - a1)
        const int64_t sg = KIND == 2 ? 1 : -1;
#pragma unroll
        for (int k = 0; k < 16; k++) V[k] += sg * ((int64_t)res[k] - (int64_t)a2[k]);
    }
#else
template <int KIND>
BSX_CALL void edt_op(const uint32_t *a1, const uint32_t *b1, const uint32_t *a2, const uint32_t *b2, uint32_t *res, uint64_t *col, size_t n_rows,
                     uint64_t *gcol, size_t grows) {
    int64_t V[31];
#pragma unroll
    for (int i = 0; i < 31; i++) V[i] = 0;
    {
        uint32_t x[16], y[16];
#pragma unroll
        for (int i = 0; i < 16; i++) { x[i] = a1[i]; y[i] = KIND >= 2 ? res[i] : b1[i]; }
#pragma unroll
        for (int i = 0; i < 16; i++)
#pragma unroll
            for (int j = 0; j < 16; j++) V[i + j] += (int64_t)((uint64_t)x[i] * y[j]);
        if (KIND == 2) {
#pragma unroll
            for (int k = 0; k < 16; k++) V[k] += (int64_t)y[k] - (int64_t)a2[k];
        }
        if (KIND == 3) {
#pragma unroll
            for (int k = 0; k < 16; k++) V[k] += (int64_t)a2[k] - (int64_t)y[k];
        }
    }
    if (KIND == 1) {
        uint32_t x[16], y[16];
#pragma unroll
        for (int i = 0; i < 16; i++) { x[i] = a2[i]; y[i] = b2[i]; }
#pragma unroll
        for (int i = 0; i < 16; i++)
#pragma unroll
            for (int j = 0; j < 16; j++) V[i + j] += (int64_t)((uint64_t)x[i] * y[j]);
    }
#endif
    // N = V(2^16) >= 0, below 2^512
    uint32_t N[16];
    {
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
    uint32_t lo19[8];
    T[7] &= 0x7fffffffu;
    {
        uint64_t cy = 19;
#pragma unroll
        for (int i = 0; i < 8; i++) { const uint64_t t = (uint64_t)T[i] + cy; lo19[i] = (uint32_t)t; cy = t >> 32; }
    }
    bump |= (lo19[7] >> 31) != 0;                         // low255(T) >= p
    lo19[7] &= 0x7fffffffu;
    uint32_t cl[16];
    {
        uint32_t inc = bump ? 1u : 0u;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint64_t t = (uint64_t)qq[i] + inc;
            cl[2 * i] = (uint32_t)t & 0xffff; cl[2 * i + 1] = ((uint32_t)t >> 16) & 0xffff;
            inc = (uint32_t)(t >> 32);
        }
    }
    if (KIND < 2) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t w = bump ? lo19[i] : T[i];
            res[2 * i] = w & 0xffff; res[2 * i + 1] = w >> 16;
            V[2 * i] -= w & 0xffff; V[2 * i + 1] -= w >> 16;
        }
    }
    EDT_PRE(EDT_IS_A);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        EDT_ST(col + (size_t)k * n_rows, (uint64_t)res[k]);
        EDT_ST(col + (size_t)(16 + k) * n_rows, (uint64_t)cl[k]);
    }
    EDT_POST(EDT_IS_A, 0, col - threadIdx.x, gcol, grows);
    // V -= carry(x) p(x): p = [0xffed, 0xffff x 14, 0x7fff] = 0xffff (1 + x + .. + x^15) - 0x12 - 0x8000 x^15
    {
        uint32_t S = 0;
#pragma unroll
        for (int k = 0; k < 31; k++) {
            if (k < 16) S += cl[k];
            if (k >= 16) S -= cl[k - 16];
            int64_t cp = (int64_t)((uint64_t)S * 0xffffu);
            if (k < 16) cp -= (int64_t)(cl[k] * 0x12u);
            if (k >= 15) cp -= (int64_t)((uint64_t)cl[k - 15] << 15);
            V[k] -= cp;
        }
    }
    EDT_PRE(EDT_IS_B);
    int64_t prev = 0;
#pragma unroll
    for (int k = 0; k < 30; k++) {
        prev = (prev - V[k]) >> 16;                       // exact
        const uint32_t sh = (uint32_t)(prev + EDT_OFFSET);
        EDT_ST(col + (size_t)(32 + k) * n_rows, (uint64_t)(sh & 0xffff));
        EDT_ST(col + (size_t)(62 + k) * n_rows, (uint64_t)(sh >> 16));
    }
    EDT_POST(EDT_IS_B, 32, col - threadIdx.x, gcol, grows);
}

// Where a row's values go.  dst / stride: this thread's slot of column 0 of the current group of columns and the distance
// between columns there (the table itself, or the CTA's staging tile); gcol: the table at the CTA's first row, same column.
struct EdtSink {
    uint64_t *dst;
    size_t stride;
    uint64_t *gcol;
    size_t n_rows;
};
// the group of ncols columns is complete: move on to the next group (staged: the tile leaves through the bulk-copy engine;
// every thread of the CTA calls this)
BSX_HD void edt_flush(EdtSink &s, int ncols) {
#if defined(__CUDA_ARCH__) && EDT_STAGE
    s.gcol += (size_t)ncols * s.n_rows;
#else
    s.dst += (size_t)ncols * s.n_rows;
    s.gcol += (size_t)ncols * s.n_rows;
#endif
}

#if EDT_UNIFIED
template <int KIND>
BSX_HD void edt_op(const uint32_t *a1, const uint32_t *b1, const uint32_t *a2, const uint32_t *b2, uint32_t *res, uint64_t *col, size_t n_rows,
                   uint64_t *gcol, size_t grows) {
    edt_op_rt(KIND, a1, b1, a2, b2, res, col, n_rows, gcol, grows);
}
#endif

// the eight operations of (x1, y1) + (x2, y2) = (x3, y3), x3 / y3 known (the chain's affine values)
BSX_HD void edt_add_emit(const uint32_t *x1, const uint32_t *y1, const uint32_t *x2, const uint32_t *y2, uint32_t *x3, uint32_t *y3,
                         EdtSink &sink) {
    // d = -121665 / 121666 mod p
    const uint32_t D16[16] = {0x78a3, 0x1359, 0x4dca, 0x75eb, 0xd8ab, 0x4141, 0x0a4d, 0x0070,
                              0xe898, 0x7779, 0x4079, 0x8cc7, 0xfe73, 0x2b6f, 0x6cee, 0x5203};
    uint32_t xn[16], yn[16], m1[16], m2[16], f[16], df[16];
    edt_op<1>(x1, y2, x2, y1, xn, sink.dst, sink.stride, sink.gcol, sink.n_rows); edt_flush(sink, EDT_OP);
    edt_op<1>(y1, y2, x1, x2, yn, sink.dst, sink.stride, sink.gcol, sink.n_rows); edt_flush(sink, EDT_OP);
    edt_op<0>(x1, y1, nullptr, nullptr, m1, sink.dst, sink.stride, sink.gcol, sink.n_rows); edt_flush(sink, EDT_OP);
    edt_op<0>(x2, y2, nullptr, nullptr, m2, sink.dst, sink.stride, sink.gcol, sink.n_rows); edt_flush(sink, EDT_OP);
    edt_op<0>(m1, m2, nullptr, nullptr, f, sink.dst, sink.stride, sink.gcol, sink.n_rows); edt_flush(sink, EDT_OP);
    edt_op<0>(D16, f, nullptr, nullptr, df, sink.dst, sink.stride, sink.gcol, sink.n_rows); edt_flush(sink, EDT_OP);
    edt_op<2>(df, nullptr, xn, nullptr, x3, sink.dst, sink.stride, sink.gcol, sink.n_rows); edt_flush(sink, EDT_OP);
    edt_op<3>(df, nullptr, yn, nullptr, y3, sink.dst, sink.stride, sink.gcol, sink.n_rows); edt_flush(sink, EDT_OP);
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
// sink: where the row's 1540 values go (EdtSink); result (may be null): 64 bytes k * P, written by the last row.
BSX_HD void edt_row_core(bool real, uint32_t j, const uint8_t *scalar, const uint8_t *point, const uint32_t *af, EdtSink sink, uint8_t *result) {
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
    {
        uint64_t *col = sink.dst;
        const size_t st = sink.stride;
        EDT_ST(col, (uint64_t)bit);
        EDT_ST(col + st, (uint64_t)(real ? 1 : 0));
        EDT_ST(col + 2 * st, (uint64_t)(j == 0));
        EDT_ST(col + 3 * st, (uint64_t)(j == 255));
#pragma unroll
        for (int i = 0; i < 16; i++) {
            EDT_ST(col + (size_t)(4 + i) * st, (uint64_t)tx[i]);
            EDT_ST(col + (size_t)(20 + i) * st, (uint64_t)ty[i]);
            EDT_ST(col + (size_t)(36 + i) * st, (uint64_t)ax[i]);
            EDT_ST(col + (size_t)(52 + i) * st, (uint64_t)ay[i]);
        }
        // the 68 header columns span both halves of the tile: first use of each, nothing to wait for
        EDT_POST(EDT_IS_A, 0, col - threadIdx.x, sink.gcol, sink.n_rows);
#if defined(__CUDA_ARCH__) && EDT_STAGE
        if (threadIdx.x < 36) {
            if (EDT_STAGE != 2) EDT_BULK(sink.gcol + (size_t)(32 + threadIdx.x) * sink.n_rows, (col - threadIdx.x) + (size_t)(32 + threadIdx.x) * EDT_TILE);
            EDT_BULK_COMMIT();
        }
#endif
        edt_flush(sink, 68);
    }
    edt_add_emit(ax, ay, tx, ty, sx, sy, sink);
    edt_add_emit(tx, ty, tx, ty, dx, dy, sink);
#if defined(__CUDA_ARCH__) && EDT_STAGE
    if (EDT_IS_A || EDT_IS_B) EDT_BULK_WAIT();              // the tile must outlive its last bulk reads
#endif
    if (result && real && j == 255) {                        // k * P = the accumulator leaving the last row
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const uint32_t lx = bit ? sx[i] : ax[i], ly = bit ? sy[i] : ay[i];
            result[2 * i] = (uint8_t)lx; result[2 * i + 1] = (uint8_t)(lx >> 8);
            result[32 + 2 * i] = (uint8_t)ly; result[32 + 2 * i + 1] = (uint8_t)(ly >> 8);
        }
    }
}

}  // namespace edt
}  // namespace bsx

// Ed25519 scalar-multiplication execution trace on the device (the EdDSA half): the table a STARK over the
// Curta Ed25519 accelerator commits to, written column-major (one polynomial per column).
// Reference: `Ed25519Stark::prove` (PX/frontend/ecc/curve25519/curta/stark.rs:182-219) fills 256 rows per ScalarMul
// operation -- `write_trace_instructions` per row inside `chunks_par(256)` -- for the two scalar multiplications of every
// signature (s*G and h*A; collected at stark.rs:93-124).  The AIR is starkyx's `scalar_mul_batch` (UN-VENDORED,
// starkyx@ad8eb4ba): its column assignment is not on disk, so this LAYOUT IS OURS and PARITY IS UNPINNED.  It carries what
// that construction constrains: an affine double-and-add, one scalar bit per row, and for each of the 16 field operations
// of the row's two Edwards additions the starkyx-style witness -- result and quotient (`carry`) as 16 limbs of 16 bits and
// the polynomial w(x) = (lhs(x) - result(x) - carry(x) p(x)) / (x - 2^16), offset and split into 16-bit halves
// (include/bsx.h BSX_ED25519_TRACE_COLS has the column table; oracle/ed_trace.py is the CPU restatement).
// Sizing: one verify_skip circuit = 100 signatures = 200 multiplications = 51 200 rows -> 2^16 rows x 1540 columns x 8 B
// = 807 MB: an HBM write stream, like the SHA traces.
//
// Three kernels, because the work has three shapes:
//   ed_trace_chain_kernel  one thread per MULTIPLICATION: the 256 dependent steps in extended coordinates (sum = acc + temp
//                          and dbl = 2 temp at every step, whatever the bit -- the row witnesses both), X Y Z to scratch
//   ed_trace_affine_kernel one thread per 16 STEPS: their 32 Z's inverted with one field inversion (prefix products),
//                          affine sum / dbl of every row to scratch (128 B / row)
//   ed_trace_rows_kernel   one thread per ROW: picks temp / acc / sum / dbl of its row from the scratch (acc = the sum of
//                          the last set bit below j), redoes the 16 field operations on 16-bit limbs with exact integer
//                          quotients, and stores its 1540 values; a warp stores 32 consecutive rows of a column (256 B).
//                          (EDT_STAGE = 1: the same through a shared-memory tile and TMA bulk stores -- measured slower, ed_trace.cuh)
#include "common.cuh"
#include "ed_trace.cuh"

namespace bsx {

using namespace edt;

// resident CTAs of the row kernel per SM the register budget is set for (2: 255 registers, no spills; 2 / 3 / 4 / 5 measured: profiles/r04c_ed_trace_minb.txt)
#ifndef EDT_ROWS_MINB
#define EDT_ROWS_MINB 2
#endif

// chain[(m * 256 + j) * 64 ..]: sum and dbl of step j in extended coordinates; aff[(m * 256 + j) * 32 ..]: sum.x sum.y dbl.x dbl.y
// (8 words each, canonical little-endian)
__global__ void __launch_bounds__(64) ed_trace_chain_kernel(const uint8_t *__restrict__ scalars, const uint8_t *__restrict__ points, uint32_t n_muls,
                                                            int32_t *__restrict__ chain) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_muls) return;
    edt_forward_core<true>(scalars + (size_t)m * 32, points + (size_t)m * 64, chain + (size_t)m * 256 * EDT_CHAIN_WORDS);
}

// The same chain with EIGHT lanes per multiplication: lanes 0..3 of a group hold X, Y, Z, T of temp (the doubling chain),
// lanes 4..7 those of acc (the addition chain), and every lane does ONE field multiplication per time slot:
//   slot 1   D: X^2, Y^2, Z * 2Z, (X + Y)^2                A: (Y1 + X1)(Yt + Xt), (Y1 - X1)(Yt - Xt), Z1 Zt, T1 (2d Tt)
//   slot 2   completed -> extended, both groups: X' T', Z' Y', Z' T', X' Y'
//   slot 3   2d T of the new temp (one lane's worth of work, the other seven idle along)
// with the operands moved by warp shuffles and chosen by lane role, so that one instruction stream serves all roles.  A
// step's dependent path is 3 multiplications instead of 18: 24 multiplication slots of work instead of 18, but a small
// batch (one circuit = 200 multiplications) is bound by that path, not by issue slots.  Same formulas as ge_dbl /
// ge_add_cached / ge_p1p1_to_p3 (ed25519.cuh), same limb bounds (f <= 3 units, g <= 2 units).
__device__ __forceinline__ fe shfl_fe(const fe &a, int src) {
    fe r;
#pragma unroll
    for (int i = 0; i < 10; i++) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src);
    return r;
}

__global__ void __launch_bounds__(128) ed_trace_chain8_kernel(const uint8_t *__restrict__ scalars, const uint8_t *__restrict__ points, uint32_t n_muls,
                                                              int32_t *__restrict__ chain) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t m_raw = gid >> 3;
    const bool live = m_raw < n_muls;
    const uint32_t m = live ? m_raw : n_muls - 1;              // idle groups compute along (the shuffles need every lane)
    const int lane = threadIdx.x & 31, gb = lane & ~7, c = lane & 3;
    const bool isA = (lane & 4) != 0;
    const int gq = gb + (isA ? 4 : 0);
    const int32_t d2c[10] = BSX_FE_2D;
    const fe d2 = fe_const(d2c);
    uint32_t k[8];
#pragma unroll
    for (int i = 0; i < 8; i++) k[i] = ld_le32(scalars + (size_t)m * 32 + 4 * i);
    fe own;
    {
        const fe px = fe_frombytes(points + (size_t)m * 64), py = fe_frombytes(points + (size_t)m * 64 + 32);
        const fe pt = fe_mul(px, py);
        // D: X Y Z T of P;  A: the identity (0, 1, 1, 0)
        own = isA ? ((c == 1 || c == 2) ? fe_one() : fe_zero()) : (c == 0 ? px : c == 1 ? py : c == 2 ? fe_one() : pt);
    }
    fe t2d = fe_mul(own, d2);                                  // used from lane D3 only
    const int srcC = isA ? (c == 0 ? gb + 5 : c == 1 ? gb + 4 : c == 2 ? gb + 2 : gb + 3) : lane;
    int32_t *out = chain + (size_t)m * 256 * EDT_CHAIN_WORDS + (isA ? 0 : 32) + 10 * c;
    // Role masks: every choice below is a bitwise blend, so the warp runs ONE instruction stream (written with ?: the
    // compiler makes branches of them and the eight roles run one after the other: 1.30 ms instead of the time below).
    const int32_t mA = isA ? -1 : 0, m0 = c == 0 ? -1 : 0, m1 = c == 1 ? -1 : 0, m2 = c == 2 ? -1 : 0, m3 = c == 3 ? -1 : 0;
    const int32_t mD3 = ~mA & m3, mA0 = mA & m0, mA1 = mA & m1;
    const int32_t gS = mA0 | mD3, gV = mA & (m2 | m3), gOO = ~mA & m2, gO = ~mA & (m0 | m1);
    const int32_t mX = m0 | m3, mT = m0 | m2;
#pragma unroll 1
    for (int j = 0; j < 256; j++) {
        const fe Xt = shfl_fe(own, gb), Yt = shfl_fe(own, gb + 1);
        fe give;
#pragma unroll
        for (int i = 0; i < 10; i++) give.v[i] = (t2d.v[i] & mD3) | (own.v[i] & ~mD3);
        const fe vc = shfl_fe(give, srcC);                       // A0 <- Y1, A1 <- X1, A2 <- Zt, A3 <- 2d Tt
        fe f, g;
#pragma unroll
        for (int i = 0; i < 10; i++) {
            const int32_t o = own.v[i], v = vc.v[i], s = Xt.v[i] + Yt.v[i], dm = Yt.v[i] - Xt.v[i];
            // f: A0 X1 + Y1, A1 Y1 - X1, A2 Z1, A3 T1;  D0 X, D1 Y, D2 Z, D3 X + Y
            f.v[i] = ((o + (v & mA0) - (v & mA1)) & ~mD3) | (s & mD3);
            // g: A0 Yt + Xt, A1 Yt - Xt, A2 Zt, A3 2d Tt;  D0 X, D1 Y, D2 2Z, D3 X + Y
            g.v[i] = (s & gS) | (dm & mA1) | (v & gV) | ((o + o) & gOO) | (o & gO);
        }
        const fe p = fe_mul(f, g);
        const fe p0 = shfl_fe(p, gq), p1 = shfl_fe(p, gq + 1), p2 = shfl_fe(p, gq + 2), p3 = shfl_fe(p, gq + 3);
        fe Tc;
#pragma unroll
        for (int i = 0; i < 10; i++) {
            // D: xx yy 2zz (x+y)^2 -> Y' = yy + xx, Z' = yy - xx, X' = a - Y', T' = 2zz - Z'
            // A: a b zz c          -> X' = a - b, Y' = a + b, Z' = 2zz + c, T' = 2zz - c
            const int32_t yy = p1.v[i] + p0.v[i], zD = p1.v[i] - p0.v[i], dA = p2.v[i] + p2.v[i];
            const int32_t Xc = ((p0.v[i] - p1.v[i]) & mA) | ((p3.v[i] - yy) & ~mA);
            const int32_t Zc = ((dA + p3.v[i]) & mA) | (zD & ~mA);
            Tc.v[i] = ((dA - p3.v[i]) & mA) | ((p2.v[i] - zD) & ~mA);
            f.v[i] = (Xc & mX) | (Zc & ~mX);                      // slot 2: X' T', Z' Y', Z' T', X' Y'
            g.v[i] = yy;                                         // Y' is yy + xx resp. a + b: the same sum
        }
        const fe tt = fe_tighten(Tc);
#pragma unroll
        for (int i = 0; i < 10; i++) g.v[i] = (tt.v[i] & mT) | (g.v[i] & ~mT);
        const fe q = fe_mul(f, g);                                // D: coordinate c of 2 temp;  A: coordinate c of acc + temp
        if (live && c < 3) {
            int2 *o2 = reinterpret_cast<int2 *>(out + (size_t)j * EDT_CHAIN_WORDS);
#pragma unroll
            for (int i = 0; i < 5; i++) o2[i] = make_int2(q.v[2 * i], q.v[2 * i + 1]);
        }
        const int32_t take = ~mA | -(int32_t)((k[j >> 5] >> (j & 31)) & 1);
#pragma unroll
        for (int i = 0; i < 10; i++) own.v[i] = (q.v[i] & take) | (own.v[i] & ~take);
        t2d = fe_mul(own, d2);
    }
}

__global__ void __launch_bounds__(128) ed_trace_affine_kernel(const int32_t *__restrict__ chain, uint32_t n_groups, uint32_t *__restrict__ aff) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    edt_affine_core(chain + (size_t)g * EDT_GROUP * EDT_CHAIN_WORDS, aff + (size_t)g * EDT_GROUP * EDT_AFF_WORDS);
}

__global__ void __launch_bounds__(EDT_TILE, EDT_ROWS_MINB) ed_trace_rows_kernel(const uint8_t *__restrict__ scalars, const uint8_t *__restrict__ points,
                                                                                uint32_t n_muls, const uint32_t *__restrict__ aff, size_t n_rows,
                                                                                uint64_t *__restrict__ trace, uint8_t *__restrict__ results) {
    // n_rows is a multiple of EDT_TILE (a power of two >= 256): every thread has a row, every CTA a whole tile
    const size_t row0 = (size_t)blockIdx.x * EDT_TILE, row = row0 + threadIdx.x;
    const uint32_t m = (uint32_t)(row >> 8), j = (uint32_t)row & 255;
    const bool real = m < n_muls;
#if EDT_STAGE
    extern __shared__ __align__(128) uint64_t edt_stage[];      // [EDT_OP columns][EDT_TILE rows]
    const EdtSink sink{edt_stage + threadIdx.x, EDT_TILE, trace + row0, n_rows};
#else
    const EdtSink sink{trace + row, n_rows, trace + row0, n_rows};
#endif
    edt_row_core(real, j, real ? scalars + (size_t)m * 32 : nullptr, real ? points + (size_t)m * 64 : nullptr,
                 real ? aff + (size_t)m * 256 * EDT_AFF_WORDS : nullptr, sink, real && results ? results + (size_t)m * 64 : nullptr);
}

// ScalarMul operands of a batch of signatures, on the device: (s, G) and (h, A) per signature from the signature bytes and
// the 576-byte witness records of bsx_ed25519_batch (h at 64, A at 200) -- the requests Ed25519Stark::new collects
// (PX/frontend/ecc/curve25519/curta/stark.rs:93-124), in the order of the EdDSA schedule (eddsa.rs:161-203: s*G, then h*A).
// s of the DUMMY signature inactive lanes run on (eddsa.rs:28-30; the bytes 32..63 of DUMMY_SIG in k_ed25519.cu)
__constant__ uint8_t EDT_DUMMY_S[32] = {1, 41, 22, 121, 249, 46, 198, 145, 155, 102, 3, 210, 168, 135, 173, 55,
                                        252, 72, 45, 126, 169, 178, 191, 7, 153, 67, 112, 90, 150, 33, 140, 7};
__global__ void ed_trace_operands_kernel(uint32_t n_sigs, const uint8_t *__restrict__ sigs, uint32_t sig_stride, const uint8_t *__restrict__ active,
                                         uint32_t active_stride, const uint8_t *__restrict__ ed_out, uint8_t *__restrict__ scalars,
                                         uint8_t *__restrict__ points) {
    const uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), l = threadIdx.x & 31;
    if (i >= n_sigs) return;
    const int32_t gx[10] = BSX_FE_GX, gy[10] = BSX_FE_GY;
    uint8_t g[64];
    fe_tobytes(g, fe_const(gx)); fe_tobytes(g + 32, fe_const(gy));
    const uint8_t *rec = ed_out + (size_t)i * BSX_SIG_OUT_BYTES;
    const bool on = !active || active[(size_t)i * active_stride] != 0;
    scalars[(size_t)(2 * i) * 32 + l] = on ? sigs[(size_t)i * sig_stride + 32 + l] : EDT_DUMMY_S[l];
    scalars[(size_t)(2 * i + 1) * 32 + l] = rec[64 + l];
    points[(size_t)(2 * i) * 64 + l] = g[l];
    points[(size_t)(2 * i) * 64 + 32 + l] = g[32 + l];
    points[(size_t)(2 * i + 1) * 64 + l] = rec[200 + l];
    points[(size_t)(2 * i + 1) * 64 + 32 + l] = rec[232 + l];
}

}  // namespace bsx

using namespace bsx;

// bytes of device scratch bsx_ed25519_trace_dev needs for n_muls multiplications
extern "C" size_t bsx_ed25519_trace_scratch_bytes(uint32_t n_muls) { return (size_t)n_muls * 256 * (EDT_CHAIN_WORDS + EDT_AFF_WORDS) * 4; }

extern "C" int bsx_ed25519_trace_operands_dev(bsx_ctx *ctx, void *stream, uint32_t n_sigs, const uint8_t *sigs, uint32_t sig_stride,
                                              const uint8_t *active, uint32_t active_stride, const uint8_t *ed_out, uint8_t *scalars,
                                              uint8_t *points) {
    BSX_REQUIRE(ctx, ctx && (n_sigs == 0 || (sigs && ed_out && scalars && points && sig_stride >= 64)));
    if (n_sigs == 0) return BSX_OK;
    ed_trace_operands_kernel<<<(n_sigs + 3) / 4, 128, 0, (cudaStream_t)stream>>>(n_sigs, sigs, sig_stride, active, active_stride, ed_out, scalars,
                                                                                  points);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

// Execution trace of n_muls scalar multiplications k_m * P_m: trace = BSX_ED25519_TRACE_COLS columns of 2^log_rows rows
// (column-major u64 field elements), 256 rows per multiplication, the rest padded with rows of 0 * (0, 1).
// scalars: n_muls x 32 bytes little-endian (any 256-bit value); points: n_muls x 64 bytes (x, y little-endian, canonical,
// ON THE CURVE -- the outputs of the decompressions); results (may be null): n_muls x 64 bytes, k_m * P_m affine.
// Two halves, callable apart so that a caller with several batches can run the latency-bound first half of one batch
// beside the bandwidth-bound second half of the previous one (two streams, two scratch buffers: bench.py --mode trace):
//   bsx_ed25519_trace_points_dev  chain + affine kernels: the affine sum / dbl of every step -> scratch
//   bsx_ed25519_trace_rows_dev    row kernel: scratch -> trace (+ results)
extern "C" int bsx_ed25519_trace_points_dev(bsx_ctx *ctx, void *stream, const uint8_t *scalars, const uint8_t *points, uint32_t n_muls, void *scratch) {
    BSX_REQUIRE(ctx, ctx && (n_muls == 0 || (scalars && points && scratch)) && n_muls <= (1u << 22));
    BSX_REQUIRE(ctx, ((uintptr_t)scratch & 15) == 0);               // 16-byte vector accesses
    if (n_muls == 0) return BSX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *chain = reinterpret_cast<int32_t *>(scratch);
    uint32_t *aff = reinterpret_cast<uint32_t *>(chain + (size_t)n_muls * 256 * EDT_CHAIN_WORDS);
    // one circuit's worth of multiplications is bound by the chain's dependent path: eight lanes each; large batches fill the
    // machine with one thread each (less work per multiplication)
    const int lanes = ctx->tun[BSX_TUN_ED_TRACE_LANES] ? ctx->tun[BSX_TUN_ED_TRACE_LANES] : (n_muls <= 16384 ? 8 : 1);
    if (lanes == 8) ed_trace_chain8_kernel<<<(n_muls + 15) / 16, 128, 0, st>>>(scalars, points, n_muls, chain);
    else ed_trace_chain_kernel<<<(n_muls + 63) / 64, 64, 0, st>>>(scalars, points, n_muls, chain);
    BSX_LAUNCHED(ctx);
    const uint32_t n_groups = n_muls * (256 / EDT_GROUP);
    ed_trace_affine_kernel<<<(n_groups + 127) / 128, 128, 0, st>>>(chain, n_groups, aff);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

extern "C" int bsx_ed25519_trace_rows_dev(bsx_ctx *ctx, void *stream, const uint8_t *scalars, const uint8_t *points, uint32_t n_muls,
                                          uint32_t log_rows, const void *scratch, uint8_t *results, uint64_t *trace) {
    static_assert(BSX_ED25519_TRACE_COLS == 68 + 16 * EDT_OP, "column table");
    BSX_REQUIRE(ctx, ctx && trace && log_rows >= 8 && log_rows <= 30);
    const size_t n_rows = (size_t)1 << log_rows;
    BSX_REQUIRE(ctx, (size_t)n_muls * 256 <= n_rows);
    BSX_REQUIRE(ctx, n_muls == 0 || (scalars && points && scratch));
    BSX_REQUIRE(ctx, ((uintptr_t)scratch & 15) == 0 && ((uintptr_t)trace & 15) == 0);
    const uint32_t *aff = reinterpret_cast<const uint32_t *>(reinterpret_cast<const int32_t *>(scratch) + (size_t)n_muls * 256 * EDT_CHAIN_WORDS);
    const size_t smem = EDT_STAGE ? sizeof(uint64_t) * EDT_OP * EDT_TILE : 0;
    if (smem) BSX_CUDA(ctx, cudaFuncSetAttribute(ed_trace_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ed_trace_rows_kernel<<<(unsigned)(n_rows / EDT_TILE), EDT_TILE, smem, (cudaStream_t)stream>>>(scalars, points, n_muls, aff, n_rows, trace, results);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

extern "C" int bsx_ed25519_trace_dev(bsx_ctx *ctx, void *stream, const uint8_t *scalars, const uint8_t *points, uint32_t n_muls,
                                     uint32_t log_rows, void *scratch, uint8_t *results, uint64_t *trace) {
    BSX_REQUIRE(ctx, ctx && trace && log_rows >= 8 && log_rows <= 30 && ((size_t)n_muls * 256 <= ((size_t)1 << log_rows)));
    const int rc = bsx_ed25519_trace_points_dev(ctx, stream, scalars, points, n_muls, scratch);
    if (rc != BSX_OK) return rc;
    return bsx_ed25519_trace_rows_dev(ctx, stream, scalars, points, n_muls, log_rows, scratch, results, trace);
}

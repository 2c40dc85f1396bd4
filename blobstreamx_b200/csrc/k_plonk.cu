// Prover inner loops over the Goldilocks field: number-theoretic transforms, coset low-degree extension,
// Poseidon Merkle caps over the extension, FRI coefficient folding.  Together with gl_gate_quotient_kernel
// (k_goldilocks.cu) they keep a witness trace on the device from wire values to committed quotient:
//     trace values (n rows x W wires) --iNTT--> coefficients --coset LDE (rate 2^r)--> 2^r n points, bit-reversed
//         --Poseidon leaves + two-to-one layers--> Merkle cap        --gate constraints, alpha powers, / Z_H--> quotient
// Reference: plonky2 0.2.1 `prove_with_partition_witness` (UN-VENDORED; call site PX/backend/circuit/build.rs:69-75,
// config PX/frontend/builder/mod.rs:69): PolynomialBatch::from_values / lde_values, MerkleTree::new, compute_quotient_polys,
// fri_committed_trees.  Restated from the published algorithm; PARITY UNPINNED against plonky2's byte layout (root of unity,
// coset shift, digest order are parameters or documented here) -- pinned by algebraic identities in tests/.
//   * transforms are decimation-in-frequency: natural-order input, BIT-REVERSED output -- the order plonky2 hashes its
//     leaves in (reverse_index_bits_in_place), so leaf i of the Merkle tree is position i of the buffer (coalesced);
//   * a rate-2^r extension of a degree-n polynomial is 2^r independent size-n coset transforms: block rev_r(q) of the
//     output holds the evaluations on (g w_N^q) H_n, so the r zero-padding stages and the zero traffic never happen;
//   * every pass stages a tile in shared memory (up to 11 butterfly stages per pass over HBM), strided passes move
//     128-byte row segments.  HBM-bound: 8 B in + 8 B out per element per pass.
#include "common.cuh"
#include "poseidon.cuh"

namespace bsx {

using glf::add;
using glf::mul;
using glf::sub;

__device__ __forceinline__ uint64_t gl_pow(uint64_t b, uint64_t e) {
    uint64_t r = 1;
    while (e) {
        if (e & 1) r = mul(r, b);
        b = mul(b, b);
        e >>= 1;
    }
    return r;
}

// tw[k] = w^k, k < count
__global__ void gl_powers_kernel(uint64_t w, uint64_t first, uint32_t count, uint64_t *__restrict__ tw) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) tw[k] = mul(first, gl_pow(w, k));
}

struct NttArgs {
    const uint64_t *src;
    uint64_t *dst;
    size_t src_poly_stride, dst_poly_stride;
    size_t src_block_stride, dst_block_stride;   // block = blockIdx.z (the coset index of an extension), placed at bit-reversed z
    const uint64_t *tw;          // w_n^k, k < n/2
    const uint64_t *scale_hi;    // optional per-block input scaling  x_i *= scale_hi[z][i >> 10] * scale_lo[z][i & 1023]
    const uint64_t *scale_lo;
    uint32_t scale_hi_stride;
    uint64_t final_scale;        // applied at the store of the last pass (1/n of an inverse transform); 0 = none
    uint32_t L, lo, hi;          // transform size 2^L; this pass runs the stages of bits [lo, hi)
    uint32_t block_perm_bits;    // destination block = bit-reversal of blockIdx.z over this many bits
};

__device__ __forceinline__ uint32_t brev(uint32_t x, uint32_t bits) { return bits ? __brev(x) >> (32 - bits) : 0; }

__device__ __forceinline__ uint64_t ntt_load(const NttArgs &a, const uint64_t *src, uint32_t i) {
    uint64_t v = glf::canon(src[i]);
    if (a.scale_hi) {
        const uint32_t z = blockIdx.z;
        v = mul(v, mul(a.scale_hi[(size_t)z * a.scale_hi_stride + (i >> 10)], a.scale_lo[(size_t)z * 1024 + (i & 1023)]));
    }
    return v;
}

// stages [lo, hi) with lo >= 4: a tile is 2^(hi-lo) rows x 16 columns (one 128-byte segment per row)
__global__ void __launch_bounds__(256) ntt_dif_strided_kernel(NttArgs a) {
    extern __shared__ uint64_t sm[];
    constexpr uint32_t C = 16;
    const uint32_t S = a.hi - a.lo, rows = 1u << S;
    const uint32_t groups_lo = (1u << a.lo) / C;
    const uint32_t u_lo = blockIdx.x % groups_lo, u_hi = blockIdx.x / groups_lo;
    const size_t zb = brev(blockIdx.z, a.block_perm_bits);
    const uint64_t *src = a.src + (size_t)blockIdx.y * a.src_poly_stride + zb * a.src_block_stride;
    uint64_t *dst = a.dst + (size_t)blockIdx.y * a.dst_poly_stride + zb * a.dst_block_stride;
    const uint32_t base = (u_hi << a.hi) | (u_lo * C);
    for (uint32_t idx = threadIdx.x; idx < rows * C; idx += blockDim.x) {
        const uint32_t t = idx / C, c = idx % C;
        sm[idx] = ntt_load(a, src, base | (t << a.lo) | c);
    }
    __syncthreads();
    for (int b = (int)a.hi - 1; b >= (int)a.lo; b--) {
        const uint32_t hb = b - a.lo, half = 1u << hb, sh = a.L - 1 - b;
        for (uint32_t k = threadIdx.x; k < (rows / 2) * C; k += blockDim.x) {
            const uint32_t c = k % C, pr = k / C;
            const uint32_t t0 = ((pr >> hb) << (hb + 1)) | (pr & (half - 1)), t1 = t0 + half;
            const uint64_t x = sm[t0 * C + c], y = sm[t1 * C + c];
            const uint32_t imod = ((t0 & (half - 1)) << a.lo) | (u_lo * C + c);      // i mod 2^b
            sm[t0 * C + c] = add(x, y);
            sm[t1 * C + c] = mul(sub(x, y), __ldg(a.tw + ((size_t)imod << sh)));
        }
        __syncthreads();
    }
    for (uint32_t idx = threadIdx.x; idx < rows * C; idx += blockDim.x) {
        const uint32_t t = idx / C, c = idx % C;
        uint64_t v = sm[idx];
        if (a.final_scale && a.lo == 0) v = mul(v, a.final_scale);
        dst[base | (t << a.lo) | c] = v;
    }
}

// stages [0, hi): a tile is 2^hi contiguous elements; the 2^(hi-1) twiddles of these stages sit in shared memory
__global__ void __launch_bounds__(256) ntt_dif_contig_kernel(NttArgs a) {
    extern __shared__ uint64_t sm[];
    const uint32_t T = 1u << a.hi;
    uint64_t *stw = sm + T;
    const size_t zb = brev(blockIdx.z, a.block_perm_bits);
    const uint64_t *src = a.src + (size_t)blockIdx.y * a.src_poly_stride + zb * a.src_block_stride;
    uint64_t *dst = a.dst + (size_t)blockIdx.y * a.dst_poly_stride + zb * a.dst_block_stride;
    const uint32_t base = blockIdx.x << a.hi;
    for (uint32_t j = threadIdx.x; j < T; j += blockDim.x) sm[j] = ntt_load(a, src, base + j);
    for (uint32_t j = threadIdx.x; j < T / 2; j += blockDim.x) stw[j] = a.tw[(size_t)j << (a.L - a.hi)];
    __syncthreads();
    for (int b = (int)a.hi - 1; b >= 0; b--) {
        const uint32_t half = 1u << b;
        for (uint32_t k = threadIdx.x; k < T / 2; k += blockDim.x) {
            const uint32_t j0 = ((k >> b) << (b + 1)) | (k & (half - 1)), j1 = j0 + half;
            const uint64_t x = sm[j0], y = sm[j1];
            sm[j0] = add(x, y);
            sm[j1] = mul(sub(x, y), stw[(k & (half - 1)) << (a.hi - 1 - b)]);
        }
        __syncthreads();
    }
    for (uint32_t j = threadIdx.x; j < T; j += blockDim.x) dst[base + j] = a.final_scale ? mul(sm[j], a.final_scale) : sm[j];
}

// dst[bitrev_L(i)] = src[i] per polynomial, through a 32 x 32 shared-memory tile so that both sides move 256-byte rows:
// i = (hi5 | mid | lo5) -> bitrev = (rev(lo5) | rev(mid) | rev(hi5))
__global__ void __launch_bounds__(256) gl_bitrev_kernel(const uint64_t *__restrict__ src, uint64_t *__restrict__ dst, uint32_t L,
                                                        size_t src_poly_stride, size_t dst_poly_stride) {
    __shared__ uint64_t tile[32][33];
    const uint64_t *s = src + (size_t)blockIdx.y * src_poly_stride;
    uint64_t *d = dst + (size_t)blockIdx.y * dst_poly_stride;
    if (L < 10) {
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (1u << L); i += gridDim.x * blockDim.x) d[brev(i, L)] = s[i];
        return;
    }
    const uint32_t mbits = L - 10, mid = blockIdx.x, rmid = brev(mid, mbits);
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (uint32_t hi5 = w; hi5 < 32; hi5 += 8) tile[hi5][lane] = s[(hi5 << (L - 5)) | (mid << 5) | lane];
    __syncthreads();
    // destination rows: fixed rev(lo5) in the top bits, the low 5 bits = rev(hi5) run over the lanes
    for (uint32_t lo5 = w; lo5 < 32; lo5 += 8) d[(brev(lo5, 5) << (L - 5)) | (rmid << 5) | lane] = tile[brev(lane, 5)][lo5];
}

// ---- Poseidon Merkle tree over a poly-major buffer ----
// leaf i = hash_or_noop(values of the `width` polynomials at position i) (plonky2 MerkleTree::new over the transposed,
// index-bit-reversed LDE: here the buffer already is in that order).  One leaf per thread; the loads of a warp are 32
// consecutive words of one polynomial.
__global__ void __launch_bounds__(128) gl_merkle_leaves_kernel(const uint64_t *__restrict__ data, size_t poly_stride, uint32_t width,
                                                               uint32_t n_leaves, uint64_t *__restrict__ digests) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_leaves) return;
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = 0;
    if (width <= 4) {   // hash_or_noop: short leaves are padded, not hashed
        for (uint32_t e = 0; e < width; e++) s[e] = glf::canon(__ldcs(data + (size_t)e * poly_stride + i));
    } else {
#pragma unroll 1
        for (uint32_t e0 = 0; e0 < width; e0 += 8) {
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (e0 + k < width) s[k] = __ldcs(data + (size_t)(e0 + k) * poly_stride + i);   // any representative
            poseidon_permute(s);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) digests[4 * (size_t)i + k] = glf::canon(s[k]);
}
// one layer: parent j = two_to_one(child 2j, child 2j+1) = first 4 words of permute(l ‖ r ‖ 0^4)
__global__ void __launch_bounds__(128) gl_merkle_layer_kernel(const uint64_t *__restrict__ children, uint32_t n_parents,
                                                              uint64_t *__restrict__ parents) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_parents) return;
    uint64_t s[12];
    const ulonglong2 *c = reinterpret_cast<const ulonglong2 *>(children + 8 * (size_t)j);
    const ulonglong2 v0 = c[0], v1 = c[1], v2 = c[2], v3 = c[3];
    s[0] = v0.x; s[1] = v0.y; s[2] = v1.x; s[3] = v1.y; s[4] = v2.x; s[5] = v2.y; s[6] = v3.x; s[7] = v3.y;
    s[8] = s[9] = s[10] = s[11] = 0;
    poseidon_permute(s);
#pragma unroll
    for (int k = 0; k < 4; k++) parents[4 * (size_t)j + k] = glf::canon(s[k]);
}

// ---- FRI coefficient folding over the quadratic extension F[X]/(X^2 - 7) ----
// out[i] = sum_{j < arity} in[i * arity + j] * beta^j   (plonky2 fri_committed_trees: `reduce_with_powers(chunk, beta)`)
// elements are (c0, c1) pairs, interleaved
__device__ __forceinline__ void ext_mul(uint64_t a0, uint64_t a1, uint64_t b0, uint64_t b1, uint64_t &r0, uint64_t &r1) {
    r0 = add(mul(a0, b0), mul(7, mul(a1, b1)));
    r1 = add(mul(a0, b1), mul(a1, b0));
}
__global__ void __launch_bounds__(256) gl_fri_fold_kernel(const uint64_t *__restrict__ in, uint32_t n_out, uint32_t arity,
                                                          uint64_t beta0, uint64_t beta1, uint64_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(in) + (size_t)i * arity;
    uint64_t acc0 = 0, acc1 = 0;
    for (int j = (int)arity - 1; j >= 0; j--) {       // Horner in beta
        uint64_t t0, t1;
        ext_mul(acc0, acc1, beta0, beta1, t0, t1);
        const ulonglong2 v = p[j];
        acc0 = add(t0, glf::canon(v.x));
        acc1 = add(t1, glf::canon(v.y));
    }
    reinterpret_cast<ulonglong2 *>(out)[i] = make_ulonglong2(acc0, acc1);
}

}  // namespace bsx

using namespace bsx;

// plonky2 GoldilocksField constants (un-vendored; verified self-consistent: w = g^((p-1)/2^32), w^(2^31) = -1)
#define GL_GENERATOR 14293326489335486720ULL          /* MULTIPLICATIVE_GROUP_GENERATOR = coset shift */
#define GL_TWO_ADIC_ROOT 7277203076849721926ULL       /* POWER_OF_TWO_GENERATOR, order 2^32 */
static const uint64_t GLP = 0xFFFFFFFF00000001ULL;

static uint64_t h_mul(uint64_t a, uint64_t b) { return (uint64_t)((unsigned __int128)a * b % GLP); }
static uint64_t h_pow(uint64_t b, uint64_t e) {
    uint64_t r = 1;
    for (; e; e >>= 1, b = h_mul(b, b))
        if (e & 1) r = h_mul(r, b);
    return r;
}
static uint64_t h_inv(uint64_t a) { return h_pow(a, GLP - 2); }
static uint64_t h_root(uint32_t log_n) { return h_pow(GL_TWO_ADIC_ROOT, 1ULL << (32 - log_n)); }

extern "C" uint64_t bsx_gl_root_of_unity(uint32_t log_n) { return log_n <= 32 ? h_root(log_n) : 0; }
extern "C" uint64_t bsx_gl_coset_shift(void) { return GL_GENERATOR; }

// twiddle tables live in the ctx, one per (log_n, direction), built on first use on the calling stream
struct TwEntry { uint32_t log_n; int inverse; uint64_t *tw; };
struct bsx_plonk_cache { TwEntry e[64]; int n; };

static int get_twiddles(bsx_ctx *ctx, cudaStream_t st, uint32_t log_n, int inverse, const uint64_t **out) {
    if (!ctx->plonk) {
        ctx->plonk = (bsx_plonk_cache *)calloc(1, sizeof(bsx_plonk_cache));
        if (!ctx->plonk) return bsx::fail(ctx, BSX_ERR_NOMEM, "out of host memory%s%s");
    }
    bsx_plonk_cache *c = ctx->plonk;
    for (int i = 0; i < c->n; i++)
        if (c->e[i].log_n == log_n && c->e[i].inverse == inverse) { *out = c->e[i].tw; return BSX_OK; }
    if (c->n == 64) return bsx::fail(ctx, BSX_ERR_INVALID, "twiddle cache full%s%s");
    const uint32_t half = log_n ? 1u << (log_n - 1) : 1;
    uint64_t *tw = nullptr;
    BSX_CUDA(ctx, cudaMalloc(&tw, sizeof(uint64_t) * half));
    uint64_t w = h_root(log_n);
    if (inverse) w = h_inv(w);
    gl_powers_kernel<<<(half + 255) / 256, 256, 0, st>>>(w, 1, half, tw);
    BSX_LAUNCHED(ctx);
    // a later call may come on another stream: the table must be complete before anyone else can see it
    BSX_CUDA(ctx, cudaStreamSynchronize(st));
    c->e[c->n++] = TwEntry{log_n, inverse, tw};
    *out = tw;
    return BSX_OK;
}

void bsx_plonk_cache_free(bsx_ctx *ctx) {
    if (!ctx->plonk) return;
    for (int i = 0; i < ctx->plonk->n; i++) cudaFree(ctx->plonk->e[i].tw);
    free(ctx->plonk);
    ctx->plonk = nullptr;
}

// all DIF passes of one size-2^L transform over n_polys x n_blocks (blocks = cosets of an extension).  The first pass
// reads a.src (with the optional per-block scaling) and writes a.dst; the following passes run in place on a.dst.
static int run_dif(bsx_ctx *ctx, cudaStream_t st, const NttArgs &a, uint32_t n_polys, uint32_t n_blocks) {
    const uint32_t L = a.L;
    // plan: the last pass takes min(L, 11) contiguous bits; the bits above go in strided passes of at most 9 bits
    uint32_t cuts[8], nc = 0;
    const uint32_t last = L < 11 ? L : 11;
    uint32_t top = L;
    while (top > last) {
        uint32_t take = top - last;
        if (take > 9) take = (take + 1) / 2 > 9 ? 9 : (take + 1) / 2;
        cuts[nc++] = top - take;
        top -= take;
    }
    auto in_place = [&](NttArgs &p) {
        p.src = a.dst; p.src_poly_stride = a.dst_poly_stride; p.src_block_stride = a.dst_block_stride;
        p.scale_hi = p.scale_lo = nullptr;
    };
    uint32_t hi = L;
    for (uint32_t k = 0; k < nc; k++) {
        NttArgs p = a;
        p.lo = cuts[k]; p.hi = hi;
        if (k) in_place(p);
        const uint32_t S = p.hi - p.lo;
        const size_t smem = sizeof(uint64_t) * ((size_t)16 << S);
        if (smem > 48 * 1024) BSX_CUDA(ctx, cudaFuncSetAttribute(ntt_dif_strided_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((1u << (L - S)) / 16, n_polys, n_blocks);
        ntt_dif_strided_kernel<<<grid, 256, smem, st>>>(p);
        BSX_LAUNCHED(ctx);
        hi = cuts[k];
    }
    NttArgs p = a;
    p.lo = 0; p.hi = hi;
    if (nc) in_place(p);
    const size_t smem = sizeof(uint64_t) * (((size_t)1 << hi) + ((size_t)1 << hi) / 2 + 1);
    dim3 grid(1u << (L - hi), n_polys, n_blocks);
    ntt_dif_contig_kernel<<<grid, 256, smem, st>>>(p);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

// ---- entry points ----
// Forward / inverse transform of n_polys polynomials of 2^log_n elements each (poly-major, `stride` words apart).
//   natural-order input -> BIT-REVERSED output (decimation in frequency); `natural_out` adds the index permutation
//   (through `scratch`, n_polys * 2^log_n words, required then).  inverse: w^-1 and the 1/n factor.
extern "C" int bsx_gl_ntt_dev(bsx_ctx *ctx, void *stream, const uint64_t *in, uint64_t *out, uint32_t log_n, uint32_t n_polys,
                              size_t in_stride, size_t out_stride, int inverse, int natural_out, uint64_t *scratch) {
    BSX_REQUIRE(ctx, ctx && in && out && log_n >= 1 && log_n <= 27 && n_polys >= 1 && n_polys <= 65535);
    BSX_REQUIRE(ctx, in_stride >= ((size_t)1 << log_n) && out_stride >= ((size_t)1 << log_n) && (!natural_out || scratch));
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t *tw = nullptr;
    int rc = get_twiddles(ctx, st, log_n, inverse, &tw);
    if (rc) return rc;
    // Polynomials are processed in groups whose working set (~48 MB) fits the 126 MB L2: the passes of one group run back
    // to back, so what a pass writes is still in L2 when the next pass (and the index permutation) reads it -- HBM then sees
    // each element about once in and once out instead of once per pass.
    const size_t n = (size_t)1 << log_n;
    uint32_t group = (uint32_t)(((size_t)48 << 20) / (n * sizeof(uint64_t)));
    if (group < 1) group = 1;
    for (uint32_t p0 = 0; p0 < n_polys; p0 += group) {
        const uint32_t g = n_polys - p0 < group ? n_polys - p0 : group;
        NttArgs a{};
        a.src = in + (size_t)p0 * in_stride;
        a.dst = natural_out ? scratch + (size_t)p0 * n : out + (size_t)p0 * out_stride;
        a.src_poly_stride = in_stride; a.dst_poly_stride = natural_out ? n : out_stride;
        a.tw = tw; a.L = log_n;
        a.final_scale = inverse ? h_inv((uint64_t)1 << log_n) : 0;
        rc = run_dif(ctx, st, a, g, 1);
        if (rc) return rc;
        if (natural_out) {
            const uint32_t gx = log_n >= 10 ? 1u << (log_n - 10) : 1;
            gl_bitrev_kernel<<<dim3(gx, g), 256, 0, st>>>(scratch + (size_t)p0 * n, out + (size_t)p0 * out_stride, log_n, n, out_stride);
            BSX_LAUNCHED(ctx);
        }
    }
    return BSX_OK;
}

// Coset low-degree extension: coefficients (natural order, 2^log_n per polynomial) -> evaluations on shift * H_N,
// N = 2^(log_n + rate_bits), in bit-reversed index order (position i holds the point shift * w_N^bitrev(i)); shift = 0
// selects plonky2's coset shift (the multiplicative generator).  out: n_polys x N words, `out_stride` apart.
extern "C" int bsx_gl_lde_dev(bsx_ctx *ctx, void *stream, const uint64_t *coeffs, uint64_t *out, uint32_t log_n, uint32_t rate_bits,
                              uint32_t n_polys, size_t in_stride, size_t out_stride, uint64_t shift) {
    BSX_REQUIRE(ctx, ctx && coeffs && out && log_n >= 1 && rate_bits <= 6 && log_n + rate_bits <= 30 && n_polys >= 1 && n_polys <= 65535);
    const size_t n = (size_t)1 << log_n, N = n << rate_bits;
    BSX_REQUIRE(ctx, in_stride >= n && out_stride >= N);
    cudaStream_t st = (cudaStream_t)stream;
    if (!shift) shift = GL_GENERATOR;
    const uint64_t *tw = nullptr;
    int rc = get_twiddles(ctx, st, log_n, 0, &tw);
    if (rc) return rc;
    // per-coset input scaling s_q^i, s_q = shift * w_N^q, as two tables of powers: s_q^(1024 h) and s_q^l
    const uint32_t Q = 1u << rate_bits, n_hi = (uint32_t)((n + 1023) >> 10);
    uint64_t *tabs = nullptr;
    BSX_CUDA(ctx, cudaMallocAsync(&tabs, sizeof(uint64_t) * Q * ((size_t)n_hi + 1024), st));
    uint64_t *t_hi = tabs, *t_lo = tabs + (size_t)Q * n_hi;
    const uint64_t wN = h_root(log_n + rate_bits);
    for (uint32_t q = 0; q < Q; q++) {
        const uint64_t sq = h_mul(shift, h_pow(wN, q));
        gl_powers_kernel<<<(n_hi + 255) / 256, 256, 0, st>>>(h_pow(sq, 1024), 1, n_hi, t_hi + (size_t)q * n_hi);
        gl_powers_kernel<<<4, 256, 0, st>>>(sq, 1, 1024, t_lo + (size_t)q * 1024);
        ctx->launches += 2;
    }
    // groups of polynomials whose extension (~48 MB) fits L2, both passes of a group back to back (see bsx_gl_ntt_dev)
    uint32_t group = (uint32_t)(((size_t)48 << 20) / (N * sizeof(uint64_t)));
    if (group < 1) group = 1;
    for (uint32_t p0 = 0; p0 < n_polys && rc == BSX_OK; p0 += group) {
        NttArgs a{};
        a.src = coeffs + (size_t)p0 * in_stride; a.dst = out + (size_t)p0 * out_stride;
        a.src_poly_stride = in_stride; a.dst_poly_stride = out_stride;
        a.src_block_stride = 0; a.dst_block_stride = n; a.block_perm_bits = rate_bits;
        a.tw = tw; a.L = log_n; a.scale_hi = t_hi; a.scale_lo = t_lo; a.scale_hi_stride = n_hi;
        rc = run_dif(ctx, st, a, n_polys - p0 < group ? n_polys - p0 : group, Q);
    }
    const cudaError_t ef = cudaFreeAsync(tabs, st);
    if (rc) return rc;
    if (ef != cudaSuccess) return bsx::fail(ctx, BSX_ERR_CUDA, "cudaFreeAsync: %s%s", cudaGetErrorString(ef));
    return BSX_OK;
}

// Poseidon Merkle tree over a poly-major buffer of n_leaves positions x `width` polynomials (leaf i = the values at
// position i).  digests: layer 0 = n_leaves leaf digests, then every layer of two_to_one parents down to 2^cap_height
// nodes, concatenated (4 words per digest); the last 2^cap_height digests are the cap.  n_leaves a power of two.
extern "C" size_t bsx_gl_merkle_digest_words(uint32_t n_leaves, uint32_t cap_height) {
    size_t w = 0;
    for (size_t l = n_leaves; l >= ((size_t)1 << cap_height) && l >= 1; l >>= 1) {
        w += 4 * l;
        if (l == 1) break;
    }
    return w;
}
extern "C" int bsx_gl_merkle_caps_dev(bsx_ctx *ctx, void *stream, const uint64_t *data, size_t poly_stride, uint32_t width,
                                      uint32_t n_leaves, uint32_t cap_height, uint64_t *digests) {
    BSX_REQUIRE(ctx, ctx && data && digests && width >= 1 && n_leaves >= 1 && (n_leaves & (n_leaves - 1)) == 0);
    BSX_REQUIRE(ctx, cap_height <= 30 && ((size_t)1 << cap_height) <= n_leaves && (reinterpret_cast<uintptr_t>(digests) & 15) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    BSX_PIN_CARVEOUT(gl_merkle_leaves_kernel); BSX_PIN_CARVEOUT(gl_merkle_layer_kernel);
    gl_merkle_leaves_kernel<<<(n_leaves + 127) / 128, 128, 0, st>>>(data, poly_stride, width, n_leaves, digests);
    BSX_LAUNCHED(ctx);
    uint64_t *cur = digests;
    for (uint32_t l = n_leaves; l > (1u << cap_height); l >>= 1) {
        uint64_t *nxt = cur + 4 * (size_t)l;
        gl_merkle_layer_kernel<<<(l / 2 + 127) / 128, 128, 0, st>>>(cur, l / 2, nxt);
        BSX_LAUNCHED(ctx);
        cur = nxt;
    }
    return BSX_OK;
}

// Tables of the quotient evaluation (bsx_gl_gate_quotient_dev) for an extension laid out by bsx_gl_lde_dev:
//   alpha_pows[a * n_constraints + c] = alphas[a]^c;   zh_inv[rev_r(q)] = 1 / ((shift w_N^q)^n - 1), q < 2^rate_bits
struct ZhArgs { uint64_t v[64]; };
__global__ void gl_store_small_kernel(ZhArgs a, uint32_t n, uint64_t *__restrict__ out) {
    if (threadIdx.x < n) out[threadIdx.x] = a.v[threadIdx.x];
}
extern "C" int bsx_gl_quotient_tables_dev(bsx_ctx *ctx, void *stream, const uint64_t *alphas, uint32_t n_alphas, uint32_t n_constraints,
                                          uint32_t log_n, uint32_t rate_bits, uint64_t shift, uint64_t *alpha_pows, uint64_t *zh_inv) {
    BSX_REQUIRE(ctx, ctx && alphas && alpha_pows && zh_inv && n_alphas >= 1 && n_alphas <= 2 && n_constraints >= 1 && rate_bits <= 6 && log_n + rate_bits <= 32);
    cudaStream_t st = (cudaStream_t)stream;
    if (!shift) shift = GL_GENERATOR;
    for (uint32_t a = 0; a < n_alphas; a++) {
        BSX_REQUIRE(ctx, alphas[a] < GLP);
        gl_powers_kernel<<<(n_constraints + 255) / 256, 256, 0, st>>>(alphas[a], 1, n_constraints, alpha_pows + (size_t)a * n_constraints);
        BSX_LAUNCHED(ctx);
    }
    ZhArgs z{};
    const uint32_t Q = 1u << rate_bits;
    const uint64_t wN = h_root(log_n + rate_bits);
    for (uint32_t q = 0; q < Q; q++) {
        const uint64_t x_n = h_pow(h_mul(shift, h_pow(wN, q)), (uint64_t)1 << log_n);
        BSX_REQUIRE(ctx, x_n > 1);      // the coset must not meet the subgroup
        uint32_t r = 0;
        for (uint32_t b = 0; b < rate_bits; b++) r |= ((q >> b) & 1) << (rate_bits - 1 - b);
        z.v[r] = h_inv(x_n - 1);     // x_n is canonical and != 0, 1
    }
    gl_store_small_kernel<<<1, 64, 0, st>>>(z, Q, zh_inv);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

// FRI fold: n_out = n_in / arity extension elements (interleaved c0, c1 pairs)
extern "C" int bsx_gl_fri_fold_dev(bsx_ctx *ctx, void *stream, const uint64_t *in, uint32_t n_in, uint32_t arity_bits,
                                   uint64_t beta0, uint64_t beta1, uint64_t *out) {
    BSX_REQUIRE(ctx, ctx && in && out && arity_bits >= 1 && arity_bits <= 6 && n_in >= (1u << arity_bits) && (n_in & ((1u << arity_bits) - 1)) == 0);
    BSX_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 && beta0 < GLP && beta1 < GLP);
    const uint32_t n_out = n_in >> arity_bits;
    gl_fri_fold_kernel<<<(n_out + 255) / 256, 256, 0, (cudaStream_t)stream>>>(in, n_out, 1u << arity_bits, beta0, beta1, out);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

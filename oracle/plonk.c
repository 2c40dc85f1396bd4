/* TEST INFRASTRUCTURE -- CPU restatement of the prover inner loops over Goldilocks: transforms, coset
 * low-degree extension, Poseidon Merkle caps, quotient combination, FRI fold.
 * The reference's implementation is plonky2 0.2.1 (git dependency plonky2@53c5bc3e, UN-VENDORED; call sites
 * PX/backend/circuit/build.rs:69-75, PX/frontend/builder/mod.rs:69).  This file restates the PUBLISHED algorithms
 * (radix-2 NTT; PolynomialCoeffs::lde + coset_fft with the multiplicative generator as shift; MerkleTree::new with
 * hash_or_noop leaves and two_to_one compression; reduce_with_powers) in their textbook O(n log n) / O(n) forms, in
 * NATURAL index order.  PARITY UNPINNED against plonky2's own vectors (none are in the tree); pinned by algebraic
 * identities in tests/test_oracle_plonk.py (naive DFT, inverse round trip, direct polynomial evaluation, the Poseidon KAT). */
#include "bsx_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
#define GL_P 0xFFFFFFFF00000001ULL
#define GL_GENERATOR 14293326489335486720ULL    /* GoldilocksField::MULTIPLICATIVE_GROUP_GENERATOR (coset shift) */
#define GL_TWO_ADIC_ROOT 7277203076849721926ULL /* GoldilocksField::POWER_OF_TWO_GENERATOR, order 2^32 */

static uint64_t fadd(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a + b) % GL_P); }
static uint64_t fsub(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a + GL_P - (b % GL_P)) % GL_P); }
static uint64_t fmul(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a * b) % GL_P); }
static uint64_t fpow(uint64_t a, uint64_t e) {
    uint64_t r = 1;
    for (; e; e >>= 1, a = fmul(a, a))
        if (e & 1) r = fmul(r, a);
    return r;
}
static uint64_t finv(uint64_t a) { return fpow(a, GL_P - 2); }

uint64_t orc_gl_root_of_unity(uint32_t log_n) { return fpow(GL_TWO_ADIC_ROOT, 1ULL << (32 - log_n)); }
uint64_t orc_gl_coset_shift(void) { return GL_GENERATOR; }

static uint32_t bitrev(uint32_t x, uint32_t bits) {
    uint32_t r = 0;
    for (uint32_t b = 0; b < bits; b++) r |= ((x >> b) & 1u) << (bits - 1 - b);
    return r;
}

/* in-place transform, natural order in and out: X[k] = sum_j x[j] w^(jk); inverse: w^-1 and 1/n */
void orc_gl_ntt(uint64_t *x, uint32_t log_n, int inverse) {
    const uint32_t n = 1u << log_n;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t j = bitrev(i, log_n);
        if (i < j) { uint64_t t = x[i]; x[i] = x[j]; x[j] = t; }
    }
    uint64_t w = orc_gl_root_of_unity(log_n);
    if (inverse) w = finv(w);
    for (uint32_t s = 1; s <= log_n; s++) {       /* Cooley-Tukey, decimation in time */
        const uint32_t m = 1u << s, h = m >> 1;
        const uint64_t wm = fpow(w, n >> s);
        for (uint32_t k = 0; k < n; k += m) {
            uint64_t t = 1;
            for (uint32_t j = 0; j < h; j++) {
                const uint64_t u = x[k + j] % GL_P, v = fmul(x[k + j + h], t);
                x[k + j] = fadd(u, v);
                x[k + j + h] = fsub(u, v);
                t = fmul(t, wm);
            }
        }
    }
    if (inverse) {
        const uint64_t ni = finv(n);
        for (uint32_t i = 0; i < n; i++) x[i] = fmul(x[i], ni);
    }
}

/* PolynomialCoeffs::lde(rate_bits) then coset_fft(shift): zero-pad to N = n 2^rate_bits, scale coefficient j by shift^j,
 * transform.  out[k] = P(shift * w_N^k), natural order.  shift = 0: the multiplicative generator. */
void orc_gl_lde(const uint64_t *coeffs, uint32_t log_n, uint32_t rate_bits, uint64_t shift, uint64_t *out) {
    const uint32_t n = 1u << log_n, N = n << rate_bits;
    if (!shift) shift = GL_GENERATOR;
    uint64_t s = 1;
    for (uint32_t j = 0; j < N; j++) {
        out[j] = j < n ? fmul(coeffs[j] % GL_P, s) : 0;
        s = fmul(s, shift);
    }
    orc_gl_ntt(out, log_n + rate_bits, 0);
}

/* MerkleTree::new(leaves, cap_height): leaf digest = hash_or_noop (<= 4 elements: padded, else hash_n_to_hash_no_pad),
 * parents = two_to_one = first 4 words of permute(left ‖ right ‖ 0^4).  leaves: n_leaves x width, leaf-major.
 * digests: layer 0 (n_leaves), then every parent layer down to 2^cap_height nodes, concatenated; the tail is the cap. */
void orc_gl_merkle(const uint64_t *leaves, uint32_t width, uint32_t n_leaves, uint32_t cap_height, uint64_t *digests) {
    for (uint32_t i = 0; i < n_leaves; i++) {
        uint64_t *d = digests + 4 * (size_t)i;
        if (width <= 4) {
            for (uint32_t k = 0; k < 4; k++) d[k] = k < width ? leaves[(size_t)i * width + k] % GL_P : 0;
        } else {
            orc_poseidon_hash_no_pad(leaves + (size_t)i * width, width, d);
        }
    }
    uint64_t *cur = digests;
    for (uint32_t l = n_leaves; l > (1u << cap_height); l >>= 1) {
        uint64_t *nxt = cur + 4 * (size_t)l;
        for (uint32_t j = 0; j < l / 2; j++) {
            uint64_t s[12];
            memcpy(s, cur + 8 * (size_t)j, 64);
            s[8] = s[9] = s[10] = s[11] = 0;
            orc_poseidon_permute(s);
            memcpy(nxt + 4 * (size_t)j, s, 32);
        }
        cur = nxt;
    }
}

/* compute_quotient_polys, restricted to one gate type on every row: out[a][x] = zh_inv[x] * sum_c alphas[a]^c C_c(x)
 * (reduce_with_powers of the constraint terms, then the division by Z_H).  constraints: n_constraints x rows. */
void orc_gl_quotient_combine(const uint64_t *constraints, uint32_t n_constraints, uint32_t rows, const uint64_t *alphas,
                             uint32_t n_alphas, const uint64_t *zh_inv, uint64_t *out) {
    for (uint32_t a = 0; a < n_alphas; a++)
        for (uint32_t r = 0; r < rows; r++) {
            uint64_t acc = 0, ap = 1;
            for (uint32_t c = 0; c < n_constraints; c++) {
                acc = fadd(acc, fmul(constraints[(size_t)c * rows + r], ap));
                ap = fmul(ap, alphas[a]);
            }
            out[(size_t)a * rows + r] = fmul(acc, zh_inv[r]);
        }
}

/* fri_committed_trees: coeffs.chunks_exact(arity).map(|chunk| reduce_with_powers(chunk, beta)) over the quadratic
 * extension F[X]/(X^2 - 7); elements are interleaved (c0, c1) pairs */
void orc_gl_fri_fold(const uint64_t *in, uint32_t n_in, uint32_t arity_bits, uint64_t beta0, uint64_t beta1, uint64_t *out) {
    const uint32_t arity = 1u << arity_bits;
    for (uint32_t i = 0; i < n_in / arity; i++) {
        uint64_t a0 = 0, a1 = 0, p0 = 1, p1 = 0;     /* acc, beta^j */
        for (uint32_t j = 0; j < arity; j++) {
            const uint64_t c0 = in[2 * ((size_t)i * arity + j)] % GL_P, c1 = in[2 * ((size_t)i * arity + j) + 1] % GL_P;
            a0 = fadd(a0, fadd(fmul(c0, p0), fmul(7, fmul(c1, p1))));
            a1 = fadd(a1, fadd(fmul(c0, p1), fmul(c1, p0)));
            const uint64_t n0 = fadd(fmul(p0, beta0), fmul(7, fmul(p1, beta1))), n1 = fadd(fmul(p0, beta1), fmul(p1, beta0));
            p0 = n0; p1 = n1;
        }
        out[2 * (size_t)i] = a0; out[2 * (size_t)i + 1] = a1;
    }
}

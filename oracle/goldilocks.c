/*
 * oracle/goldilocks.c -- Goldilocks field, the five vendored u32 gates (constraint evaluation +
 * witness generators) and the Poseidon sponge.  TEST INFRASTRUCTURE ONLY (see bsx_oracle.h).
 *
 * Gates restate `eval_unfiltered` / `eval_unfiltered_base_packed` and `SimpleGenerator::run_once` of
 *   U32ArithmeticGate   PX/frontend/uint/num/u32/gates/arithmetic_u32.rs:107-166, 280-349, 383-431
 *   U32AddManyGate      PX/frontend/uint/num/u32/gates/add_many_u32.rs:107-146, 340-391
 *   U32SubtractionGate  PX/frontend/uint/num/u32/gates/subtraction_u32.rs:101-135, 305-350
 *   ComparisonGate      PX/frontend/uint/num/u32/gates/comparison.rs:118-195, 441-540
 *   U32RangeCheckGate   PX/frontend/uint/num/u32/gates/range_check_u32.rs:69-91, 202-224
 * Layout: wires[w*rows + r], constraints[c*rows + r] (plonky2's EvaluationVarsBaseBatch convention).
 * Field arithmetic here is the plain `% p` on unsigned __int128 -- deliberately not the reduction the
 * CUDA kernels use.
 *
 * Poseidon: plonky2 0.2.1 (un-vendored) width-12 permutation, x^7, 4+22+4 rounds, hash_n_to_hash_no_pad in
 * overwrite mode (call sites PX/frontend/hash/poseidon/poseidon256.rs:68,107; PX/utils/poseidon/mod.rs:31-36).
 * Round constants: poseidon_constants.h (generated, see scripts/gen_poseidon_constants.py); parity pinned by
 * the single reference KAT PX/frontend/hash/poseidon/poseidon256.rs:172-178.
 */
#include "bsx_oracle.h"
#include "poseidon_constants.h"
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
#define GL_P 0xFFFFFFFF00000001ULL

static uint64_t gl(uint64_t x) { return x >= GL_P ? x - GL_P : x; }
static uint64_t gl_add(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a + b) % GL_P); }
static uint64_t gl_sub(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a + GL_P - (b % GL_P)) % GL_P); }
static uint64_t gl_mul(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a * b) % GL_P); }
static uint64_t gl_pow(uint64_t a, uint64_t e) {
    uint64_t r = 1;
    a %= GL_P;
    while (e) { if (e & 1) r = gl_mul(r, a); a = gl_mul(a, a); e >>= 1; }
    return r;
}
static uint64_t gl_inv(uint64_t a) { return gl_pow(a, GL_P - 2); }

/* product over x in 0..base of (limb - x): the range-check polynomial every gate uses */
static uint64_t limb_product(uint64_t limb, uint32_t base) {
    uint64_t p = 1;
    for (uint32_t x = 0; x < base; x++) p = gl_mul(p, gl_sub(limb, x));
    return p;
}

static uint32_t ceil_div(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

uint32_t orc_gate_num_wires(uint32_t gate, uint32_t p0, uint32_t p1) {
    switch (gate) {
        case ORC_GATE_U32_ARITHMETIC: return p0 * (6 + 32);
        case ORC_GATE_U32_ADD_MANY: return p1 * (p0 + 3 + 19);
        case ORC_GATE_U32_SUBTRACTION: return p0 * (5 + 16);
        case ORC_GATE_U32_COMPARISON: return 4 + 5 * p1 + ceil_div(p0, p1) + 1;
        case ORC_GATE_U32_RANGE_CHECK: return p0 * 17;
    }
    return 0;
}
uint32_t orc_gate_num_constraints(uint32_t gate, uint32_t p0, uint32_t p1) {
    switch (gate) {
        case ORC_GATE_U32_ARITHMETIC: return p0 * (4 + 32);
        case ORC_GATE_U32_ADD_MANY: return p1 * (3 + 19);
        case ORC_GATE_U32_SUBTRACTION: return p0 * (3 + 16);
        case ORC_GATE_U32_COMPARISON: return 6 + 5 * p1 + ceil_div(p0, p1);
        case ORC_GATE_U32_RANGE_CHECK: return p0 * 17;
    }
    return 0;
}

#define W(col) (wires[(size_t)(col) * rows + r] % GL_P)
#define C(v) constraints[(size_t)(c++) * rows + r] = (v)

static void eval_arithmetic(uint32_t num_ops, const uint64_t *wires, uint32_t rows, uint32_t r, uint64_t *constraints) {
    uint32_t c = 0;
    for (uint32_t i = 0; i < num_ops; i++) {
        uint64_t m0 = W(6 * i), m1 = W(6 * i + 1), addend = W(6 * i + 2);
        uint64_t out_lo = W(6 * i + 3), out_hi = W(6 * i + 4), inverse = W(6 * i + 5);
        uint64_t computed = gl_add(gl_mul(m0, m1), addend);
        uint64_t diff = gl_sub(0xFFFFFFFFULL, out_hi);
        uint64_t hi_not_max = gl_sub(gl_mul(inverse, diff), 1);
        C(gl_mul(hi_not_max, out_lo));
        uint64_t combined = gl_add(gl_mul(out_hi, 1ULL << 32), out_lo);
        C(gl_sub(combined, computed));
        uint64_t lo = 0, hi = 0;
        for (int j = 31; j >= 0; j--) {
            uint64_t limb = W(6 * num_ops + 32 * i + j);
            C(limb_product(limb, 4));
            if (j < 16) lo = gl_add(gl_mul(lo, 4), limb);
            else hi = gl_add(gl_mul(hi, 4), limb);
        }
        C(gl_sub(lo, out_lo));
        C(gl_sub(hi, out_hi));
    }
}

static void eval_add_many(uint32_t na, uint32_t num_ops, const uint64_t *wires, uint32_t rows, uint32_t r, uint64_t *constraints) {
    uint32_t c = 0;
    for (uint32_t i = 0; i < num_ops; i++) {
        uint64_t computed = 0;
        for (uint32_t j = 0; j < na; j++) computed = gl_add(computed, W((na + 3) * i + j));
        computed = gl_add(computed, W((na + 3) * i + na));
        uint64_t out_res = W((na + 3) * i + na + 1), out_carry = W((na + 3) * i + na + 2);
        C(gl_sub(gl_add(gl_mul(out_carry, 1ULL << 32), out_res), computed));
        uint64_t res = 0, carry = 0;
        for (int j = 18; j >= 0; j--) {
            uint64_t limb = W((na + 3) * num_ops + 19 * i + j);
            C(limb_product(limb, 4));
            if (j < 16) res = gl_add(gl_mul(res, 4), limb);
            else carry = gl_add(gl_mul(carry, 4), limb);
        }
        C(gl_sub(res, out_res));
        C(gl_sub(carry, out_carry));
    }
}

static void eval_subtraction(uint32_t num_ops, const uint64_t *wires, uint32_t rows, uint32_t r, uint64_t *constraints) {
    uint32_t c = 0;
    for (uint32_t i = 0; i < num_ops; i++) {
        uint64_t x = W(5 * i), y = W(5 * i + 1), bin = W(5 * i + 2), out_res = W(5 * i + 3), out_b = W(5 * i + 4);
        uint64_t initial = gl_sub(gl_sub(x, y), bin);
        C(gl_sub(out_res, gl_add(initial, gl_mul(1ULL << 32, out_b))));
        uint64_t comb = 0;
        for (int j = 15; j >= 0; j--) {
            uint64_t limb = W(5 * num_ops + 16 * i + j);
            C(limb_product(limb, 4));
            comb = gl_add(gl_mul(comb, 4), limb);
        }
        C(gl_sub(comb, out_res));
        C(gl_mul(out_b, gl_sub(1, out_b)));
    }
}

static void eval_comparison(uint32_t num_bits, uint32_t nc, const uint64_t *wires, uint32_t rows, uint32_t r, uint64_t *constraints) {
    uint32_t c = 0, cb = ceil_div(num_bits, nc), chunk_size = 1u << cb;
    uint64_t first = W(0), second = W(1);
    uint64_t fc = 0, sc = 0; /* reduce_with_powers: sum chunk_i * (2^cb)^i */
    for (int i = (int)nc - 1; i >= 0; i--) {
        fc = gl_add(gl_mul(fc, chunk_size), W(4 + i));
        sc = gl_add(gl_mul(sc, chunk_size), W(4 + nc + i));
    }
    C(gl_sub(fc, first));
    C(gl_sub(sc, second));
    uint64_t msd_so_far = 0;
    for (uint32_t i = 0; i < nc; i++) {
        uint64_t f = W(4 + i), s = W(4 + nc + i);
        C(limb_product(f, chunk_size));
        C(limb_product(s, chunk_size));
        uint64_t diff = gl_sub(s, f), dummy = W(4 + 2 * nc + i), eq = W(4 + 3 * nc + i);
        C(gl_sub(gl_mul(diff, dummy), gl_sub(1, eq)));
        C(gl_mul(eq, diff));
        uint64_t inter = W(4 + 4 * nc + i);
        C(gl_sub(inter, gl_mul(eq, msd_so_far)));
        msd_so_far = gl_add(inter, gl_mul(gl_sub(1, eq), diff));
    }
    uint64_t msd = W(3);
    C(gl_sub(msd, msd_so_far));
    uint64_t bits_comb = 0;
    for (uint32_t i = 0; i <= cb; i++) {
        uint64_t bit = W(4 + 5 * nc + i);
        C(gl_mul(bit, gl_sub(1, bit)));
    }
    for (int i = (int)cb; i >= 0; i--) bits_comb = gl_add(gl_mul(bits_comb, 2), W(4 + 5 * nc + i));
    C(gl_sub(gl_add(chunk_size, msd), bits_comb));
    C(gl_sub(W(2), W(4 + 5 * nc + cb)));
}

static void eval_range_check(uint32_t nl, const uint64_t *wires, uint32_t rows, uint32_t r, uint64_t *constraints) {
    uint32_t c = 0;
    for (uint32_t i = 0; i < nl; i++) {
        uint64_t sum = 0;
        for (int j = 15; j >= 0; j--) sum = gl_add(gl_mul(sum, 4), W(nl + 16 * i + j));
        C(gl_sub(sum, W(i)));
        for (uint32_t j = 0; j < 16; j++) C(limb_product(W(nl + 16 * i + j), 4));
    }
}

int orc_gate_eval(uint32_t gate, uint32_t p0, uint32_t p1, const uint64_t *wires, uint32_t rows,
                  uint64_t *constraints, int threads) {
    if (gate > ORC_GATE_U32_RANGE_CHECK) return -1;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (uint32_t r = 0; r < rows; r++) {
        switch (gate) {
            case ORC_GATE_U32_ARITHMETIC: eval_arithmetic(p0, wires, rows, r, constraints); break;
            case ORC_GATE_U32_ADD_MANY: eval_add_many(p0, p1, wires, rows, r, constraints); break;
            case ORC_GATE_U32_SUBTRACTION: eval_subtraction(p0, wires, rows, r, constraints); break;
            case ORC_GATE_U32_COMPARISON: eval_comparison(p0, p1, wires, rows, r, constraints); break;
            default: eval_range_check(p0, wires, rows, r, constraints); break;
        }
    }
    return 0;
}

/* ---- witness generators (SimpleGenerator::run_once): fill the dependent wires of each row ---- */
#undef W
#define W(col) wires[(size_t)(col) * rows + r]
static void split_limbs(uint64_t v, uint32_t n, uint32_t bits, uint64_t *wires, uint32_t rows, uint32_t r, uint32_t col0) {
    for (uint32_t j = 0; j < n; j++) { W(col0 + j) = v & ((1ULL << bits) - 1); v >>= bits; }
}

int orc_gate_witness(uint32_t gate, uint32_t p0, uint32_t p1, uint64_t *wires, uint32_t rows, int threads) {
    if (gate > ORC_GATE_U32_RANGE_CHECK) return -1;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (uint32_t r = 0; r < rows; r++) {
        if (gate == ORC_GATE_U32_ARITHMETIC) {
            for (uint32_t i = 0; i < p0; i++) {
                uint64_t out = gl_add(gl_mul(W(6 * i) % GL_P, W(6 * i + 1) % GL_P), W(6 * i + 2) % GL_P);
                uint64_t hi = out >> 32, lo = out & 0xFFFFFFFFULL;
                W(6 * i + 3) = lo;
                W(6 * i + 4) = hi;
                uint64_t diff = 0xFFFFFFFFULL - hi;
                W(6 * i + 5) = diff == 0 ? 0 : gl_inv(diff);
                split_limbs(out, 32, 2, wires, rows, r, 6 * p0 + 32 * i);
            }
        } else if (gate == ORC_GATE_U32_ADD_MANY) {
            uint32_t na = p0;
            for (uint32_t i = 0; i < p1; i++) {
                uint64_t out = 0;
                for (uint32_t j = 0; j <= na; j++) out = gl_add(out, W((na + 3) * i + j) % GL_P);
                uint64_t carry = out >> 32, res = out & 0xFFFFFFFFULL;
                W((na + 3) * i + na + 1) = res;
                W((na + 3) * i + na + 2) = carry;
                split_limbs(res, 16, 2, wires, rows, r, (na + 3) * p1 + 19 * i);
                split_limbs(carry, 3, 2, wires, rows, r, (na + 3) * p1 + 19 * i + 16);
            }
        } else if (gate == ORC_GATE_U32_SUBTRACTION) {
            for (uint32_t i = 0; i < p0; i++) {
                uint64_t initial = gl_sub(gl_sub(W(5 * i) % GL_P, W(5 * i + 1)), W(5 * i + 2));
                uint64_t borrow = initial > (1ULL << 32) ? 1 : 0; /* subtraction_u32.rs:319 (strict >) */
                uint64_t res = gl_add(initial, gl_mul(1ULL << 32, borrow));
                W(5 * i + 3) = res;
                W(5 * i + 4) = borrow;
                split_limbs(res, 16, 2, wires, rows, r, 5 * p0 + 16 * i);
            }
        } else if (gate == ORC_GATE_U32_COMPARISON) {
            uint32_t nc = p1, cb = ceil_div(p0, p1);
            uint64_t a = gl(W(0) % GL_P), b = gl(W(1) % GL_P), msd = 0;
            W(2) = a <= b;
            for (uint32_t i = 0; i < nc; i++) {
                uint32_t sh = cb * i;
                uint64_t f = sh < 64 ? (a >> sh) & ((1ULL << cb) - 1) : 0, s = sh < 64 ? (b >> sh) & ((1ULL << cb) - 1) : 0;
                W(4 + i) = f;
                W(4 + nc + i) = s;
                W(4 + 2 * nc + i) = f == s ? 1 : gl_inv(gl_sub(s, f));
                W(4 + 3 * nc + i) = f == s;
                if (f != s) { msd = gl_sub(s, f); W(4 + 4 * nc + i) = 0; }
                else W(4 + 4 * nc + i) = msd;
            }
            W(3) = msd;
            uint64_t t = gl_add(1ULL << cb, msd);
            for (uint32_t i = 0; i <= cb; i++) { W(4 + 5 * nc + i) = t & 1; t >>= 1; }
        } else {
            for (uint32_t i = 0; i < p0; i++) split_limbs((uint32_t)gl(W(i) % GL_P), 16, 2, wires, rows, r, p0 + 16 * i);
        }
    }
    return 0;
}

/* ---- Poseidon ---- */
static const uint64_t MDS_CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
static const uint64_t MDS_DIAG[12] = {8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

const uint64_t *orc_poseidon_round_constants(void) { return ORC_POSEIDON_RC; }

void orc_poseidon_permute(uint64_t s[12]) {
    for (int i = 0; i < 12; i++) s[i] %= GL_P;
    for (int r = 0; r < 30; r++) {
        for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], ORC_POSEIDON_RC[12 * r + i]);
        int full = r < 4 || r >= 26;
        for (int i = 0; i < (full ? 12 : 1); i++) {
            uint64_t x2 = gl_mul(s[i], s[i]), x4 = gl_mul(x2, x2);
            s[i] = gl_mul(gl_mul(x4, x2), s[i]);
        }
        uint64_t t[12];
        for (int k = 0; k < 12; k++) {
            u128 acc = 0;
            for (int i = 0; i < 12; i++) acc += (u128)s[(i + k) % 12] * MDS_CIRC[i];
            acc += (u128)s[k] * MDS_DIAG[k];
            t[k] = (uint64_t)(acc % GL_P);
        }
        memcpy(s, t, sizeof t);
    }
}

/* plonky2 hash_n_to_hash_no_pad: rate 8, overwrite mode, first 4 state elements out */
void orc_poseidon_hash_no_pad(const uint64_t *in, uint32_t n, uint64_t out[4]) {
    uint64_t st[12] = {0};
    for (uint32_t i = 0; i < n; i += 8) {
        uint32_t k = n - i < 8 ? n - i : 8;
        for (uint32_t j = 0; j < k; j++) st[j] = in[i + j] % GL_P;
        orc_poseidon_permute(st);
    }
    memcpy(out, st, 4 * sizeof(uint64_t));
}

void orc_poseidon_batch(const uint64_t *in, const uint32_t *offsets, uint32_t n, uint64_t *out, int threads) {
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (uint32_t i = 0; i < n; i++) orc_poseidon_hash_no_pad(in + offsets[i], offsets[i + 1] - offsets[i], out + 4 * (size_t)i);
}

// Poseidon permutation over Goldilocks (width 12, x^7, 4 + 22 + 4 rounds) for the hash kernels (k_goldilocks.cu: sponge
// batches, k_plonk.cu: Merkle leaves and two-to-one compression).  plonky2 0.2.1 `PoseidonHash` (un-vendored; call sites
// PX/frontend/hash/poseidon/poseidon256.rs:61-86, PX/utils/poseidon/mod.rs:9-66); constants regenerated and pinned by the
// reference's single KAT (scripts/gen_poseidon_constants.py).
#pragma once
#include "goldilocks.cuh"
#include "poseidon_constants.cuh"

namespace bsx {

// ---- Poseidon (width 12, x^7, 4 + 22 + 4 rounds) ----
// The state is kept as arbitrary 64-bit representatives (weak reductions, glf::): every step below is correct for
// any representative, and only the four squeezed words are made canonical at the end.
__device__ __forceinline__ uint64_t gl_mul_weak(uint64_t a, uint64_t b) {
    uint64_t hi, lo;
    glf::mul128(a, b, 0, hi, lo);
    return glf::reduce128_weak(hi, lo);
}
__device__ __forceinline__ uint64_t gl_pow7(uint64_t x) {
    const uint64_t x2 = gl_mul_weak(x, x), x4 = gl_mul_weak(x2, x2);
    return gl_mul_weak(gl_mul_weak(x4, x2), x);
}
// s + c for any 64-bit s and canonical c: wrapped + eps on carry (cannot carry twice because c < p)
__device__ __forceinline__ uint64_t gl_add_const_weak(uint64_t s, uint64_t c) {
    uint64_t r;
    asm("{\n\t"
        ".reg .u32 cy;\n\t"
        "add.cc.u64 %0, %1, %2;\n\t"
        "addc.u32 cy, 0, 0;\n\t"
        "mad.wide.u32 %0, cy, 0xFFFFFFFF, %0;\n\t"
        "}"
        : "=&l"(r)
        : "l"(s), "l"(c));
    return r;
}

// MDS = circulant(CIRC) + diag(8, 0, ...).  The constants are below 2^6 and sum to 284 < 2^8.2, so with the state cut
// into limbs of 22 / 21 / 21 bits every column sum stays below 2^31.2: plain 32-bit IMADs (2 issue cycles on the FMA
// pipe) instead of IMAD.WIDE (4 cycles), 3 x 144 of them per round instead of 2 x 144 wide ones; the three sums of an
// output are folded into one 128-bit value (a0 + a1 2^22 + a2 2^43) and reduced once, weakly.  Measured against the
// form with two IMAD.WIDE sums of 32-bit halves: 490 -> 623 M permutations/s (profiles/r01o_poseidon_*.json).
__device__ __forceinline__ void poseidon_mds(uint64_t s[12]) {
    constexpr uint32_t CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    uint32_t l0[12], l1[12], l2[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        l0[i] = (uint32_t)s[i] & 0x3fffffu;
        l1[i] = (uint32_t)(s[i] >> 22) & 0x1fffffu;
        l2[i] = (uint32_t)(s[i] >> 43);
    }
#pragma unroll
    for (int k = 0; k < 12; k++) {
        uint32_t a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            a0 += l0[(i + k) % 12] * CIRC[i];
            a1 += l1[(i + k) % 12] * CIRC[i];
            a2 += l2[(i + k) % 12] * CIRC[i];
        }
        if (k == 0) { a0 += l0[0] * 8; a1 += l1[0] * 8; a2 += l2[0] * 8; }  // MDS_MATRIX_DIAG = [8, 0, ...]
        // value = a0 + a1 2^22 + a2 2^43  (< 2^75)
        uint64_t lo = (uint64_t)a0 + ((uint64_t)a1 << 22), hi = (uint64_t)a2 >> 21;
        asm("add.cc.u64 %0, %0, %2;\n\t"
            "addc.u64 %1, %1, 0;"
            : "+l"(lo), "+l"(hi)
            : "l"((uint64_t)a2 << 43));
        s[k] = glf::reduce128_weak(hi, lo);
    }
}

__device__ __forceinline__ void poseidon_permute(uint64_t s[12]) {
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl_add_const_weak(s[i], BSX_POSEIDON_RC[12 * r + i]);
        if (r < 4 || r >= 26) {
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = gl_pow7(s[i]);
        } else {
            s[0] = gl_pow7(s[0]);
        }
        poseidon_mds(s);
    }
}

}  // namespace bsx

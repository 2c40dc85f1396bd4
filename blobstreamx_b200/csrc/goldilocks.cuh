// Goldilocks field (p = 2^64 - 2^32 + 1) primitives for the gate kernels, written against the two integer pipes
// of an SM sub-partition: 32x32->64 multiply-adds (IMAD.WIDE.U32, FMA pipe) build the 128-bit products, and every
// conditional +-epsilon of the reduction is a carry chain instead of a 64-bit compare + select.  PTX keeps the
// carry flag ARM-style (after sub.cc, CF = NOT borrow), so `subc m, 0, 0` after a subtraction is the borrow mask
// (0 or 0xFFFFFFFF = epsilon = 2^64 mod p), and after an addition `addc c, 0, 0` is the carry bit, applied as
// x + c*epsilon by one mad.wide.u32 on the FMA pipe.
//   reference arithmetic: plonky2 GoldilocksField (un-vendored; call sites PX/frontend/uint/num/u32/gates/*.rs)
#pragma once
#include <stdint.h>

namespace bsx {
namespace glf {

constexpr uint64_t P = 0xFFFFFFFF00000001ULL;
constexpr uint64_t EPS = 0xFFFFFFFFULL;

// x mod p for any 64-bit x:  x >= p  <=>  x + eps carries out of 64 bits, and then x - p = (x + eps) mod 2^64
__device__ __forceinline__ uint64_t canon(uint64_t x) {
    uint64_t r;
    asm("{\n\t"
        ".reg .u64 s;\n\t"
        ".reg .u32 c;\n\t"
        "add.cc.u64 s, %1, 0xFFFFFFFF;\n\t"
        "addc.u32 c, 0, 0;\n\t"                       // c = (x >= p)
        "mad.wide.u32 %0, c, 0xFFFFFFFF, %1;\n\t"     // x + c eps, on the FMA pipe
        "}"
        : "=l"(r)
        : "l"(x));
    return r;
}
// a - b mod p, canonical inputs -> canonical output
__device__ __forceinline__ uint64_t sub(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("{\n\t"
        ".reg .u64 mm;\n\t"
        ".reg .u32 m;\n\t"
        "sub.cc.u64 %0, %1, %2;\n\t"
        "subc.u32 m, 0, 0;\n\t"            // borrow ? 0xFFFFFFFF : 0
        "cvt.u64.u32 mm, m;\n\t"
        "sub.u64 %0, %0, mm;\n\t"          // wrapped - eps = a - b + p
        "}"
        : "=l"(r)
        : "l"(a), "l"(b));
    return r;
}
// a + b mod p, canonical inputs -> canonical output
__device__ __forceinline__ uint64_t add(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("{\n\t"
        ".reg .u32 c;\n\t"
        "add.cc.u64 %0, %1, %2;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "mad.wide.u32 %0, c, 0xFFFFFFFF, %0;\n\t"     // wrapped + eps on carry (cannot carry again: a + b < 2p)
        "}"
        : "=&l"(r)
        : "l"(a), "l"(b));
    return canon(r);
}

// (hi, lo) = a * b + c, c < 2^32: four independent IMAD.WIDE.U32, then the middle column is one 3-input add with
// two carry-outs (IADD3) -- no zero-extended register pairs to set up for the addends
__device__ __forceinline__ void mul128(uint64_t a, uint64_t b, uint32_t c, uint64_t &hi, uint64_t &lo) {
    const uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    const uint64_t p00 = (uint64_t)a0 * b0 + c, p01 = (uint64_t)a0 * b1, p10 = (uint64_t)a1 * b0, p11 = (uint64_t)a1 * b1;
    const uint64_t m = (p00 >> 32) + (uint32_t)p01 + (uint32_t)p10;          // < 3 * 2^32
    lo = (m << 32) | (uint32_t)p00;
    // hi = p11 + (p01 >> 32) + (p10 >> 32) + (m >> 32): each "64-bit + zero-extended 32-bit" is one mad.wide.u32 by 1
    // on the FMA pipe instead of an IADD3 / IADD3.X pair on the (busier) ALU pipe
    hi = p11;
    asm("mad.wide.u32 %0, %1, 1, %0;\n\t"
        "mad.wide.u32 %0, %2, 1, %0;\n\t"
        "mad.wide.u32 %0, %3, 1, %0;"
        : "+l"(hi)
        : "r"((uint32_t)(p01 >> 32)), "r"((uint32_t)(p10 >> 32)), "r"((uint32_t)(m >> 32)));
}

// any 64-bit representative of (hi 2^64 + lo) mod p:  2^64 = eps, 2^96 = -1 (mod p)
//   lo - hi_hi (+p on borrow)  +  hi_lo * eps (= (hi_lo << 32) - hi_lo)  (+eps on carry)
__device__ __forceinline__ uint64_t reduce128_weak(uint64_t hi, uint64_t lo) {
    uint64_t r;
    asm("{\n\t"
        ".reg .u32 h0, h1, m, z;\n\t"
        ".reg .u64 t, x, y;\n\t"
        "mov.b64 {h0, h1}, %1;\n\t"
        "cvt.u64.u32 x, h1;\n\t"
        "sub.cc.u64 t, %2, x;\n\t"         // lo - hi_hi
        "subc.u32 m, 0, 0;\n\t"
        "cvt.u64.u32 x, m;\n\t"
        "sub.u64 t, t, x;\n\t"             // - eps on borrow
        "mov.u32 z, 0;\n\t"
        "mov.b64 x, {z, h0};\n\t"          // hi_lo << 32
        "cvt.u64.u32 y, h0;\n\t"
        "sub.u64 x, x, y;\n\t"             // hi_lo * eps
        "add.cc.u64 t, t, x;\n\t"
        "addc.u32 m, 0, 0;\n\t"
        "mad.wide.u32 %0, m, 0xFFFFFFFF, t;\n\t"   // + eps on carry
        "}"
        : "=l"(r)
        : "l"(hi), "l"(lo));
    return r;
}
__device__ __forceinline__ uint64_t mul(uint64_t a, uint64_t b) {
    uint64_t hi, lo;
    mul128(a, b, 0, hi, lo);
    return canon(reduce128_weak(hi, lo));
}

// prod_{x < 4} (l - x) for ANY 64-bit representative l (sub() then returns some representative of l - 3, which is
// all the product needs):  with w = l (l - 3) + 1 the product is (w - 1)(w + 1) = w^2 - 1.
// Two 128-bit products, two weak reductions; the "+ 1" rides in the first product's addend and the "- 1" is
// "+ (p - 1)" on the second (w^2 + p - 1 < 2^128 for any 64-bit w).
__device__ __forceinline__ uint64_t limb_product4(uint64_t l) {
    uint64_t hi, lo;
    mul128(l, sub(l, 3), 1, hi, lo);
    const uint64_t w = reduce128_weak(hi, lo);
    mul128(w, w, 0, hi, lo);
    asm("add.cc.u64 %0, %0, 0xFFFFFFFF00000000;\n\t"
        "addc.u64 %1, %1, 0;"
        : "+l"(lo), "+l"(hi));
    return canon(reduce128_weak(hi, lo));
}

// sum_t limb_t 4^t over 64-bit limbs, kept as two 64-bit sums of 32-bit halves times 4^t (IMAD.WIDE.U32
// with an immediate multiplier): 16 limbs stay below 2^63.  value() folds them into one canonical element.
struct Radix4Sum {
    uint64_t lo = 0, hi = 0;
    // += limb * 4^t, t < 16 (any 64-bit limb: the sum is linear, representatives need not be canonical)
    __device__ __forceinline__ void add(uint64_t limb, int t) {
        const uint32_t l0 = (uint32_t)limb, l1 = (uint32_t)(limb >> 32), k = 1u << (2 * t);
        asm("mad.wide.u32 %0, %2, %4, %0;\n\t"
            "mad.wide.u32 %1, %3, %4, %1;"
            : "+l"(lo), "+l"(hi)
            : "r"(l0), "r"(l1), "r"(k));
    }
    // this = this + other * 4^8 (other holds limbs 8..15 accumulated with T = 0..7)
    __device__ __forceinline__ void append_high8(const Radix4Sum &o) { lo += o.lo << 16; hi += o.hi << 16; }
    __device__ __forceinline__ uint64_t value() const {
        uint64_t l = lo, h = hi >> 32;
        asm("add.cc.u64 %0, %0, %2;\n\t"
            "addc.u64 %1, %1, 0;"
            : "+l"(l), "+l"(h)
            : "l"(hi << 32));
        return canon(reduce128_weak(h, l));
    }
};

}  // namespace glf
}  // namespace bsx

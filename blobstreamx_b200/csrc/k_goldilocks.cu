// K6/K7/K8: Goldilocks u32-gate constraint evaluation, gate witness generators, Poseidon sponge.
//   U32ArithmeticGate   PX/frontend/uint/num/u32/gates/arithmetic_u32.rs:280-349 (eval), 383-431 (generator)
//   U32AddManyGate      PX/frontend/uint/num/u32/gates/add_many_u32.rs:107-146, 340-391
//   U32SubtractionGate  PX/frontend/uint/num/u32/gates/subtraction_u32.rs:101-135, 305-350
//   ComparisonGate      PX/frontend/uint/num/u32/gates/comparison.rs:118-195, 441-540
//   U32RangeCheckGate   PX/frontend/uint/num/u32/gates/range_check_u32.rs:69-91, 202-224
//   Poseidon            plonky2 0.2.1 hash_n_to_hash_no_pad (call sites PX/frontend/hash/poseidon/poseidon256.rs:68,
//                       PX/utils/poseidon/mod.rs:31-36,54-59)
// Gate kernels: one thread per row over plonky2's wire-major batch layout (wires[w*rows + r],
// constraints[c*rows + r]) so that every load and store of a warp is one contiguous 256-byte segment;
// the traffic is streamed (ld.global.cs / st.global.cs, nothing is reused).  ~1.8 KB of traffic against
// ~400 field multiplications per row puts the arithmetic gate near the HBM/ALU balance point.
#include "common.cuh"
#include "goldilocks.cuh"
#include "poseidon.cuh"

namespace bsx {

constexpr uint64_t GL_P = 0xFFFFFFFF00000001ULL;
constexpr uint64_t GL_EPS = 0xFFFFFFFFULL;

// all values canonical (< p)
__device__ __forceinline__ uint64_t gl_canon(uint64_t x) { return x >= GL_P ? x - GL_P : x; }
__device__ __forceinline__ uint64_t gl_add(uint64_t a, uint64_t b) {
    uint64_t s = a + b;
    return (s < a) ? s + GL_EPS : gl_canon(s);
}
__device__ __forceinline__ uint64_t gl_sub(uint64_t a, uint64_t b) {
    uint64_t d = a - b;
    return (a < b) ? d - GL_EPS : d;
}
__device__ __forceinline__ uint64_t gl_reduce128(uint64_t hi, uint64_t lo) {
    const uint64_t hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    uint64_t t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= GL_EPS;
    const uint64_t t1 = hi_lo * GL_EPS;
    uint64_t t2 = t0 + t1;
    if (t2 < t1) t2 += GL_EPS;
    return gl_canon(t2);
}
__device__ __forceinline__ uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_reduce128(__umul64hi(a, b), a * b); }
__device__ __forceinline__ uint64_t gl_dbl(uint64_t a) { return gl_add(a, a); }
__device__ __forceinline__ uint64_t gl_mul4(uint64_t a) { return gl_dbl(gl_dbl(a)); }
__device__ __forceinline__ uint64_t gl_inv(uint64_t a) {  // a^(p-2), p-2 = 0xFFFFFFFEFFFFFFFF
    uint64_t r = 1;
#pragma unroll 1
    for (int i = 63; i >= 0; i--) {
        r = gl_mul(r, r);
        if ((0xFFFFFFFEFFFFFFFFULL >> i) & 1) r = gl_mul(r, a);
    }
    return r;
}
// ---- lazy 128-bit forms used by the gate kernels (fewer reductions per constraint) ----
// weak reduction: any u64 representative of (hi*2^64 + lo) mod p (no final conditional subtraction)
__device__ __forceinline__ uint64_t gl_reduce128_weak(uint64_t hi, uint64_t lo) {
    const uint64_t hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    uint64_t t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= GL_EPS;
    const uint64_t t1 = hi_lo * GL_EPS;
    uint64_t t2 = t0 + t1;
    if (t2 < t1) t2 += GL_EPS;
    return t2;
}
// prod_{x < 4} (l - x) for canonical l, with two multiplications and two reductions:
//   u = l^2 - 3l  (computed as l^2 + 3(p - l) >= 0),   l(l-1)(l-2)(l-3) = u (u + 2) = u^2 + 2u
// u may be any 64-bit representative: u^2 + 2u <= 2^128 - 1 never overflows.
__device__ __forceinline__ uint64_t limb_product4(uint64_t l) {
    const uint64_t m = GL_P - l;                              // in (0, p]
    uint64_t lo = l * l, hi = __umul64hi(l, l);
    const uint64_t m3lo = m * 3, m3hi = __umul64hi(m, 3);     // 3m < 2^66
    lo += m3lo;
    hi += m3hi + (lo < m3lo ? 1 : 0);
    const uint64_t u = gl_reduce128_weak(hi, lo);
    lo = u * u;
    hi = __umul64hi(u, u);
    const uint64_t u2lo = u << 1, u2hi = u >> 63;
    lo += u2lo;
    hi += u2hi + (lo < u2lo ? 1 : 0);
    return gl_canon(gl_reduce128_weak(hi, lo));
}
// sum_j limb_j * 4^j accumulated as a plain 128-bit integer (16 canonical limbs: < 2^97), one reduction at the end
struct Horner4 {
    uint64_t hi = 0, lo = 0;
    __device__ __forceinline__ void push(uint64_t limb) {      // acc = 4*acc + limb  (limbs arrive most significant first)
        hi = (hi << 2) | (lo >> 62);
        lo <<= 2;
        lo += limb;
        hi += (lo < limb ? 1 : 0);
    }
    __device__ __forceinline__ uint64_t value() const { return gl_canon(gl_reduce128_weak(hi, lo)); }
};
__device__ __forceinline__ uint64_t limb_product(uint64_t l, uint32_t base) {
    uint64_t p = l;
    for (uint32_t x = 1; x < base; x++) p = gl_mul(p, gl_sub(l, x));
    return p;
}

struct RowIO {
    const uint64_t *__restrict__ wires;
    uint64_t *__restrict__ out;
    size_t rows, r;
    uint32_t c;
    __device__ __forceinline__ uint64_t raw(uint32_t col) const { return __ldcs(wires + (size_t)col * rows + r); }
    __device__ __forceinline__ uint64_t w(uint32_t col) const { return glf::canon(raw(col)); }
    __device__ __forceinline__ void put(uint64_t v) { __stcs(out + (size_t)(c++) * rows + r, v); }
    __device__ __forceinline__ void put_at(uint32_t col, uint64_t v) { __stcs(out + (size_t)col * rows + r, v); }
    // constraints col_top, col_top - 1, ...: the column pointer steps by one row stride (no per-access 64-bit multiply)
    struct Cursor {
        uint64_t *cp; size_t rows;
        __device__ __forceinline__ void put_down(int t, uint64_t v) const { __stcs(cp - (size_t)t * rows, v); }
    };
    __device__ __forceinline__ Cursor cursor(uint32_t col_top) { return Cursor{out + (size_t)col_top * rows + r, rows}; }
};

// The quotient form of the same evaluation (plonky2's compute_quotient_polys reduces the constraint terms of a point with
// powers of alpha before they ever reach memory): every constraint is folded into NA running sums  acc_a += alpha_a^c C_c
// instead of being stored, so a point costs its wire reads and NA words written.
template <int NA>
struct RowReduce {
    const uint64_t *__restrict__ wires;
    const uint64_t *apow;            // shared memory: apow[a * ncn + c] = alpha_a^c
    size_t rows, r;
    uint32_t c, ncn;
    uint64_t acc[NA];
    __device__ __forceinline__ uint64_t raw(uint32_t col) const { return __ldcs(wires + (size_t)col * rows + r); }
    __device__ __forceinline__ uint64_t w(uint32_t col) const { return glf::canon(raw(col)); }
    __device__ __forceinline__ void put_at(uint32_t col, uint64_t v) {
#pragma unroll
        for (int a = 0; a < NA; a++) acc[a] = glf::add(acc[a], glf::mul(v, apow[a * ncn + col]));
    }
    __device__ __forceinline__ void put(uint64_t v) { put_at(c++, v); }
    struct Cursor {
        RowReduce *io; uint32_t top;
        __device__ __forceinline__ void put_down(int t, uint64_t v) const { io->put_at(top - (uint32_t)t, v); }
    };
    __device__ __forceinline__ Cursor cursor(uint32_t col_top) { return Cursor{this, col_top}; }
};

// N 2-bit limbs (wires w0 .. w0+N-1, least significant first): their range-check products go to constraints
// c_top, c_top+1, ... in order of DECREASING significance (the reference iterates the limbs in reverse), and the
// return value is sum_t limb_t 4^t as a Radix4Sum.  All N loads are issued before the arithmetic.  The callers loop
// over groups of <= 8 limbs WITHOUT unrolling, so the hot code is one copy of this body (~600 instructions): fully
// unrolled the arithmetic gate was 58 KB of SASS and a quarter of its issue stalls were instruction-cache misses
// (profiles/r01i_gl_gate_eval_ncu_full.csv).
template <int N, class IO>
__device__ __forceinline__ glf::Radix4Sum limb_group(IO &io, uint32_t w0, uint32_t c_top) {
    // column pointers advance by one row stride per limb (no per-access 64-bit multiply)
    const uint64_t *wp = io.wires + (size_t)w0 * io.rows + io.r;
    const auto cur = io.cursor(c_top + N - 1);
    uint64_t limb[N];
#pragma unroll
    for (int t = 0; t < N; t++) limb[t] = __ldcs(wp + (size_t)t * io.rows);   // raw: limb_product4 and the sum take any representative
    glf::Radix4Sum s;
#pragma unroll
    for (int t = 0; t < N; t++) {
        cur.put_down(t, glf::limb_product4(limb[t]));
        s.add(limb[t], t);
    }
    return s;
}

// arithmetic_u32.rs:280-349: per op  [hi_not_max * out_lo, out_hi 2^32 + out_lo - (m0 m1 + addend),
//   32 limb products (limb 31 first), low16 - out_lo, high16 - out_hi]
template <class IO>
__device__ __forceinline__ void eval_arithmetic(IO &io, uint32_t num_ops) {
#pragma unroll 1
    for (uint32_t i = 0; i < num_ops; i++) {
        const uint32_t cb = 36 * i, wl = 6 * num_ops + 32 * i;
        const uint64_t m0 = io.w(6 * i), m1 = io.w(6 * i + 1), addend = io.w(6 * i + 2);
        const uint64_t out_lo = io.w(6 * i + 3), out_hi = io.w(6 * i + 4), inverse = io.w(6 * i + 5);
        const uint64_t computed = glf::add(glf::mul(m0, m1), addend);
        const uint64_t hi_not_max = glf::sub(glf::mul(inverse, glf::sub(GL_EPS, out_hi)), 1);
        io.put_at(cb, glf::mul(hi_not_max, out_lo));
        io.put_at(cb + 1, glf::sub(glf::add(glf::mul(out_hi, 1ULL << 32), out_lo), computed));
        glf::Radix4Sum top;
#pragma unroll 1
        for (int g = 3; g >= 0; g--) {               // limbs 8g .. 8g+7; g = 3, 2 build the high word, g = 1, 0 the low word
            glf::Radix4Sum s = limb_group<8>(io, wl + 8 * g, cb + 2 + 8 * (3 - g));
            if (g & 1) top = s;
            else {
                s.append_high8(top);
                io.put_at(cb + 34 + (g >> 1), glf::sub(s.value(), g ? out_hi : out_lo));
            }
        }
    }
}

// add_many_u32.rs:107-146: per op  [out_carry 2^32 + out_res - sum(addends, carry_in),
//   19 limb products (limb 18 first: 3 carry limbs, then 16 result limbs), res16 - out_res, carry3 - out_carry]
template <class IO>
__device__ __forceinline__ void eval_add_many(IO &io, uint32_t na, uint32_t num_ops) {
#pragma unroll 1
    for (uint32_t i = 0; i < num_ops; i++) {
        const uint32_t b = (na + 3) * i, cb = 22 * i, wl = (na + 3) * num_ops + 19 * i;
        uint64_t computed = 0;
#pragma unroll 1
        for (uint32_t j = 0; j <= na; j++) computed = glf::add(computed, io.w(b + j));
        const uint64_t out_res = io.w(b + na + 1), out_carry = io.w(b + na + 2);
        io.put_at(cb, glf::sub(glf::add(glf::mul(out_carry, 1ULL << 32), out_res), computed));
        const glf::Radix4Sum carry = limb_group<3>(io, wl + 16, cb + 1);
        glf::Radix4Sum top;
#pragma unroll 1
        for (int g = 1; g >= 0; g--) {
            glf::Radix4Sum s = limb_group<8>(io, wl + 8 * g, cb + 4 + 8 * (1 - g));
            if (g) top = s;
            else {
                s.append_high8(top);
                io.put_at(cb + 20, glf::sub(s.value(), out_res));
            }
        }
        io.put_at(cb + 21, glf::sub(carry.value(), out_carry));
    }
}

// subtraction_u32.rs:101-135: per op  [out_res - (x - y - borrow_in + 2^32 out_borrow), 16 limb products (limb 15
//   first), limbs16 - out_res, out_borrow (1 - out_borrow)]
template <class IO>
__device__ __forceinline__ void eval_subtraction(IO &io, uint32_t num_ops) {
#pragma unroll 1
    for (uint32_t i = 0; i < num_ops; i++) {
        const uint32_t cb = 19 * i, wl = 5 * num_ops + 16 * i;
        const uint64_t x = io.w(5 * i), y = io.w(5 * i + 1), bin = io.w(5 * i + 2), out_res = io.w(5 * i + 3),
                       out_b = io.w(5 * i + 4);
        const uint64_t initial = glf::sub(glf::sub(x, y), bin);
        io.put_at(cb, glf::sub(out_res, glf::add(initial, glf::mul(1ULL << 32, out_b))));
        glf::Radix4Sum top;
#pragma unroll 1
        for (int g = 1; g >= 0; g--) {
            glf::Radix4Sum s = limb_group<8>(io, wl + 8 * g, cb + 1 + 8 * (1 - g));
            if (g) top = s;
            else {
                s.append_high8(top);
                io.put_at(cb + 17, glf::sub(s.value(), out_res));
            }
        }
        io.put_at(cb + 18, glf::mul(out_b, glf::sub(1, out_b)));
    }
}

// prod_{x < base} (l - x) for canonical l; base = 4 (2-bit chunks, the shape the circuits use) takes the two-product form
__device__ __forceinline__ uint64_t chunk_product(uint64_t l, uint32_t base) {
    if (base == 4) return glf::limb_product4(l);
    uint64_t p = l;
    for (uint32_t x = 1; x < base; x++) p = glf::mul(p, glf::sub(l, x));
    return p;
}

// comparison.rs:118-195
template <class IO>
__device__ __forceinline__ void eval_comparison(IO &io, uint32_t num_bits, uint32_t nc) {
    const uint32_t cb = (num_bits + nc - 1) / nc;
    const uint64_t chunk_size = 1ULL << cb;
    // first / second input recombined from their chunks: sum chunk_i 2^(cb i) as Horner steps (chunks are canonical,
    // 2^cb <= 2^16, so each step is one 128-bit multiply-add reduced weakly)
    uint64_t fc = 0, sc = 0;
#pragma unroll 1
    for (int i = (int)nc - 1; i >= 0; i--) {
        fc = glf::add(glf::mul(fc, chunk_size), io.w(4 + i));
        sc = glf::add(glf::mul(sc, chunk_size), io.w(4 + nc + i));
    }
    io.put(glf::sub(fc, io.w(0)));
    io.put(glf::sub(sc, io.w(1)));
    uint64_t msd_so_far = 0;
#pragma unroll 1
    for (uint32_t i = 0; i < nc; i++) {          // (unrolling four chunks per pass measured slower: 0.47 vs 0.44 ms)
        const uint64_t f = io.w(4 + i), s = io.w(4 + nc + i);
        io.put(chunk_product(f, (uint32_t)chunk_size));
        io.put(chunk_product(s, (uint32_t)chunk_size));
        const uint64_t diff = glf::sub(s, f), dummy = io.w(4 + 2 * nc + i), eq = io.w(4 + 3 * nc + i);
        io.put(glf::sub(glf::mul(diff, dummy), glf::sub(1, eq)));
        io.put(glf::mul(eq, diff));
        const uint64_t inter = io.w(4 + 4 * nc + i);
        io.put(glf::sub(inter, glf::mul(eq, msd_so_far)));
        msd_so_far = glf::add(inter, glf::mul(glf::sub(1, eq), diff));
    }
    const uint64_t msd = io.w(3);
    io.put(glf::sub(msd, msd_so_far));
#pragma unroll 1
    for (uint32_t i = 0; i <= cb; i++) {
        const uint64_t bit = io.w(4 + 5 * nc + i);
        io.put(glf::mul(bit, glf::sub(1, bit)));
    }
    uint64_t bits = 0;
#pragma unroll 1
    for (int i = (int)cb; i >= 0; i--) bits = glf::add(glf::add(bits, bits), io.w(4 + 5 * nc + i));
    io.put(glf::sub(glf::add(chunk_size, msd), bits));
    io.put(glf::sub(io.w(2), io.w(4 + 5 * nc + cb)));
}

// range_check_u32.rs:69-91: per value  [aux16 - value, then the 16 limb products in INCREASING limb order]
template <class IO>
__device__ __forceinline__ void eval_range_check(IO &io, uint32_t nl) {
#pragma unroll 1
    for (uint32_t i = 0; i < nl; i++) {
        const uint32_t cb = 17 * i;
        glf::Radix4Sum top, sum;
#pragma unroll 1
        for (int g = 1; g >= 0; g--) {
            uint64_t aux[8];
#pragma unroll
            for (int t = 0; t < 8; t++) aux[t] = io.raw(nl + 16 * i + 8 * g + t);
            glf::Radix4Sum s;
#pragma unroll
            for (int t = 0; t < 8; t++) {
                io.put_at(cb + 1 + 8 * g + t, glf::limb_product4(aux[t]));
                s.add(aux[t], t);
            }
            if (g) top = s;
            else { s.append_high8(top); sum = s; }
        }
        io.put_at(cb, glf::sub(sum.value(), io.w(i)));
    }
}

template <int GATE>
__global__ void __launch_bounds__(128) gl_gate_eval_kernel(uint32_t p0, uint32_t p1, const uint64_t *__restrict__ wires,
                                                           uint32_t rows, uint64_t *__restrict__ constraints) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    RowIO io{wires, constraints, rows, r, 0};
    if (GATE == BSX_GATE_U32_ARITHMETIC) eval_arithmetic(io, p0);
    else if (GATE == BSX_GATE_U32_ADD_MANY) eval_add_many(io, p0, p1);
    else if (GATE == BSX_GATE_U32_SUBTRACTION) eval_subtraction(io, p0);
    else if (GATE == BSX_GATE_U32_COMPARISON) eval_comparison(io, p0, p1);
    else eval_range_check(io, p0);
}

// Quotient-style evaluation over a low-degree extension (plonky2 compute_quotient_polys, un-vendored; call site
// PX/backend/circuit/build.rs:69-75): for every point x of the coset LDE, out[a][x] = zh_inv(x) * sum_c alpha_a^c C_c(x).
// `wires` are the LDE values poly-major (wire w of point r at wires[w * rows + r], the layout bsx_gl_lde_dev writes),
// zh_inv has one entry per block of 2^log_block points (Z_H = x^n - 1 takes 2^rate_bits values on the coset).
template <int GATE, int NA>
__global__ void __launch_bounds__(128) gl_gate_quotient_kernel(uint32_t p0, uint32_t p1, const uint64_t *__restrict__ wires, uint32_t rows,
                                                               const uint64_t *__restrict__ alpha_pows, uint32_t ncn,
                                                               const uint64_t *__restrict__ zh_inv, uint32_t log_block,
                                                               uint64_t *__restrict__ out) {
    extern __shared__ uint64_t s_apow[];
    for (uint32_t k = threadIdx.x; k < NA * ncn; k += blockDim.x) s_apow[k] = alpha_pows[k];
    __syncthreads();
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    RowReduce<NA> io{wires, s_apow, rows, r, 0, ncn, {}};
#pragma unroll
    for (int a = 0; a < NA; a++) io.acc[a] = 0;
    if (GATE == BSX_GATE_U32_ARITHMETIC) eval_arithmetic(io, p0);
    else if (GATE == BSX_GATE_U32_ADD_MANY) eval_add_many(io, p0, p1);
    else if (GATE == BSX_GATE_U32_SUBTRACTION) eval_subtraction(io, p0);
    else if (GATE == BSX_GATE_U32_COMPARISON) eval_comparison(io, p0, p1);
    else eval_range_check(io, p0);
    const uint64_t zi = zh_inv[r >> log_block];
#pragma unroll
    for (int a = 0; a < NA; a++) __stcs(out + (size_t)a * rows + r, glf::mul(io.acc[a], zi));
}

// ---- witness generators (SimpleGenerator::run_once): fill the dependent wires of each row in place ----
struct RowRW {
    uint64_t *wires;
    size_t rows, r;
    __device__ __forceinline__ uint64_t get(uint32_t col) const { return gl_canon(wires[(size_t)col * rows + r]); }
    __device__ __forceinline__ void set(uint32_t col, uint64_t v) { wires[(size_t)col * rows + r] = v; }
    __device__ __forceinline__ void split(uint64_t v, uint32_t n, uint32_t bits, uint32_t col0) {
        for (uint32_t j = 0; j < n; j++) { set(col0 + j, v & ((1ULL << bits) - 1)); v >>= bits; }
    }
};

__global__ void __launch_bounds__(128) gl_gate_witness_kernel(uint32_t gate, uint32_t p0, uint32_t p1, uint64_t *wires,
                                                              uint32_t rows) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    RowRW w{wires, rows, r};
    if (gate == BSX_GATE_U32_ARITHMETIC) {
        for (uint32_t i = 0; i < p0; i++) {
            const uint64_t out = gl_add(gl_mul(w.get(6 * i), w.get(6 * i + 1)), w.get(6 * i + 2));
            const uint64_t hi = out >> 32, lo = out & GL_EPS;
            w.set(6 * i + 3, lo);
            w.set(6 * i + 4, hi);
            const uint64_t diff = GL_EPS - hi;
            w.set(6 * i + 5, diff == 0 ? 0 : gl_inv(diff));
            w.split(out, 32, 2, 6 * p0 + 32 * i);
        }
    } else if (gate == BSX_GATE_U32_ADD_MANY) {
        const uint32_t na = p0;
        for (uint32_t i = 0; i < p1; i++) {
            uint64_t out = 0;
            for (uint32_t j = 0; j <= na; j++) out = gl_add(out, w.get((na + 3) * i + j));
            const uint64_t carry = out >> 32, res = out & GL_EPS;
            w.set((na + 3) * i + na + 1, res);
            w.set((na + 3) * i + na + 2, carry);
            w.split(res, 16, 2, (na + 3) * p1 + 19 * i);
            w.split(carry, 3, 2, (na + 3) * p1 + 19 * i + 16);
        }
    } else if (gate == BSX_GATE_U32_SUBTRACTION) {
        for (uint32_t i = 0; i < p0; i++) {
            const uint64_t initial = gl_sub(gl_sub(w.get(5 * i), w.get(5 * i + 1)), w.get(5 * i + 2));
            const uint64_t borrow = initial > (1ULL << 32) ? 1 : 0;  // subtraction_u32.rs:319 (strict >)
            const uint64_t res = gl_add(initial, gl_mul(1ULL << 32, borrow));
            w.set(5 * i + 3, res);
            w.set(5 * i + 4, borrow);
            w.split(res, 16, 2, 5 * p0 + 16 * i);
        }
    } else if (gate == BSX_GATE_U32_COMPARISON) {
        const uint32_t nc = p1, cb = (p0 + p1 - 1) / p1;
        const uint64_t a = w.get(0), b = w.get(1);
        uint64_t msd = 0;
        w.set(2, a <= b ? 1 : 0);
        for (uint32_t i = 0; i < nc; i++) {
            const uint32_t sh = cb * i;
            const uint64_t f = sh < 64 ? (a >> sh) & ((1ULL << cb) - 1) : 0, s = sh < 64 ? (b >> sh) & ((1ULL << cb) - 1) : 0;
            w.set(4 + i, f);
            w.set(4 + nc + i, s);
            w.set(4 + 2 * nc + i, f == s ? 1 : gl_inv(gl_sub(s, f)));
            w.set(4 + 3 * nc + i, f == s ? 1 : 0);
            if (f != s) { msd = gl_sub(s, f); w.set(4 + 4 * nc + i, 0); }
            else w.set(4 + 4 * nc + i, msd);
        }
        w.set(3, msd);
        uint64_t t = gl_add(1ULL << cb, msd);
        for (uint32_t i = 0; i <= cb; i++) { w.set(4 + 5 * nc + i, t & 1); t >>= 1; }
    } else {
        for (uint32_t i = 0; i < p0; i++) w.split((uint32_t)w.get(i), 16, 2, p0 + 16 * i);
    }
}

// hash_n_to_hash_no_pad: rate 8, overwrite mode; one thread per hash
__global__ void __launch_bounds__(128) gl_poseidon_batch_kernel(const uint64_t *__restrict__ in, const uint32_t *__restrict__ offsets,
                                                                uint32_t n, uint64_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = offsets[i], e = offsets[i + 1];
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = 0;
#pragma unroll 1
    for (uint32_t p = b; p < e; p += 8) {
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (p + k < e) s[k] = in[p + k];          // any representative
        poseidon_permute(s);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) out[4 * (size_t)i + k] = glf::canon(s[k]);
}

}  // namespace bsx

using namespace bsx;

extern "C" uint32_t bsx_gate_num_wires(uint32_t gate, uint32_t p0, uint32_t p1) {
    switch (gate) {
        case BSX_GATE_U32_ARITHMETIC: return p0 * 38;
        case BSX_GATE_U32_ADD_MANY: return p1 * (p0 + 3 + 19);
        case BSX_GATE_U32_SUBTRACTION: return p0 * 21;
        case BSX_GATE_U32_COMPARISON: return p1 ? 4 + 5 * p1 + (p0 + p1 - 1) / p1 + 1 : 0;
        case BSX_GATE_U32_RANGE_CHECK: return p0 * 17;
    }
    return 0;
}
extern "C" uint32_t bsx_gate_num_constraints(uint32_t gate, uint32_t p0, uint32_t p1) {
    switch (gate) {
        case BSX_GATE_U32_ARITHMETIC: return p0 * 36;
        case BSX_GATE_U32_ADD_MANY: return p1 * 22;
        case BSX_GATE_U32_SUBTRACTION: return p0 * 19;
        case BSX_GATE_U32_COMPARISON: return p1 ? 6 + 5 * p1 + (p0 + p1 - 1) / p1 : 0;
        case BSX_GATE_U32_RANGE_CHECK: return p0 * 17;
    }
    return 0;
}

static int gate_args_ok(bsx_ctx *ctx, uint32_t gate, uint32_t p0, uint32_t p1) {
    BSX_REQUIRE(ctx, gate <= BSX_GATE_U32_RANGE_CHECK && p0 >= 1);
    BSX_REQUIRE(ctx, gate != BSX_GATE_U32_ADD_MANY || (p0 <= 64 && p1 >= 1));
    BSX_REQUIRE(ctx, gate != BSX_GATE_U32_COMPARISON || (p1 >= 1 && p0 <= 63 && (p0 + p1 - 1) / p1 <= 16));
    return BSX_OK;
}

extern "C" int bsx_gl_gate_eval_dev(bsx_ctx *ctx, void *stream, uint32_t gate, uint32_t p0, uint32_t p1, const uint64_t *wires,
                                    uint32_t rows, uint64_t *constraints) {
    BSX_REQUIRE(ctx, ctx && wires && constraints);
    int rc = gate_args_ok(ctx, gate, p0, p1);
    if (rc) return rc;
    if (rows == 0) return BSX_OK;
    const uint32_t blocks = (rows + 127) / 128;
    cudaStream_t st = (cudaStream_t)stream;
    switch (gate) {
        case BSX_GATE_U32_ARITHMETIC: gl_gate_eval_kernel<BSX_GATE_U32_ARITHMETIC><<<blocks, 128, 0, st>>>(p0, p1, wires, rows, constraints); break;
        case BSX_GATE_U32_ADD_MANY: gl_gate_eval_kernel<BSX_GATE_U32_ADD_MANY><<<blocks, 128, 0, st>>>(p0, p1, wires, rows, constraints); break;
        case BSX_GATE_U32_SUBTRACTION: gl_gate_eval_kernel<BSX_GATE_U32_SUBTRACTION><<<blocks, 128, 0, st>>>(p0, p1, wires, rows, constraints); break;
        case BSX_GATE_U32_COMPARISON: gl_gate_eval_kernel<BSX_GATE_U32_COMPARISON><<<blocks, 128, 0, st>>>(p0, p1, wires, rows, constraints); break;
        default: gl_gate_eval_kernel<BSX_GATE_U32_RANGE_CHECK><<<blocks, 128, 0, st>>>(p0, p1, wires, rows, constraints); break;
    }
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

template <int NA>
static int launch_quotient(bsx_ctx *ctx, cudaStream_t st, uint32_t gate, uint32_t p0, uint32_t p1, const uint64_t *wires, uint32_t rows,
                           const uint64_t *alpha_pows, uint32_t ncn, const uint64_t *zh_inv, uint32_t log_block, uint64_t *out) {
    const unsigned grid = (rows + 127) / 128;
    const size_t smem = sizeof(uint64_t) * NA * ncn;
    switch (gate) {
        case BSX_GATE_U32_ARITHMETIC: gl_gate_quotient_kernel<BSX_GATE_U32_ARITHMETIC, NA><<<grid, 128, smem, st>>>(p0, p1, wires, rows, alpha_pows, ncn, zh_inv, log_block, out); break;
        case BSX_GATE_U32_ADD_MANY: gl_gate_quotient_kernel<BSX_GATE_U32_ADD_MANY, NA><<<grid, 128, smem, st>>>(p0, p1, wires, rows, alpha_pows, ncn, zh_inv, log_block, out); break;
        case BSX_GATE_U32_SUBTRACTION: gl_gate_quotient_kernel<BSX_GATE_U32_SUBTRACTION, NA><<<grid, 128, smem, st>>>(p0, p1, wires, rows, alpha_pows, ncn, zh_inv, log_block, out); break;
        case BSX_GATE_U32_COMPARISON: gl_gate_quotient_kernel<BSX_GATE_U32_COMPARISON, NA><<<grid, 128, smem, st>>>(p0, p1, wires, rows, alpha_pows, ncn, zh_inv, log_block, out); break;
        default: gl_gate_quotient_kernel<BSX_GATE_U32_RANGE_CHECK, NA><<<grid, 128, smem, st>>>(p0, p1, wires, rows, alpha_pows, ncn, zh_inv, log_block, out); break;
    }
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

// Quotient-style constraint evaluation over an LDE (see gl_gate_quotient_kernel).  alpha_pows: device, n_alphas x
// num_constraints (alpha_a^c); zh_inv: device, rows >> log_block entries; out: n_alphas x rows.
extern "C" int bsx_gl_gate_quotient_dev(bsx_ctx *ctx, void *stream, uint32_t gate, uint32_t p0, uint32_t p1, const uint64_t *wires,
                                        uint32_t rows, const uint64_t *alpha_pows, uint32_t n_alphas, const uint64_t *zh_inv,
                                        uint32_t log_block, uint64_t *out) {
    BSX_REQUIRE(ctx, ctx && wires && alpha_pows && zh_inv && out && gate <= BSX_GATE_U32_RANGE_CHECK && (n_alphas == 1 || n_alphas == 2));
    BSX_REQUIRE(ctx, log_block <= 31);
    if (rows == 0) return BSX_OK;
    const uint32_t ncn = bsx_gate_num_constraints(gate, p0, p1);
    BSX_REQUIRE(ctx, ncn >= 1 && sizeof(uint64_t) * 2 * ncn <= 48 * 1024);
    return n_alphas == 1 ? launch_quotient<1>(ctx, (cudaStream_t)stream, gate, p0, p1, wires, rows, alpha_pows, ncn, zh_inv, log_block, out)
                         : launch_quotient<2>(ctx, (cudaStream_t)stream, gate, p0, p1, wires, rows, alpha_pows, ncn, zh_inv, log_block, out);
}

extern "C" int bsx_gl_gate_witness_dev(bsx_ctx *ctx, void *stream, uint32_t gate, uint32_t p0, uint32_t p1, uint64_t *wires,
                                       uint32_t rows) {
    BSX_REQUIRE(ctx, ctx && wires);
    int rc = gate_args_ok(ctx, gate, p0, p1);
    if (rc) return rc;
    if (rows == 0) return BSX_OK;
    gl_gate_witness_kernel<<<(rows + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gate, p0, p1, wires, rows);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

extern "C" int bsx_gl_poseidon_batch_dev(bsx_ctx *ctx, void *stream, const uint64_t *in, const uint32_t *offsets, uint32_t n,
                                         uint64_t *out) {
    BSX_REQUIRE(ctx, ctx && offsets && out);
    if (n == 0) return BSX_OK;
    gl_poseidon_batch_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(in, offsets, n, out);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

// ---- host-buffer entry points ----
extern "C" int bsx_gl_gate_eval(bsx_ctx *ctx, uint32_t gate, uint32_t p0, uint32_t p1, const uint64_t *wires, uint32_t rows,
                                uint64_t *constraints) {
    BSX_REQUIRE(ctx, ctx && wires && constraints);
    int rc = gate_args_ok(ctx, gate, p0, p1);
    if (rc) return rc;
    if (rows == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t sw = 8 * (size_t)bsx_gate_num_wires(gate, p0, p1) * rows, sc = 8 * (size_t)bsx_gate_num_constraints(gate, p0, p1) * rows;
    rc = ws_begin(ctx, ws_size(sw) + ws_size(sc));
    if (rc) return rc;
    uint64_t *d_w = ws_take<uint64_t>(ctx, sw / 8), *d_c = ws_take<uint64_t>(ctx, sc / 8);
    BSX_CUDA(ctx, cudaMemcpyAsync(d_w, wires, sw, cudaMemcpyHostToDevice, ctx->stream));
    rc = bsx_gl_gate_eval_dev(ctx, ctx->stream, gate, p0, p1, d_w, rows, d_c);
    if (rc) return rc;
    BSX_CUDA(ctx, cudaMemcpyAsync(constraints, d_c, sc, cudaMemcpyDeviceToHost, ctx->stream));
    BSX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BSX_OK;
}

extern "C" int bsx_gl_gate_witness(bsx_ctx *ctx, uint32_t gate, uint32_t p0, uint32_t p1, uint64_t *wires, uint32_t rows) {
    BSX_REQUIRE(ctx, ctx && wires);
    int rc = gate_args_ok(ctx, gate, p0, p1);
    if (rc) return rc;
    if (rows == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t sw = 8 * (size_t)bsx_gate_num_wires(gate, p0, p1) * rows;
    rc = ws_begin(ctx, ws_size(sw));
    if (rc) return rc;
    uint64_t *d_w = ws_take<uint64_t>(ctx, sw / 8);
    BSX_CUDA(ctx, cudaMemcpyAsync(d_w, wires, sw, cudaMemcpyHostToDevice, ctx->stream));
    rc = bsx_gl_gate_witness_dev(ctx, ctx->stream, gate, p0, p1, d_w, rows);
    if (rc) return rc;
    BSX_CUDA(ctx, cudaMemcpyAsync(wires, d_w, sw, cudaMemcpyDeviceToHost, ctx->stream));
    BSX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BSX_OK;
}

extern "C" int bsx_gl_poseidon_batch(bsx_ctx *ctx, const uint64_t *in, const uint32_t *offsets, uint32_t n, uint64_t *out) {
    BSX_REQUIRE(ctx, ctx && offsets && out);
    if (n == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t total = offsets[n];
    BSX_REQUIRE(ctx, total == 0 || in);
    int rc = ws_begin(ctx, ws_size(8 * total + 8) + ws_size(4 * (size_t)(n + 1)) + ws_size(32 * (size_t)n));
    if (rc) return rc;
    uint64_t *d_in = ws_take<uint64_t>(ctx, total + 1);
    uint32_t *d_off = ws_take<uint32_t>(ctx, n + 1);
    uint64_t *d_out = ws_take<uint64_t>(ctx, 4 * (size_t)n);
    if (total) BSX_CUDA(ctx, cudaMemcpyAsync(d_in, in, 8 * total, cudaMemcpyHostToDevice, ctx->stream));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_off, offsets, 4 * (size_t)(n + 1), cudaMemcpyHostToDevice, ctx->stream));
    rc = bsx_gl_poseidon_batch_dev(ctx, ctx->stream, d_in, d_off, n, d_out);
    if (rc) return rc;
    BSX_CUDA(ctx, cudaMemcpyAsync(out, d_out, 32 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    BSX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BSX_OK;
}

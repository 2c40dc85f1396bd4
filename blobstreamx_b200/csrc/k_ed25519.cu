// K4+K5: batched Ed25519 witness generation.  Two paths, chosen by batch size (BSX_ED_QUAD_MAX = 16 384; BSX_ED_MODE forces one):
//
// (a) up to 16 384 signatures -- the latency path (one proof = 100 signatures): three kernels, because the parallel width of
//     the work changes along a signature:
//   ed25519_prep_kernel    one thread per POINT (2 per signature): SHA-512, h = digest div/rem l, s < l, and the
//                          decompression of A or R (a serial chain of ~265 squarings -- nothing to split)
//   ed25519_quad_kernel    FOUR lanes per signature, one extended coordinate (X, Y, Z, T) each: the four independent
//                          field multiplications / squarings of every point doubling, addition and p1p1 -> p3
//                          conversion run side by side and the operands move between the lanes with warp shuffles;
//                          s*G (8-bit windows), the 16-entry table of A, h*A and R + h*A
//   ed25519_finish_kernel  one thread per signature: one inversion for the three Z's, affine bytes, flags
//     The latency of one signature drops from ~3 700 dependent field operations to ~265 + ~750 + ~280.  r01l: 100 signatures
//     0.57 ms instead of 1.26, 10 000: 0.96 instead of 1.30; r03a/b, all three kernels on the FP64 pipe (441 / 652 cycles per
//     squaring / multiplication with one warp per sub-partition against 570 / 772): 0.46 and 0.79 ms.  The quad kernel
//     executes ~1.5x the instructions per signature (shuffles, re-carries, sign selections), hence the size threshold.
//
// (b) above it -- the throughput path: one thread per signature (ed25519_batch_kernel: decompression of A, s*G from the
//     8-bit window table, h*A by 252 doublings + signed 4-bit windows, R checked against sG - hA before a decompression is
//     spent on it, one inversion), in the register budget whose wave the batch fills.  When public keys repeat in the batch
//     (one validator set signing many ranges) the key-table kernels below run first and ed25519_keyed_kernel takes h*A from
//     per-key window tables instead (no doublings, A decompressed once per key); which of the two kernels does the work is
//     decided on the device, the other returns at once.
//
// Replaces, per signature, the CPU hints of curta_eddsa_verify_sigs (PX/frontend/ecc/curve25519/
// ed25519/eddsa.rs:161-203): HashDigestHint<SHA512> (PX/frontend/hash/sha/sha512/curta.rs:103-111),
// BigUintDivRemGenerator (PX/frontend/uint/num/biguint/mod.rs:451-488) and the seven EcOpResultHint
// calls (PX/frontend/ecc/curve25519/curta/result_hint.rs:21-50).  Inactive lanes run on the DUMMY
// triple exactly like curta_eddsa_verify_sigs_conditional (eddsa.rs:72-127).
#include "common.cuh"
#include "ed25519.cuh"
#include "sha512.cuh"

#include <stdlib.h>

namespace bsx {

using namespace ed;

__device__ __constant__ uint8_t DUMMY_PK[32] = {138, 136, 227, 221, 116, 9, 241, 149, 253, 82, 219, 45, 60, 186, 93, 114,
                                                 202, 103, 9, 191, 29, 148, 18, 27, 243, 116, 136, 1, 180, 15, 111, 92};
__device__ __constant__ uint8_t DUMMY_SIG[64] = {55, 20, 104, 158, 84, 120, 194, 17, 6, 237, 157, 164, 85, 88, 158, 137,
                                                  187, 119, 187, 240, 159, 73, 80, 63, 133, 162, 74, 91, 48, 53, 6, 138,
                                                  1, 41, 22, 121, 249, 46, 198, 145, 155, 102, 3, 210, 168, 135, 173, 55,
                                                  252, 72, 45, 126, 169, 178, 191, 7, 153, 67, 112, 90, 150, 33, 140, 7};

__global__ void __launch_bounds__(64) ed25519_base_table_kernel(ge_niels_slot *table) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BSX_ED_BASE_WINDOWS * BSX_ED_BASE_ENTRIES) return;
    ge_niels_store(table + i, ge_base_table_entry(i / BSX_ED_BASE_ENTRIES, i % BSX_ED_BASE_ENTRIES + 1));
}

struct EdIn {
    const uint8_t *pks, *sigs, *msgs, *lens, *active;   // lens: u32 LE at lens + i*len_stride (NULL = msg_max)
    uint32_t pk_stride, sig_stride, msg_stride, msg_max, len_stride, active_stride;
};

// ---- per-key tables (ed25519.cuh, "per-key window tables"): which signatures share a public key is found on the device ----
// An open-addressing table of BSX_ED_KEY_SLOTS slots holds, per distinct key, the index of the first signature that
// carries it; a batch with more than BSX_ED_KEY_MAX distinct keys (or a probe sequence that runs too long) sets the
// overflow word and the whole batch takes the general path.  No host round trip: every later kernel reads the verdict.
#define BSX_ED_KEY_SLOTS 4096
#define BSX_ED_KEY_MAX 1024
#define BSX_ED_KEY_PROBES 64
#define BSX_ED_KEY_MIN_USE 16        // tables are used when a key serves at least this many signatures on average
struct EdKeys {
    int32_t *slots;        // [SLOTS] -1 or the first signature with this key
    int32_t *slot_id;      // [SLOTS] dense key id (assigned by the bases kernel)
    int32_t *state;        // [0] distinct keys  [1] overflow  [2] next dense id
    int32_t *key_slot;     // [n] slot of each signature's key
    uint8_t *recs;         // [KEY_MAX] BSX_ED_KEYREC_BYTES
    double *bases;         // [KEY_MAX][WINDOWS][20]
    double *tab;           // [KEY_MAX][WINDOWS][ENTRIES][20]
    int32_t force;         // ED_KEYTAB = 1: use the tables whenever the keys fit
    int32_t pair;          // table path: two signatures per thread sharing one inversion (large batches only)
};
__device__ __forceinline__ bool ed_keys_in_use(const EdKeys &k, uint32_t n) {
    return k.state != nullptr && k.state[1] == 0 && (k.force || (uint64_t)k.state[0] * BSX_ED_KEY_MIN_USE <= n);
}
__device__ __forceinline__ void ed_effective_pk(const EdIn &in, uint32_t i, uint32_t w[8]) {
    const bool on = !in.active || in.active[(size_t)in.active_stride * i];
    const uint8_t *p = on ? in.pks + (size_t)in.pk_stride * i : DUMMY_PK;
#pragma unroll
    for (int k = 0; k < 8; k++)
        w[k] = (uint32_t)p[4 * k] | ((uint32_t)p[4 * k + 1] << 8) | ((uint32_t)p[4 * k + 2] << 16) | ((uint32_t)p[4 * k + 3] << 24);
}
__global__ void __launch_bounds__(128) ed25519_key_assign_kernel(uint32_t n, EdIn in, EdKeys keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[8];
    ed_effective_pk(in, i, w);
    uint32_t h = 0x811c9dc5u;
#pragma unroll
    for (int k = 0; k < 8; k++) h = (h ^ w[k]) * 0x01000193u + (h >> 15);
    h ^= h >> 16;
    int32_t found = -1;
    for (int probe = 0; probe < BSX_ED_KEY_PROBES; probe++) {
        const uint32_t slot = (h + probe) & (BSX_ED_KEY_SLOTS - 1);
        int32_t cur = keys.slots[slot];
        if (cur < 0) {
            cur = atomicCAS(&keys.slots[slot], -1, (int32_t)i);
            if (cur < 0) {                                   // this signature is the key's first
                if (atomicAdd(&keys.state[0], 1) >= BSX_ED_KEY_MAX) keys.state[1] = 1;
                found = (int32_t)slot;
                break;
            }
        }
        uint32_t o[8];
        ed_effective_pk(in, (uint32_t)cur, o);
        bool same = true;
#pragma unroll
        for (int k = 0; k < 8; k++) same = same && o[k] == w[k];
        if (same) { found = (int32_t)slot; break; }
    }
    if (found < 0) keys.state[1] = 1;
    keys.key_slot[i] = found;
}
// one thread per occupied slot: the key's record and its 64 window bases (a dependent chain: ~2 100 field operations)
__global__ void __launch_bounds__(32) ed25519_key_bases_kernel(uint32_t n, EdIn in, EdKeys keys) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= BSX_ED_KEY_SLOTS || !ed_keys_in_use(keys, n)) return;
    const int32_t first = keys.slots[s];
    if (first < 0) return;
    const int32_t id = atomicAdd(&keys.state[2], 1);
    keys.slot_id[s] = id;
    uint32_t w[8];
    ed_effective_pk(in, (uint32_t)first, w);
    uint8_t pk[32];
#pragma unroll
    for (int k = 0; k < 32; k++) pk[k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
    edd::ed25519_key_bases(pk, keys.recs + (size_t)BSX_ED_KEYREC_BYTES * id, keys.bases + (size_t)id * BSX_ED_KEY_WINDOWS * 20);
}
// one CTA per key, one thread per window: the multiples 1..8 of the window base in addend form
__global__ void __launch_bounds__(BSX_ED_KEY_WINDOWS) ed25519_key_table_kernel(uint32_t n, EdKeys keys) {
    const uint32_t id = blockIdx.x, w = threadIdx.x;
    if (!ed_keys_in_use(keys, n) || (int32_t)id >= keys.state[0]) return;
    const size_t e = (size_t)id * BSX_ED_KEY_WINDOWS + w;
    edd::ed25519_key_window(keys.bases + e * 20, keys.tab + e * BSX_ED_KEY_ENTRIES * 20);
}

// FP64: the field arithmetic of the whole signature on the FP64 pipe (fe51d.cuh) instead of IMAD.WIDE
template <bool INL, bool FP64>
__device__ __forceinline__ void ed25519_batch_body(uint32_t n, const EdIn &in, const ge_niels_slot *__restrict__ table, uint8_t *__restrict__ out,
                                                   const EdKeys &keys);

template <int MIN_CTAS, bool INL, bool FP64 = false>
__global__ void __launch_bounds__(64, MIN_CTAS) ed25519_batch_kernel(uint32_t n, EdIn in, const ge_niels_slot *__restrict__ table,
                                                           uint8_t *__restrict__ out, EdKeys keys) {
    ed25519_batch_body<INL, FP64>(n, in, table, out, keys);
}
// the same kernel under an explicit register cap (64-thread CTAs): still 2 warps per SM sub-partition, but more of the
// register file left to the SHA-256 warps that run beside it
template <int REGS, bool INL, bool FP64 = false>
__global__ void __maxnreg__(REGS) ed25519_batch_kernel_capped(uint32_t n, EdIn in, const ge_niels_slot *__restrict__ table,
                                                               uint8_t *__restrict__ out, EdKeys keys) {
    ed25519_batch_body<INL, FP64>(n, in, table, out, keys);
}

// inputs of signature i (the DUMMY triple for an inactive lane, eddsa.rs:28-30,62-63) and SHA-512(R || A || M)
__device__ __forceinline__ void ed_load_and_hash(const EdIn &in, uint32_t i, uint8_t pk[32], uint8_t sig[64], uint8_t digest[64]) {
    const bool on = !in.active || in.active[(size_t)in.active_stride * i];
    const uint8_t *m = in.msgs + (size_t)in.msg_stride * i;
    uint32_t len = in.msg_max;
    if (in.lens) {
        const uint8_t *lp = in.lens + (size_t)in.len_stride * i;
        len = (uint32_t)lp[0] | ((uint32_t)lp[1] << 8) | ((uint32_t)lp[2] << 16) | ((uint32_t)lp[3] << 24);
    }
    if (len > in.msg_max) len = in.msg_max;
    if (on) {
        for (int k = 0; k < 32; k++) pk[k] = in.pks[(size_t)in.pk_stride * i + k];
        for (int k = 0; k < 64; k++) sig[k] = in.sigs[(size_t)in.sig_stride * i + k];
    } else {
        for (int k = 0; k < 32; k++) pk[k] = DUMMY_PK[k];
        for (int k = 0; k < 64; k++) sig[k] = DUMMY_SIG[k];
        len = 32;  // DUMMY_MSG_LENGTH_BYTES: 32 zero bytes
    }
    uint64_t st[8];
    sha512_bytes(
        [&](uint32_t k) -> uint8_t { return k < 32 ? sig[k] : (k < 64 ? pk[k - 32] : (on ? m[k - 64] : (uint8_t)0)); },
        64 + len, st);
#pragma unroll
    for (int k = 0; k < 8; k++)
#pragma unroll
        for (int j = 0; j < 8; j++) digest[8 * k + j] = (uint8_t)(st[k] >> (56 - 8 * j));
}
__device__ __forceinline__ void ed_load_sig(const EdIn &in, uint32_t i, uint8_t sig[64]) {
    const bool on = !in.active || in.active[(size_t)in.active_stride * i];
    for (int k = 0; k < 64; k++) sig[k] = on ? in.sigs[(size_t)in.sig_stride * i + k] : DUMMY_SIG[k];
}

// the table path: thread t takes signatures t and t + ceil(n / 2) and inverts once for both (Montgomery's trick): with h*A
// down to 43 table additions the inversion is a quarter of a signature's field operations.  The grid is launched for one
// signature per thread (the verdict on the tables falls on the device), so the upper half of the threads leaves at once.
template <bool INL>
__device__ __forceinline__ void ed25519_keyed_pair(uint32_t n, const EdIn &in, const ge_niels_slot *__restrict__ table, uint8_t *__restrict__ out,
                                                   const EdKeys &keys) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, half = (n + 1) / 2;
    if (t >= half) return;
    edd::ed_pending pend[2];
    uint32_t flags[2];
    const bool two = t + half < n;
#pragma unroll 1
    for (int j = 0; j < (two ? 2 : 1); j++) {
        const uint32_t i = t + j * half;
        uint8_t pk[32], sig[64], digest[64];
        ed_load_and_hash(in, i, pk, sig, digest);
        const int32_t id = keys.slot_id[keys.key_slot[i]];
        pend[j] = edd::ed25519_witness_keyed_points<INL>(sig, digest, table, keys.recs + (size_t)BSX_ED_KEYREC_BYTES * id,
                                                         keys.tab + (size_t)id * BSX_ED_KEY_WINDOWS * BSX_ED_KEY_ENTRIES * 20,
                                                         out + (size_t)BSX_SIG_OUT_BYTES * i, flags[j]);
    }
    // Z is never zero (complete addition law; undecodable points were replaced by the identity): the product is invertible
    const edd::fed inv01 = edd::fed_invert(two ? edd::fed_mul(pend[0].z, pend[1].z) : pend[0].z);
#pragma unroll 1
    for (int j = 0; j < (two ? 2 : 1); j++) {
        const uint32_t i = t + j * half;
        uint8_t sig[64];
        ed_load_sig(in, i, sig);
        const edd::fed inv = two ? edd::fed_mul(inv01, pend[1 - j].z) : inv01;
        edd::ed25519_witness_finish(pend[j], inv, sig, flags[j], out + (size_t)BSX_SIG_OUT_BYTES * i);
    }
}

// The table path as a kernel of its own, so that its registers are allocated for its own code (no 8-entry local table, no
// doubling chain): does nothing unless the device-side verdict is "tables".
template <int MIN_CTAS, bool INL>
__global__ void __launch_bounds__(64, MIN_CTAS) ed25519_keyed_kernel(uint32_t n, EdIn in, const ge_niels_slot *__restrict__ table,
                                                                      uint8_t *__restrict__ out, EdKeys keys) {
    if (!ed_keys_in_use(keys, n)) return;
    if (keys.pair) {
        ed25519_keyed_pair<INL>(n, in, table, out, keys);
        return;
    }
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t pk[32], sig[64], digest[64];
    ed_load_and_hash(in, i, pk, sig, digest);
    const int32_t id = keys.slot_id[keys.key_slot[i]];
    edd::ed25519_witness_core_keyed<INL>(sig, digest, table, keys.recs + (size_t)BSX_ED_KEYREC_BYTES * id,
                                         keys.tab + (size_t)id * BSX_ED_KEY_WINDOWS * BSX_ED_KEY_ENTRIES * 20, out + (size_t)BSX_SIG_OUT_BYTES * i);
}

template <bool INL, bool FP64>
__device__ __forceinline__ void ed25519_batch_body(uint32_t n, const EdIn &in, const ge_niels_slot *__restrict__ table, uint8_t *__restrict__ out,
                                                   const EdKeys &keys) {
    // the table path has a kernel of its own (ed25519_keyed_kernel, launched just before this one): when the device-side
    // verdict is "tables", this kernel has nothing to do
    if (FP64 && ed_keys_in_use(keys, n)) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t pk[32], sig[64], digest[64];
    ed_load_and_hash(in, i, pk, sig, digest);
    if (FP64) {
        edd::ed25519_witness_core<INL>(pk, sig, digest, table, out + (size_t)BSX_SIG_OUT_BYTES * i);
    } else {
        ed25519_witness_core<INL>(pk, sig, digest, table, out + (size_t)BSX_SIG_OUT_BYTES * i);
    }
}


// ---------------------------------------------------------------------------------------------------------------
// three-stage path
// ---------------------------------------------------------------------------------------------------------------
// scratch per signature between the quad and finish kernels: X, Y, Z of sG, hA, R + hA (9 field elements)
#define BSX_ED_SCRATCH_WORDS 90
// default of the ED_FP64 tunable (-1): set by measurement, see DESIGN.md
// r02b (profiles/r02b_ab_fp64.txt): 37 888 signatures alone 1.81 -> 1.48 ms (compact), 1.67 -> 1.44 (inlined); header_range step
// 2.79 -> 2.64 ms.  Beside the SHA-256 kernels the uncapped FP64 build (240 registers) beat the 192-register cap (2.64 vs 2.68 ms).
#ifndef BSX_ED_FP64_DEFAULT
#define BSX_ED_FP64_DEFAULT true
#endif

// stage 1: thread 2i handles A of signature i (and the hashing), thread 2i+1 handles R
template <int DUMMY>
__global__ void __launch_bounds__(128) ed25519_prep_kernel(uint32_t n, EdIn in, uint8_t *__restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, i = t >> 1, which = t & 1;
    if (i >= n) return;
    uint8_t *rec = out + (size_t)BSX_SIG_OUT_BYTES * i;
    const bool on = !in.active || in.active[(size_t)in.active_stride * i];
    uint8_t pt[32];
    fe x, y;
    if (which == 0) {
        uint8_t pk[32], sig[64];
        const uint8_t *m = in.msgs + (size_t)in.msg_stride * i;
        uint32_t len = in.msg_max;
        if (in.lens) {
            const uint8_t *lp = in.lens + (size_t)in.len_stride * i;
            len = (uint32_t)lp[0] | ((uint32_t)lp[1] << 8) | ((uint32_t)lp[2] << 16) | ((uint32_t)lp[3] << 24);
        }
        if (len > in.msg_max) len = in.msg_max;
        if (on) {
            for (int k = 0; k < 32; k++) pk[k] = in.pks[(size_t)in.pk_stride * i + k];
            for (int k = 0; k < 64; k++) sig[k] = in.sigs[(size_t)in.sig_stride * i + k];
        } else {
            for (int k = 0; k < 32; k++) pk[k] = DUMMY_PK[k];
            for (int k = 0; k < 64; k++) sig[k] = DUMMY_SIG[k];
            len = 32;  // DUMMY_MSG_LENGTH_BYTES: 32 zero bytes (eddsa.rs:28-30,62-63)
        }
        uint64_t st[8];
        sha512_bytes(
            [&](uint32_t k) -> uint8_t { return k < 32 ? sig[k] : (k < 64 ? pk[k - 32] : (on ? m[k - 64] : (uint8_t)0)); },
            64 + len, st);
#pragma unroll
        for (int k = 0; k < 8; k++)
#pragma unroll
            for (int j = 0; j < 8; j++) rec[8 * k + j] = (uint8_t)(st[k] >> (56 - 8 * j));
        sc_divrem_l(rec, rec + 64, rec + 96);
        rec[521] = sc_lt_l(sig + 32) ? 1 : 0;
        for (int k = 0; k < 32; k++) pt[k] = pk[k];
        // the scalar s travels to the quad kernel in the (still unused) sG slot
        for (int k = 0; k < 32; k++) rec[136 + k] = sig[32 + k];
    } else {
        for (int k = 0; k < 32; k++) pt[k] = on ? in.sigs[(size_t)in.sig_stride * i + k] : DUMMY_SIG[k];
    }
    // A -> [200..296), flag byte 522;  R -> [360..456), flag byte 523
    uint8_t *dst = rec + (which ? 360 : 200);
    // the square-root chain (~265 dependent squarings: the latency of this kernel) on the FP64 pipe: 441 against 570 cycles
    // per squaring with one warp per sub-partition (profiles/r02a_ubench_fp64.txt); same bytes (tests: every build, host check)
    edd::fed fx, fy;
    const bool ok = edd::ged_decompress(pt, fx, fy, dst, dst + 32, dst + 64);
    rec[522 + which] = ok ? 1 : 0;
}

// ---- four lanes per signature: lane k of a quad holds coordinate k of the point (0 X, 1 Y, 2 Z, 3 T) ----
__device__ __forceinline__ fe fe_quad_get(const fe &a, int src) {
    fe r;
#pragma unroll
    for (int i = 0; i < 10; i++) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src, 4);
    return r;
}
__device__ __forceinline__ fe fe_quad_swap(const fe &a) {     // lanes 0<->1, 2<->3
    fe r;
#pragma unroll
    for (int i = 0; i < 10; i++) r.v[i] = __shfl_xor_sync(0xffffffffu, a.v[i], 1, 4);
    return r;
}
__device__ __forceinline__ fe fe_lin(int ca, const fe &a, int cb, const fe &b) {
    fe r;
#pragma unroll
    for (int i = 0; i < 10; i++) r.v[i] = ca * a.v[i] + cb * b.v[i];
    return r;
}
__device__ __forceinline__ fe fe_small(int32_t c) { fe r = fe_zero(); r.v[0] = c; return r; }

// completed point (one coordinate per lane) -> extended:  X3 = X T, Y3 = Z Y, Z3 = Z T, T3 = X Y.
// The right-hand factors (T or Y) go through fe_tighten, so a 4-unit T is fine (fe_mul bounds, ed25519.cuh).
__device__ __forceinline__ fe quad_p3(const fe &c, int k) {
    const fe t = fe_tighten(c);
    const fe f = fe_quad_get(c, (k == 1 || k == 2) ? 2 : 0);
    const fe g = fe_quad_get(t, (k & 1) ? 1 : 3);
    return fe_mul(f, g);
}
// doubling: lanes square X, Y, Z, X + Y;  X' = A - YY - XX, Y' = YY + XX, Z' = YY - XX, T' = 2 ZZ - YY + XX
__device__ __forceinline__ fe quad_dbl(const fe &c, int k) {
    const fe x = fe_quad_get(c, 0), y = fe_quad_get(c, 1);
    const fe s = fe_sq(k == 3 ? fe_add(x, y) : c);
    const fe xx = fe_quad_get(s, 0), yy = fe_quad_get(s, 1);
    const fe o = fe_quad_get(s, k == 0 ? 3 : 2);          // lane 0 takes A, lane 3 takes ZZ
    const int co = (k == 0) ? 1 : (k == 3 ? 2 : 0), cy = (k == 0 || k == 3) ? -1 : 1, cx = (k == 0 || k == 2) ? -1 : 1;
    fe r;
#pragma unroll
    for (int i = 0; i < 10; i++) r.v[i] = co * o.v[i] + cy * yy.v[i] + cx * xx.v[i];
    return r;
}
// addition of a precomputed point whose components sit one per lane: g = (Y+X, Y-X, 2Z, 2dT) of the addend
// (2Z = the constant 2 for an affine table entry).  Lanes multiply (Y+X) g0, (Y-X) g1, Z g2, T g3, then
// X' = a - b, Y' = a + b, Z' = zz2 + c, T' = zz2 - c with one exchange between neighbouring lanes.
__device__ __forceinline__ fe quad_add(const fe &c, const fe &g, int k) {
    const fe o = fe_quad_swap(c);
    const fe m = fe_mul(fe_lin(1, c, k == 0 ? 1 : (k == 1 ? -1 : 0), o), g);
    const fe o2 = fe_quad_swap(m);
    return fe_lin(k == 3 ? -1 : 1, m, k == 0 ? -1 : 1, o2);
}
// extended point -> its addend form, one component per lane
__device__ __forceinline__ fe quad_to_cached(const fe &c, int k) {
    const int32_t d2[10] = BSX_FE_2D;
    const fe o = fe_quad_swap(c);
    const fe v = fe_lin(k == 2 ? 2 : 1, c, k == 0 ? 1 : (k == 1 ? -1 : 0), o);
    return fe_mul(v, k == 3 ? fe_const(d2) : fe_one());
}

__global__ void __launch_bounds__(64) ed25519_quad_kernel(uint32_t n, const ge_niels_slot *__restrict__ table,
                                                           const uint8_t *__restrict__ out, int32_t *__restrict__ scratch) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = t & 3;
    uint32_t i = t >> 2;
    const bool live = i < n;
    if (!live) i = n - 1;                     // keep whole warps in the shuffles; no stores
    const uint8_t *rec = out + (size_t)BSX_SIG_OUT_BYTES * i;
    int32_t *scr = scratch + (size_t)BSX_ED_SCRATCH_WORDS * i;
    const fe ident = fe_small((k == 1 || k == 2) ? 1 : 0);            // (0, 1, 1, 0)
    const fe ident_addend = fe_small(k == 3 ? 0 : (k == 2 ? 2 : 1));  // (1, 1, 2, 0)
    const int32_t d2[10] = BSX_FE_2D;

    // ---- s*G: 32 windows of 8 bits over the affine table ----
    fe acc = ident;
#pragma unroll 1
    for (int w = 0; w < BSX_ED_BASE_WINDOWS; w++) {
        const uint32_t dgt = rec[136 + w];
        fe g = ident_addend;
        if (dgt && k != 2) {
            const int32_t *q = table[w * BSX_ED_BASE_ENTRIES + (dgt - 1)].v + (k == 3 ? 20 : 10 * k);
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const int2 v = __ldg(reinterpret_cast<const int2 *>(q) + j);
                g.v[2 * j] = v.x; g.v[2 * j + 1] = v.y;
            }
        }
        acc = quad_p3(quad_add(acc, g, k), k);
    }
    if (live && k < 3)
#pragma unroll
        for (int j = 0; j < 10; j++) scr[10 * k + j] = acc.v[j];

    // ---- table of A: tab[d] = addend form of d*A, d = 0..15 ----
    const fe ax = fe_frombytes(rec + 200), ay = fe_frombytes(rec + 232);
    const fe axy = fe_mul(ax, ay);
    fe cur = k == 0 ? ax : (k == 1 ? ay : (k == 2 ? fe_one() : axy));
    fe tab[16];
    tab[0] = ident_addend;
    tab[1] = quad_to_cached(cur, k);
#pragma unroll 1
    for (int d = 2; d < 16; d++) {
        cur = quad_p3(quad_add(cur, tab[1], k), k);
        tab[d] = quad_to_cached(cur, k);
    }
    // ---- h*A: 64 windows of 4 bits, most significant first ----
    acc = ident;
#pragma unroll 1
    for (int w = 63; w >= 0; w--) {
        if (w != 63) {
#pragma unroll 1
            for (int r = 0; r < 4; r++) acc = quad_p3(quad_dbl(acc, k), k);
        }
        const uint32_t dgt = (rec[64 + (w >> 1)] >> ((w & 1) * 4)) & 15;
        acc = quad_p3(quad_add(acc, tab[dgt], k), k);
    }
    if (live && k < 3)
#pragma unroll
        for (int j = 0; j < 10; j++) scr[30 + 10 * k + j] = acc.v[j];

    // ---- R + h*A ----
    const fe rx = fe_frombytes(rec + 360), ry = fe_frombytes(rec + 392);
    const fe rt = fe_mul(fe_mul(rx, ry), fe_const(d2));
    const fe gr = k == 0 ? fe_add(ry, rx) : (k == 1 ? fe_sub(ry, rx) : (k == 2 ? fe_small(2) : rt));
    acc = quad_p3(quad_add(acc, gr, k), k);
    if (live && k < 3)
#pragma unroll
        for (int j = 0; j < 10; j++) scr[60 + 10 * k + j] = acc.v[j];
}

// ---- the same kernel on the FP64 pipe (fe51d.cuh): the quad kernel is a chain of dependent field operations per
// signature (252 doublings x (one squaring + one multiplication)), and with one warp per sub-partition an FP64-limb
// squaring / multiplication takes 441 / 652 cycles against 570 / 772 for the integer limbs (profiles/r02a_ubench_fp64.txt).
// Units (fe51d.cuh: a product needs |f_i| |g_j| < 2^103, i.e. 2u x 2u or 3u x 1u of carried values): a doubling leaves
// X' = A - YY - XX 3u, Y' 2u, Z' 2u, T' = 2ZZ - YY + XX 4u; quad_p3d multiplies (X' or Z') by a re-carried (Y' or T').
using edd::fed;
__device__ __forceinline__ fed fed_quad_get(const fed &a, int src) {
    fed r;
#pragma unroll
    for (int i = 0; i < 5; i++) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src, 4);
    return r;
}
__device__ __forceinline__ fed fed_quad_swap(const fed &a) {     // lanes 0<->1, 2<->3
    fed r;
#pragma unroll
    for (int i = 0; i < 5; i++) r.v[i] = __shfl_xor_sync(0xffffffffu, a.v[i], 1, 4);
    return r;
}
__device__ __forceinline__ fed fed_lin(int ca, const fed &a, int cb, const fed &b) {   // small integer coefficients: exact
    fed r;
#pragma unroll
    for (int i = 0; i < 5; i++) r.v[i] = (double)ca * a.v[i] + (double)cb * b.v[i];
    return r;
}
__device__ __forceinline__ fed fed_small(int c) { fed r = edd::fed_zero(); r.v[0] = (double)c; return r; }
__device__ __forceinline__ fed quad_p3d(const fed &c, int k) {
    const fed t = edd::fed_reduce(c);
    const fed f = fed_quad_get(c, (k == 1 || k == 2) ? 2 : 0);
    const fed g = fed_quad_get(t, (k & 1) ? 1 : 3);
    return edd::fed_mul(f, g);
}
__device__ __forceinline__ fed quad_dbld(const fed &c, int k) {
    const fed x = fed_quad_get(c, 0), y = fed_quad_get(c, 1);
    const fed s = edd::fed_sq(k == 3 ? edd::fed_add(x, y) : c);
    const fed xx = fed_quad_get(s, 0), yy = fed_quad_get(s, 1);
    const fed o = fed_quad_get(s, k == 0 ? 3 : 2);          // lane 0 takes A, lane 3 takes ZZ
    const int co = (k == 0) ? 1 : (k == 3 ? 2 : 0), cy = (k == 0 || k == 3) ? -1 : 1, cx = (k == 0 || k == 2) ? -1 : 1;
    fed r;
#pragma unroll
    for (int i = 0; i < 5; i++) r.v[i] = (double)co * o.v[i] + (double)cy * yy.v[i] + (double)cx * xx.v[i];
    return r;
}
__device__ __forceinline__ fed quad_addd(const fed &c, const fed &g, int k) {
    const fed o = fed_quad_swap(c);
    const fed m = edd::fed_mul(fed_lin(1, c, k == 0 ? 1 : (k == 1 ? -1 : 0), o), g);
    const fed o2 = fed_quad_swap(m);
    return fed_lin(k == 3 ? -1 : 1, m, k == 0 ? -1 : 1, o2);
}
__device__ __forceinline__ fed quad_to_cachedd(const fed &c, int k) {
    const double d2[5] = BSX_FED_2D;
    const fed o = fed_quad_swap(c);
    const fed v = fed_lin(k == 2 ? 2 : 1, c, k == 0 ? 1 : (k == 1 ? -1 : 0), o);
    return edd::fed_mul(v, k == 3 ? edd::fed_const(d2) : edd::fed_one());
}
__device__ __forceinline__ void scr_store_fed(int32_t *scr, int p, const fed &a) {
    const fe f = edd::fe_from_fed(a);
#pragma unroll
    for (int j = 0; j < 10; j++) scr[10 * p + j] = f.v[j];
}

__global__ void __launch_bounds__(64) ed25519_quad_kernel_fp64(uint32_t n, const ge_niels_slot *__restrict__ table,
                                                                const uint8_t *__restrict__ out, int32_t *__restrict__ scratch) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = t & 3;
    uint32_t i = t >> 2;
    const bool live = i < n;
    if (!live) i = n - 1;                     // keep whole warps in the shuffles; no stores
    const uint8_t *rec = out + (size_t)BSX_SIG_OUT_BYTES * i;
    int32_t *scr = scratch + (size_t)BSX_ED_SCRATCH_WORDS * i;
    const fed ident = fed_small((k == 1 || k == 2) ? 1 : 0);            // (0, 1, 1, 0)
    const fed ident_addend = fed_small(k == 3 ? 0 : (k == 2 ? 2 : 1));  // (1, 1, 2, 0)
    const double d2[5] = BSX_FED_2D;
    // the two scalars once, into registers (the integer kernel re-reads a byte of the record from global memory per window)
    uint32_t sw[8], hw[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint8_t *ps = rec + 136 + 4 * j, *ph = rec + 64 + 4 * j;
        sw[j] = (uint32_t)ps[0] | ((uint32_t)ps[1] << 8) | ((uint32_t)ps[2] << 16) | ((uint32_t)ps[3] << 24);
        hw[j] = (uint32_t)ph[0] | ((uint32_t)ph[1] << 8) | ((uint32_t)ph[2] << 16) | ((uint32_t)ph[3] << 24);
    }
    auto word = [](const uint32_t (&a)[8], int j) {   // a[j] without a local-memory array
        uint32_t v = a[0];
#pragma unroll
        for (int q = 1; q < 8; q++) v = j == q ? a[q] : v;
        return v;
    };

    // ---- s*G: 32 windows of 8 bits over the affine table ----
    fed acc = ident;
#pragma unroll 1
    for (int w = 0; w < BSX_ED_BASE_WINDOWS; w++) {
        const uint32_t dgt = (word(sw, w >> 2) >> (8 * (w & 3))) & 255u;
        fed g = ident_addend;
        if (dgt && k != 2) {
            const int32_t *q = table[w * BSX_ED_BASE_ENTRIES + (dgt - 1)].v + (k == 3 ? 20 : 10 * k);
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const int2 v = __ldg(reinterpret_cast<const int2 *>(q) + j);
                g.v[j] = (double)v.x + 67108864.0 * (double)v.y;
            }
        }
        acc = quad_p3d(quad_addd(acc, g, k), k);
    }
    if (live && k < 3) scr_store_fed(scr, k, acc);

    // ---- table of A: tab[d] = addend form of d*A, d = 0..15 ----
    const fed ax = edd::fed_from_fe(fe_frombytes(rec + 200)), ay = edd::fed_from_fe(fe_frombytes(rec + 232));
    const fed axy = edd::fed_mul(ax, ay);
    fed cur = k == 0 ? ax : (k == 1 ? ay : (k == 2 ? edd::fed_one() : axy));
    fed tab[16];
    tab[0] = ident_addend;
    tab[1] = quad_to_cachedd(cur, k);
#pragma unroll 1
    for (int d = 2; d < 16; d++) {
        cur = quad_p3d(quad_addd(cur, tab[1], k), k);
        tab[d] = quad_to_cachedd(cur, k);
    }
    // ---- h*A: 64 windows of 4 bits, most significant first ----
    acc = ident;
#pragma unroll 1
    for (int w = 63; w >= 0; w--) {
        if (w != 63) {
#pragma unroll 1
            for (int r = 0; r < 4; r++) acc = quad_p3d(quad_dbld(acc, k), k);
        }
        const uint32_t dgt = (word(hw, w >> 3) >> (4 * (w & 7))) & 15u;
        acc = quad_p3d(quad_addd(acc, tab[dgt], k), k);
    }
    if (live && k < 3) scr_store_fed(scr, 3 + k, acc);

    // ---- R + h*A ----
    const fed rx = edd::fed_from_fe(fe_frombytes(rec + 360)), ry = edd::fed_from_fe(fe_frombytes(rec + 392));
    const fed rt = edd::fed_mul(edd::fed_mul(rx, ry), edd::fed_const(d2));
    const fed gr = k == 0 ? edd::fed_add(ry, rx) : (k == 1 ? edd::fed_sub(ry, rx) : (k == 2 ? fed_small(2) : rt));
    acc = quad_p3d(quad_addd(acc, gr, k), k);
    if (live && k < 3) scr_store_fed(scr, 6 + k, acc);
}

__device__ __forceinline__ fe scr_load(const int32_t *scr, int p) {
    fe r;
#pragma unroll
    for (int j = 0; j < 10; j++) r.v[j] = scr[10 * p + j];
    return r;
}

// stage 3: K signatures per thread.  The three projective results of a signature share one inversion, and the K
// signatures of a thread share it too (Montgomery's trick: prefix products, ONE field inversion, back-substitution):
// (265 + 3K) instead of 265 K field operations.  The quad-lane path uses K = 1 (more threads, lower latency).
// Z is never zero (complete addition law; undecodable inputs were replaced by the identity), so the product is invertible.
template <int K>
__global__ void __launch_bounds__(128) ed25519_finish_kernel(uint32_t n, const int32_t *__restrict__ scratch, uint8_t *__restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, i0 = t * K;
    if (i0 >= n) return;
    const uint32_t cnt = (n - i0 < (uint32_t)K) ? n - i0 : (uint32_t)K;
    fe z12[K], pre[K];                 // z12_j = Z_sg Z_ha;  pre_j = prod_{l <= j} z12_l Z_sum_l
#pragma unroll
    for (int j = 0; j < K; j++) {
        if ((uint32_t)j < cnt) {
            const int32_t *scr = scratch + (size_t)BSX_ED_SCRATCH_WORDS * (i0 + j);
            z12[j] = fe_mul(scr_load(scr, 2), scr_load(scr, 5));
            const fe z = fe_mul(z12[j], scr_load(scr, 8));
            pre[j] = j ? fe_mul(pre[j - 1], z) : z;
        }
    }
    // 1 / (z_0 ... z_{cnt-1}); the inversion chain on the FP64 pipe (as in the prep kernel)
    fe run = edd::fe_from_fed(edd::fed_invert(edd::fed_from_fe(pre[cnt - 1])));
#pragma unroll
    for (int j = K - 1; j >= 0; j--) {
        if ((uint32_t)j >= cnt) continue;
        const uint32_t i = i0 + j;
        uint8_t *rec = out + (size_t)BSX_SIG_OUT_BYTES * i;
        const int32_t *scr = scratch + (size_t)BSX_ED_SCRATCH_WORDS * i;
        const fe smZ = scr_load(scr, 8);
        fe inv = run;                                          // 1 / (z_0 ... z_j)
        if (j) {
            inv = fe_mul(run, pre[j - 1]);                     // 1 / z_j
            run = fe_mul(run, fe_mul(z12[j], smZ));            // 1 / (z_0 ... z_{j-1})
        }
        const fe sgZ = scr_load(scr, 2), haZ = scr_load(scr, 5);
        const fe isum = fe_mul(inv, z12[j]);
        const fe i12 = fe_mul(inv, smZ);
        const fe isg = fe_mul(i12, haZ), iha = fe_mul(i12, sgZ);
        fe_tobytes(rec + 136, fe_mul(scr_load(scr, 0), isg)); fe_tobytes(rec + 168, fe_mul(scr_load(scr, 1), isg));
        fe_tobytes(rec + 296, fe_mul(scr_load(scr, 3), iha)); fe_tobytes(rec + 328, fe_mul(scr_load(scr, 4), iha));
        fe_tobytes(rec + 456, fe_mul(scr_load(scr, 6), isum)); fe_tobytes(rec + 488, fe_mul(scr_load(scr, 7), isum));
        uint32_t flags = (rec[521] ? 1u : 0u) | (rec[522] ? 2u : 0u) | (rec[523] ? 4u : 0u);
        if (bytes_eq32(rec + 136, rec + 456) && bytes_eq32(rec + 168, rec + 488)) flags |= 8u;
        rec[520] = (uint8_t)flags; rec[521] = 0; rec[522] = 0; rec[523] = 0;
        for (int q = 524; q < 576; q++) rec[q] = 0;
    }
}

}  // namespace bsx

using namespace bsx;

// the s*G table lives in the ctx (built on first use, on the stream of that first call).  A later call on another stream
// must not read it before the build kernel has finished: every consumer stream waits on ev_table until the event has
// been seen complete once.
static int ensure_base_table(bsx_ctx *ctx, cudaStream_t st) {
    if (!ctx->ed_table) {
        const int entries = BSX_ED_BASE_WINDOWS * BSX_ED_BASE_ENTRIES;
        void *tab = nullptr;
        BSX_CUDA(ctx, cudaMalloc(&tab, sizeof(ed::ge_niels_slot) * entries));
        if (!ctx->ev_table && cudaEventCreateWithFlags(&ctx->ev_table, cudaEventDisableTiming) != cudaSuccess) {
            cudaFree(tab);
            return bsx::fail(ctx, BSX_ERR_CUDA, "cudaEventCreate (s*G table)%s%s");
        }
        ed25519_base_table_kernel<<<(entries + 63) / 64, 64, 0, st>>>(reinterpret_cast<ed::ge_niels_slot *>(tab));
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess || cudaEventRecord(ctx->ev_table, st) != cudaSuccess) {
            cudaStreamSynchronize(st);
            cudaFree(tab);
            return bsx::fail(ctx, BSX_ERR_CUDA, "s*G table build failed%s%s");
        }
        ctx->ed_table = tab;
        ctx->ed_table_pending = 1;
        return BSX_OK;
    }
    if (ctx->ed_table_pending) {
        if (cudaEventQuery(ctx->ev_table) == cudaSuccess) ctx->ed_table_pending = 0;
        else BSX_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_table, 0));
    }
    return BSX_OK;
}

// three-stage path; the 360-byte-per-signature scratch is stream-ordered (pool allocation, no ctx state)
static int launch_quad(bsx_ctx *ctx, cudaStream_t st, uint32_t n, const EdIn &in, const ed::ge_niels_slot *tab, uint8_t *out) {
    int32_t *scratch = nullptr;
    BSX_CUDA(ctx, cudaMallocAsync(&scratch, sizeof(int32_t) * BSX_ED_SCRATCH_WORDS * (size_t)n, st));
    BSX_PIN_CARVEOUT(ed25519_prep_kernel<0>); BSX_PIN_CARVEOUT(ed25519_quad_kernel); BSX_PIN_CARVEOUT(ed25519_finish_kernel<1>);
    cudaError_t e = cudaSuccess;
    ed25519_prep_kernel<0><<<(2 * n + 127) / 128, 128, 0, st>>>(n, in, out);
    ctx->launches++;
    if ((e = cudaGetLastError()) == cudaSuccess) {
        const int fp64_t = ctx->tun[BSX_TUN_ED_FP64];
        if (fp64_t < 0 ? BSX_ED_FP64_DEFAULT : fp64_t != 0) {
            BSX_PIN_CARVEOUT(ed25519_quad_kernel_fp64);
            ed25519_quad_kernel_fp64<<<(4 * n + 63) / 64, 64, 0, st>>>(n, tab, out, scratch);
        } else {
            ed25519_quad_kernel<<<(4 * n + 63) / 64, 64, 0, st>>>(n, tab, out, scratch);
        }
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) {
        ed25519_finish_kernel<1><<<(n + 127) / 128, 128, 0, st>>>(n, scratch, out);
        ctx->launches++;
        e = cudaGetLastError();
    }
    const cudaError_t ef = cudaFreeAsync(scratch, st);   // stream-ordered: also on the error paths
    if (e != cudaSuccess) return bsx::fail(ctx, BSX_ERR_CUDA, "kernel launch: %s%s", cudaGetErrorString(e));
    if (ef != cudaSuccess) return bsx::fail(ctx, BSX_ERR_CUDA, "cudaFreeAsync: %s%s", cudaGetErrorString(ef));
    return BSX_OK;
}

// one thread per signature.  The register file is per SM sub-partition (16 K registers): 2 warps of this kernel per
// sub-partition (4 CTAs/SM) -> 216 registers, 3 (6 CTAs) -> 168, 4 (8 CTAs) -> 128 with spills.
// `alone`: the batch is not issued next to SHA-256 kernels (bsx_ed25519_batch*), so the build with inlined point
// arithmetic is used; the verify_* / header_range paths share the SMs with the hash kernels and use the compact build.
// `corun` (pipelined host path): the 128-register build.  One wave of the 216-register build leaves room for a
// single 80-register warp per sub-partition, so the map kernels of the first chunks -- whose digests the D2H engine is
// waiting for -- queue behind it; the 128-register build leaves half the register file (single call 6.9 -> 6.35 ms).
static int launch_mono(bsx_ctx *ctx, cudaStream_t st, uint32_t n, const EdIn &in, const ed::ge_niels_slot *tab, uint8_t *out, bool alone,
                       bool corun) {
    const int env_occ = ctx->tun[BSX_TUN_ED_OCC];
    const int fp64_t = ctx->tun[BSX_TUN_ED_FP64];
    const bool fp64 = fp64_t < 0 ? BSX_ED_FP64_DEFAULT : fp64_t != 0;
    // FP64 limbs: the register budget follows the batch size (the widest build whose wave the batch fills, common.cuh);
    // ED_OCC = 4 / 6 / 8 forces one
    // Table path, two signatures per thread (one shared inversion): only when half the threads still give every SM five
    // CTAs -- r02n: 1135 ranges 4.76 -> 4.64 ms per step, but 757 ranges 3.26 -> 3.30 and 378 ranges 1.94 -> 2.33 (too few
    // warps to hide the latency).  The wave is then counted over n / 2 threads; a batch that falls back to the general path
    // on the device runs it in a build chosen for half its size.  ED_PAIR = 0 / 1 forces the choice.
    const int pair_t = ctx->tun[BSX_TUN_ED_PAIR];
    const bool pairs = fp64 && ctx->tun[BSX_TUN_ED_KEYTAB] != 0 &&
                       (pair_t >= 0 ? pair_t != 0 : (uint64_t)(n + 1) / 2 >= (uint64_t)ctx->sm_count * 5 * 64);
    const int wave_occ = fp64 ? bsx_ed_wave_occ(ctx, pairs ? (n + 1) / 2 : n) : 0;
    const int occ = env_occ ? env_occ : corun ? 8 : (wave_occ > 4 ? wave_occ : 4);
    const int inl = ctx->tun[BSX_TUN_ED_INLINE];   // -1: by call site
    const bool use_inl = inl < 0 ? alone : inl != 0;
    // (Splitting this path into prep / main / finish kernels with a 4-way batched inversion was measured slower:
    // 19.3 vs 21.4 M sig/s at 37 800 signatures, 2.98 vs 2.91 ms for the header_range step -- not kept.)
    // beside the SHA-256 kernels, for batches that fill whole waves, the build capped at 192 registers (no spills) is used:
    // each sub-partition keeps room for two 64-register hash warps instead of one (header_range step 2.81 -> 2.77 ms).
    // BSX_ED_REGS: 0 = never, non-zero = always (A/B).
    const int env_cap = ctx->tun[BSX_TUN_ED_REGS];
    const int cap = env_cap >= 0 ? env_cap : (bsx_ed_fills_waves_at(ctx, n, 4) && !fp64 ? 192 : 0);
    const unsigned grid = (n + 63) / 64;
    // ED_RESIDENT = k: unused dynamic shared memory as ballast so that at most k CTAs of this kernel are resident per SM
    // (80 KB of the pinned 114 KB carveout shared between them), which leaves registers and issue slots to the hash kernels
    const int resident = ctx->tun[BSX_TUN_ED_RESIDENT];
    const size_t ballast = resident > 1 ? (size_t)((80 * 1024 / resident) & ~1023) : 0;
    // per-key tables (FP64 build): ED_KEYTAB -1 = when keys repeat at least BSX_ED_KEY_MIN_USE times on average, 0 = never,
    // 1 = whenever the distinct keys fit.  The state is stream-ordered pool memory: nothing of it lives in the ctx.
    EdKeys keys = {};
    uint8_t *kmem = nullptr;
    const int keytab = ctx->tun[BSX_TUN_ED_KEYTAB];
    const size_t o_slot_id = sizeof(int32_t) * BSX_ED_KEY_SLOTS, o_state = 2 * o_slot_id, o_key_slot = o_state + 256,
                 o_recs = (o_key_slot + sizeof(int32_t) * (size_t)n + 255) & ~(size_t)255,
                 o_bases = o_recs + (size_t)BSX_ED_KEY_MAX * BSX_ED_KEYREC_BYTES,
                 o_tab = o_bases + sizeof(double) * 20 * BSX_ED_KEY_WINDOWS * BSX_ED_KEY_MAX,
                 k_total = o_tab + sizeof(double) * 20 * BSX_ED_KEY_ENTRIES * BSX_ED_KEY_WINDOWS * BSX_ED_KEY_MAX;
    if (fp64 && keytab != 0 && cudaMallocAsync((void **)&kmem, k_total, st) != cudaSuccess) {
        cudaGetLastError();   // no room for the tables: the batch simply takes the general path
        kmem = nullptr;
    }
    if (kmem) {
        keys.slots = reinterpret_cast<int32_t *>(kmem);
        keys.slot_id = reinterpret_cast<int32_t *>(kmem + o_slot_id);
        keys.state = reinterpret_cast<int32_t *>(kmem + o_state);
        keys.key_slot = reinterpret_cast<int32_t *>(kmem + o_key_slot);
        keys.recs = kmem + o_recs;
        keys.bases = reinterpret_cast<double *>(kmem + o_bases);
        keys.tab = reinterpret_cast<double *>(kmem + o_tab);
        keys.force = keytab > 0;
        keys.pair = pairs;
        cudaError_t e = cudaMemsetAsync(keys.slots, 0xff, o_slot_id, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(keys.state, 0, 256, st);
        if (e == cudaSuccess) {
            BSX_PIN_CARVEOUT(ed25519_key_assign_kernel); BSX_PIN_CARVEOUT(ed25519_key_bases_kernel); BSX_PIN_CARVEOUT(ed25519_key_table_kernel);
            ed25519_key_assign_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, in, keys);
            ed25519_key_bases_kernel<<<BSX_ED_KEY_SLOTS / 32, 32, 0, st>>>(n, in, keys);
            ed25519_key_table_kernel<<<BSX_ED_KEY_MAX, BSX_ED_KEY_WINDOWS, 0, st>>>(n, keys);
            ctx->launches += 3;
            e = cudaGetLastError();
            if (e == cudaSuccess) {   // the table-path kernel; the general kernel below then returns at once (and vice versa)
                const int kocc = ctx->tun[BSX_TUN_ED_KOCC] ? ctx->tun[BSX_TUN_ED_KOCC] : occ;
                if (kocc >= 8) {
                    BSX_PIN_CARVEOUT((ed25519_keyed_kernel<8, false>));
                    ed25519_keyed_kernel<8, false><<<grid, 64, 0, st>>>(n, in, tab, out, keys);
                } else if (kocc >= 6) {
                    BSX_PIN_CARVEOUT((ed25519_keyed_kernel<6, false>));
                    ed25519_keyed_kernel<6, false><<<grid, 64, 0, st>>>(n, in, tab, out, keys);
                } else if (use_inl) {
                    BSX_PIN_CARVEOUT((ed25519_keyed_kernel<4, true>));
                    ed25519_keyed_kernel<4, true><<<grid, 64, 0, st>>>(n, in, tab, out, keys);
                } else {
                    BSX_PIN_CARVEOUT((ed25519_keyed_kernel<4, false>));
                    ed25519_keyed_kernel<4, false><<<grid, 64, 0, st>>>(n, in, tab, out, keys);
                }
                ctx->launches++;
                e = cudaGetLastError();
            }
        }
        if (e != cudaSuccess) {
            cudaFreeAsync(kmem, st);
            return bsx::fail(ctx, BSX_ERR_CUDA, "Ed25519 key tables: %s%s", cudaGetErrorString(e));
        }
    }
#define BSX_ED_LAUNCH(K)                                      \
    do {                                                      \
        BSX_PIN_CARVEOUT((K));                                \
        K<<<grid, 64, ballast, st>>>(n, in, tab, out, keys);  \
    } while (0)
    if (cap && (!alone || env_cap > 0) && !env_occ && !corun && inl <= 0) {
        // (caps of 176 and 160 registers spill and were slower: profiles/r01p_step_ab.txt)
        // (the same cap with inlined point arithmetic: 378 ranges per step 2.768 -> 2.818 ms, 756 ranges 5.595 -> 5.530 ms -- not kept)
        if (fp64) BSX_ED_LAUNCH((ed25519_batch_kernel_capped<192, false, true>));
        else BSX_ED_LAUNCH((ed25519_batch_kernel_capped<192, false, false>));
    } else if (occ >= 8) {
        if (fp64) BSX_ED_LAUNCH((ed25519_batch_kernel<8, false, true>));
        else BSX_ED_LAUNCH((ed25519_batch_kernel<8, false, false>));
    } else if (occ >= 6) {
        if (fp64) BSX_ED_LAUNCH((ed25519_batch_kernel<6, false, true>));
        else BSX_ED_LAUNCH((ed25519_batch_kernel<6, false, false>));
    } else if (use_inl) {
        if (fp64) BSX_ED_LAUNCH((ed25519_batch_kernel<4, true, true>));
        else BSX_ED_LAUNCH((ed25519_batch_kernel<4, true, false>));
    } else {
        if (fp64) BSX_ED_LAUNCH((ed25519_batch_kernel<4, false, true>));
        else BSX_ED_LAUNCH((ed25519_batch_kernel<4, false, false>));
    }
#undef BSX_ED_LAUNCH
    const cudaError_t el = cudaGetLastError();
    if (kmem) cudaFreeAsync(kmem, st);   // stream-ordered: after the kernel above, also on the error path
    if (el != cudaSuccess) return bsx::fail(ctx, BSX_ERR_CUDA, "kernel launch: %s%s", cudaGetErrorString(el));
    ctx->launches++;
    return BSX_OK;
}

static int ed25519_strided(bsx_ctx *ctx, void *stream, uint32_t n, const uint8_t *pks, uint32_t pk_stride,
                           const uint8_t *sigs, uint32_t sig_stride, const uint8_t *msgs, uint32_t msg_stride,
                           uint32_t msg_max, const uint8_t *msg_lens, uint32_t len_stride, const uint8_t *active,
                           uint32_t active_stride, uint8_t *out, bool alone, bool corun = false) {
    BSX_REQUIRE(ctx, ctx && pks && sigs && (msgs || msg_max == 0) && out);
    BSX_REQUIRE(ctx, ((uintptr_t)out & 7) == 0);   // the records are stored as 8-byte words (cudaMalloc'd memory always is)
    if (n == 0) return BSX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_base_table(ctx, st);
    if (rc) return rc;
    EdIn in{pks, sigs, msgs, msg_lens, active, pk_stride, sig_stride, msg_stride, msg_max, len_stride, active_stride};
    const ed::ge_niels_slot *tab = reinterpret_cast<const ed::ge_niels_slot *>(ctx->ed_table);
    // BSX_ED_MODE: 1 = three-stage quad-lane path, 2 = one thread per signature, unset = by batch size
    const int forced = ctx->tun[BSX_TUN_ED_MODE];
    const uint32_t quad_max = (uint32_t)ctx->tun[BSX_TUN_ED_QUAD_MAX];
    // (A split of one large batch over both paths at once was measured: 25 600 signatures alone 1.84 -> 1.72 ms at a
    // 35 % quad share, but the header_range step next to the map kernels gets slower beyond 20 % -- not kept.)
    const bool quad = forced ? forced == 1 : n <= quad_max;
    return quad ? launch_quad(ctx, st, n, in, tab, out) : launch_mono(ctx, st, n, in, tab, out, alone, corun);
}

// strided form: used by the verify_* entry points to run straight over validator records, next to their SHA-256 kernels
extern "C" int bsx_ed25519_strided_dev(bsx_ctx *ctx, void *stream, uint32_t n, const uint8_t *pks, uint32_t pk_stride,
                                       const uint8_t *sigs, uint32_t sig_stride, const uint8_t *msgs, uint32_t msg_stride,
                                       uint32_t msg_max, const uint8_t *msg_lens, uint32_t len_stride, const uint8_t *active,
                                       uint32_t active_stride, uint8_t *out) {
    return ed25519_strided(ctx, stream, n, pks, pk_stride, sigs, sig_stride, msgs, msg_stride, msg_max, msg_lens, len_stride, active,
                           active_stride, out, false);
}

// internal: the strided form with the co-run register budget of the pipelined host path (k_header_range.cu)
int bsx_ed25519_strided_corun(bsx_ctx *ctx, void *stream, uint32_t n, const uint8_t *pks, uint32_t pk_stride, const uint8_t *sigs,
                              uint32_t sig_stride, const uint8_t *msgs, uint32_t msg_stride, uint32_t msg_max, const uint8_t *msg_lens,
                              uint32_t len_stride, const uint8_t *active, uint32_t active_stride, uint8_t *out, int corun) {
    return ed25519_strided(ctx, stream, n, pks, pk_stride, sigs, sig_stride, msgs, msg_stride, msg_max, msg_lens, len_stride, active,
                           active_stride, out, false, corun != 0);
}

extern "C" int bsx_ed25519_batch_dev(bsx_ctx *ctx, void *stream, uint32_t n, const uint8_t *pks, const uint8_t *sigs,
                                     const uint8_t *msgs, uint32_t msg_stride, const uint32_t *msg_lens,
                                     const uint8_t *active, uint8_t *out) {
    return ed25519_strided(ctx, stream, n, pks, 32, sigs, 64, msgs, msg_stride, msg_stride,
                           reinterpret_cast<const uint8_t *>(msg_lens), 4, active, 1, out, true);
}

extern "C" int bsx_ed25519_batch(bsx_ctx *ctx, uint32_t n, const uint8_t *pks, const uint8_t *sigs, const uint8_t *msgs,
                                 uint32_t msg_stride, const uint32_t *msg_lens, const uint8_t *active, uint8_t *out) {
    BSX_REQUIRE(ctx, ctx && pks && sigs && (msgs || msg_stride == 0) && out);
    if (n == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t N = n, s_msg = N * msg_stride;
    int rc = ws_begin(ctx, ws_size(32 * N) + ws_size(64 * N) + ws_size(s_msg + 16) + ws_size(4 * N) + ws_size(N) +
                               ws_size(BSX_SIG_OUT_BYTES * N));
    if (rc) return rc;
    uint8_t *d_pk = ws_take<uint8_t>(ctx, 32 * N), *d_sig = ws_take<uint8_t>(ctx, 64 * N);
    uint8_t *d_msg = ws_take<uint8_t>(ctx, s_msg + 16);
    uint32_t *d_len = ws_take<uint32_t>(ctx, N);
    uint8_t *d_act = ws_take<uint8_t>(ctx, N), *d_out = ws_take<uint8_t>(ctx, BSX_SIG_OUT_BYTES * N);
    cudaStream_t st = ctx->stream;
    BSX_CUDA(ctx, cudaMemcpyAsync(d_pk, pks, 32 * N, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_sig, sigs, 64 * N, cudaMemcpyHostToDevice, st));
    if (s_msg) BSX_CUDA(ctx, cudaMemcpyAsync(d_msg, msgs, s_msg, cudaMemcpyHostToDevice, st));
    if (msg_lens) BSX_CUDA(ctx, cudaMemcpyAsync(d_len, msg_lens, 4 * N, cudaMemcpyHostToDevice, st));
    if (active) BSX_CUDA(ctx, cudaMemcpyAsync(d_act, active, N, cudaMemcpyHostToDevice, st));
    rc = bsx_ed25519_batch_dev(ctx, st, n, d_pk, d_sig, d_msg, msg_stride, msg_lens ? d_len : nullptr,
                               active ? d_act : nullptr, d_out);
    if (rc) return rc;
    BSX_CUDA(ctx, cudaMemcpyAsync(out, d_out, BSX_SIG_OUT_BYTES * N, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaStreamSynchronize(st));
    return BSX_OK;
}

// prove_subchain<B> (the header_range map circuit) and the mapreduce reduce stage, fused per job.
//   BX/circuits/builder.rs:150-271 (map), :337-395 (reduce), :273-409 (prove_data_commitment)
// One CTA per map job: 2B threads verify the 2B Merkle inclusion proofs (leaf hash + 4 levels,
// both orderings per level), then the same CTA hashes the B data-root tuples and evaluates the
// fixed-shape Tendermint tree.  Inputs of the job are staged into shared memory with TMA bulk
// copies; every digest of the Curta request schedule is written out.
#include "common.cuh"
// addition groups of the SHA-256 rounds issued on the FMA pipe (sha256.cuh): map stage 2.02 -> 1.87 ms per 757 ranges (r02j)
#ifndef BSX_SHA_FMA_ADDS
#define BSX_SHA_FMA_ADDS 13
#endif
#include "sha256.cuh"
#include "tm_tree.cuh"

#include <stdlib.h>

namespace bsx {

struct SubchainArgs {
    const uint8_t *dh_leaf, *dh_aunts, *lb_leaf, *lb_aunts, *start_headers, *end_headers;
    // explicit per-job scalars (range_jobs == 0) ...
    const uint64_t *batch_start, *batch_end, *global_end;
    const uint8_t *global_end_header;
    // ... or derived from per-range public inputs (range_jobs = jobs per range)
    uint32_t range_jobs;
    const uint64_t *start_blocks, *end_blocks;
    const uint8_t *range_end_header;
    uint8_t *digests, *subchains;
    // peer placement of the subchain records (multi-GPU, bsx_prove_subchain_batch_p2p_dev): with n_peers > 0 the record
    // of local job j2 = (range r, local job jl) is stored straight into the gathered [ranges, total_jobs] array of the
    // rank that reduces range r -- peers[r / ranges_per_owner] + ((r % ranges_per_owner) * total_jobs + rank * per + jl) * 128
    // -- over NVLink peer memory; the exchange step of the map/reduce needs no collective kernel.
    uint8_t *peers[BSX_MAX_PEERS];
    uint32_t n_peers, p2p_rank, p2p_per, p2p_total_jobs, p2p_ranges_per_owner;
    // exchange completion folded into the kernel that writes the records (bsx_shard_step_dev): every CTA counts itself
    // on `sig_done` after its record stores; the last one publishes `sig_step` into sig_flags[w] of every peer with a
    // system-scope release store -- the reduce kernel of each rank acquires those flags, so no barrier kernel has to find
    // an SM slot between the resident Ed25519 CTAs.  sig_done == nullptr: no signalling.
    uint32_t *sig_done;
    uint32_t *sig_flags[BSX_MAX_PEERS];
    uint32_t sig_step;
};
#define BSX_SUBCHAIN_ARGS_NO_P2P {nullptr}, 0u, 0u, 0u, 0u, 0u, nullptr, {nullptr}, 0u

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// End of a record-writing kernel.  `wrote`: this thread stored records (peer memory) in this kernel.
__device__ __forceinline__ void subchain_signal(const SubchainArgs &a, bool wrote) {
    if (!a.sig_done) return;
    if (wrote) __threadfence_system();          // this thread's record stores are visible system-wide before the count
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(a.sig_done, 1u);
        if (prev == gridDim.x - 1) {             // last CTA of the launch: every record of this rank has been stored
            *a.sig_done = 0;                     // re-armed for the next launch (stream-ordered)
            __threadfence_system();
            for (uint32_t w = 0; w < a.n_peers; w++) st_release_sys(a.sig_flags[w], a.sig_step);
        }
    }
}

__device__ __forceinline__ uint32_t *subchain_record(const SubchainArgs &a, size_t j2) {
    if (a.n_peers == 0) return reinterpret_cast<uint32_t *>(a.subchains + j2 * BSX_SUBCHAIN_BYTES);
    const size_t r = j2 / a.p2p_per, jl = j2 % a.p2p_per;
    const size_t owner = r / a.p2p_ranges_per_owner, rl = r % a.p2p_ranges_per_owner;
    return reinterpret_cast<uint32_t *>(a.peers[owner] + (rl * a.p2p_total_jobs + (size_t)a.p2p_rank * a.p2p_per + jl) * BSX_SUBCHAIN_BYTES);
}

__device__ __forceinline__ void load_words_be(const uint8_t *p, uint32_t d[8]) {  // any alignment
#pragma unroll
    for (int k = 0; k < 8; k++)
        d[k] = ((uint32_t)p[4 * k] << 24) | ((uint32_t)p[4 * k + 1] << 16) | ((uint32_t)p[4 * k + 2] << 8) | p[4 * k + 3];
}

template <int B, bool TMA>
__global__ void __launch_bounds__((2 * B < 32) ? 32 : 2 * B) prove_subchain_kernel(SubchainArgs a) {
    constexpr int LEAF_DH = 34, LEAF_LB = 72;
    constexpr uint32_t SZ_DHL = (B * LEAF_DH + 15) & ~15u, SZ_LBL = (B * LEAF_LB + 15) & ~15u, SZ_AUNT = B * 128;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *s_dh_leaf = smem_raw;
    uint8_t *s_lb_leaf = s_dh_leaf + SZ_DHL;
    uint8_t *s_dh_aunts = s_lb_leaf + SZ_LBL;
    uint8_t *s_lb_aunts = s_dh_aunts + SZ_AUNT;
    uint32_t *s_dh_root = reinterpret_cast<uint32_t *>(s_lb_aunts + SZ_AUNT);
    uint32_t *s_lb_root = s_dh_root + 8 * B;
    uint32_t *s_A = s_lb_root + 8 * B;
    uint32_t *s_Bf = s_A + 8 * B;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_Bf + 4 * B + 4);
    uint32_t *s_fail = reinterpret_cast<uint32_t *>(s_bar + 1);

    const uint32_t tid = threadIdx.x;
    const size_t job = blockIdx.x;
    const uint8_t *g_dh_leaf = a.dh_leaf + job * (B * LEAF_DH), *g_lb_leaf = a.lb_leaf + job * (B * LEAF_LB);
    const uint8_t *g_dh_aunts = a.dh_aunts + job * SZ_AUNT, *g_lb_aunts = a.lb_aunts + job * SZ_AUNT;

    if (TMA) {
        if (tid == 0) {
            *s_fail = 0;
            mbar_init(s_bar, 1);
        }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(s_bar, B * LEAF_DH + B * LEAF_LB + 2 * SZ_AUNT);
            bulk_g2s(s_dh_leaf, g_dh_leaf, B * LEAF_DH, s_bar);
            bulk_g2s(s_lb_leaf, g_lb_leaf, B * LEAF_LB, s_bar);
            bulk_g2s(s_dh_aunts, g_dh_aunts, SZ_AUNT, s_bar);
            bulk_g2s(s_lb_aunts, g_lb_aunts, SZ_AUNT, s_bar);
        }
        mbar_wait(s_bar, 0);
    } else {
        if (tid == 0) *s_fail = 0;
        for (uint32_t k = tid; k < B * LEAF_DH; k += blockDim.x) s_dh_leaf[k] = g_dh_leaf[k];
        for (uint32_t k = tid; k < B * LEAF_LB; k += blockDim.x) s_lb_leaf[k] = g_lb_leaf[k];
        for (uint32_t k = tid; k < SZ_AUNT; k += blockDim.x) { s_dh_aunts[k] = g_dh_aunts[k]; s_lb_aunts[k] = g_lb_aunts[k]; }
        __syncthreads();
    }

    // per-job scalars
    uint64_t batch_start, batch_end, global_end;
    const uint8_t *g_end_hdr;
    if (a.range_jobs) {
        size_t r = job / a.range_jobs, j = job % a.range_jobs;
        batch_start = a.start_blocks[r] + (uint64_t)j * B;
        batch_end = batch_start + B;
        global_end = a.end_blocks[r];
        g_end_hdr = a.range_end_header + 32 * r;
    } else {
        batch_start = a.batch_start[job];
        batch_end = a.batch_end[job];
        global_end = a.global_end[job];
        g_end_hdr = a.global_end_header + 32 * job;
    }
    uint8_t *out = a.digests + job * (size_t)(20 * B - 1) * 32;

    // ---- phase 1: the 2B inclusion proofs (tendermint.rs:62-93) ----
    if (tid < 2 * B) {
        const uint32_t kind = tid / B, i = tid % B;  // 0: data_hash proof, 1: last_block_id proof
        const uint8_t *leaf = kind ? s_lb_leaf + LEAF_LB * i : s_dh_leaf + LEAF_DH * i;
        const uint8_t *aunts = (kind ? s_lb_aunts : s_dh_aunts) + 128 * i;
        const uint32_t bits = kind ? 0x4u : 0x6u;  // path of leaf 4 / leaf 6, LSB first (builder.rs:166-169)
        uint8_t *o = out + 32 * (size_t)(18 * i + 9 * kind);
        uint32_t h[8];
        tm_leaf_hash([&](uint32_t k) -> uint8_t { return leaf[k]; }, kind ? LEAF_LB : LEAF_DH, h);
        store_digest_be(o, h);
#pragma unroll 1
        for (int l = 0; l < 4; l++) {
            uint32_t au[8], left[8], right[8];
            const uint4 *ap = reinterpret_cast<const uint4 *>(aunts + 32 * l);
            uint4 a0 = ap[0], a1 = ap[1];
            au[0] = bswap32(a0.x); au[1] = bswap32(a0.y); au[2] = bswap32(a0.z); au[3] = bswap32(a0.w);
            au[4] = bswap32(a1.x); au[5] = bswap32(a1.y); au[6] = bswap32(a1.z); au[7] = bswap32(a1.w);
            tm_inner_hash(h, au, left);
            tm_inner_hash(au, h, right);
            store_digest_be(o + 32 + 64 * l, left);
            store_digest_be(o + 64 + 64 * l, right);
            bool sel = (bits >> l) & 1;
#pragma unroll
            for (int k = 0; k < 8; k++) h[k] = sel ? right[k] : left[k];
        }
        uint32_t *r = (kind ? s_lb_root : s_dh_root) + 8 * i;
#pragma unroll
        for (int k = 0; k < 8; k++) r[k] = h[k];
    }
    __syncthreads();

    // ---- phase 2: link checks (builder.rs:174-232) + tuple leaves (:133-139) ----
    const bool batch_enabled = batch_start < global_end;
    const uint64_t last = global_end - 1;           // last_block_to_process
    const uint64_t kk = last - batch_start;         // index of the last enabled iteration (if enabled)
    if (tid < B) {
        const uint32_t i = tid;
        uint32_t f = 0;
        const bool e_i = batch_enabled && (uint64_t)i <= kk;
        const bool is_last = (last == batch_start + i);
        uint32_t hh[8], cur[8];
        load_words_be(s_lb_leaf + LEAF_LB * i + 2, hh);  // header hash of block curr_idx inside last_block_id
        if (i == 0) load_words_be(a.start_headers + 32 * job, cur);
        else {
#pragma unroll
            for (int k = 0; k < 8; k++) cur[k] = s_lb_root[8 * (i - 1) + k];
        }
        if (e_i && !digest_eq(cur, hh)) f |= BSX_FAIL_PREV_HEADER;
        if (e_i && !digest_eq(s_dh_root + 8 * i, hh)) f |= BSX_FAIL_DATA_HASH;
        if (is_last) {
            uint32_t ge[8];
            load_words_be(g_end_hdr, ge);
            if (!digest_eq(s_lb_root + 8 * i, ge)) f |= BSX_FAIL_END_HEADER;
        }
        if (f) atomicOr(s_fail, f);
        // tuple leaf: 0x00 ‖ 0^24 ‖ u64be(batch_start+i) ‖ data_hash_i   (data_hash = dh_leaf[2..34])
        uint32_t x[8], d[8], w[16];
        load_words_be(s_dh_leaf + LEAF_DH * i + 2, x);
        const uint64_t hgt = batch_start + i;
#pragma unroll
        for (int k = 0; k < 6; k++) w[k] = 0;
        w[6] = (uint32_t)(hgt >> 40);
        w[7] = (uint32_t)(hgt >> 8);
        w[8] = ((uint32_t)hgt << 24) | (x[0] >> 8);
#pragma unroll
        for (int k = 1; k < 8; k++) w[8 + k] = __funnelshift_r(x[k], x[k - 1], 8);
        sha256_init(d);
        sha256_compress(d, w);
        sha256_tail65(d, x[7] & 0xffu);
        store_digest_be(out + 32 * (size_t)(18 * B + i), d);
#pragma unroll
        for (int k = 0; k < 8; k++) s_A[8 * i + k] = d[k];
    }
    __syncthreads();

    // ---- phase 3: data-commitment tree over the batch (builder.rs:234-252, tendermint.rs:165-204) ----
    const uint64_t temp_end = batch_end < global_end ? batch_end : global_end;
    const uint64_t end_block = temp_end < batch_start ? batch_start : temp_end;
    const uint64_t nb_blocks = end_block - batch_start;
    uint32_t root[8];
    tm_tree_cta(s_A, s_Bf, B, nb_blocks & 0xffffffffull, out + 32 * (size_t)(19 * B), root);

    if (tid == 0) {
        uint32_t f = *s_fail;
        if (nb_blocks >> 32) f |= BSX_FAIL_END_LT_START;
        // curr_header after the loop and the batch-end check (:228-232)
        uint32_t cur[8], eh[8];
        if (batch_enabled) {
            uint64_t idx = kk < (uint64_t)(B - 1) ? kk : (uint64_t)(B - 1);
#pragma unroll
            for (int k = 0; k < 8; k++) cur[k] = s_lb_root[8 * idx + k];
        } else {
            load_words_be(a.start_headers + 32 * job, cur);
        }
        const bool e_B = batch_enabled && (uint64_t)B <= kk;
        load_words_be(a.end_headers + 32 * job, eh);
        if (e_B && !digest_eq(cur, eh)) f |= BSX_FAIL_BATCH_END_HEADER;
        uint32_t *rec = subchain_record(a, job);
        rec[0] = batch_enabled ? 1u : 0u;
        rec[1] = f;
        rec[2] = (uint32_t)batch_start; rec[3] = (uint32_t)(batch_start >> 32);
        rec[4] = (uint32_t)end_block;   rec[5] = (uint32_t)(end_block >> 32);
        const uint32_t *sh = reinterpret_cast<const uint32_t *>(a.start_headers + 32 * job);
#pragma unroll
        for (int k = 0; k < 8; k++) rec[6 + k] = sh[k];
#pragma unroll
        for (int k = 0; k < 8; k++) rec[14 + k] = bswap32(cur[k]);
#pragma unroll
        for (int k = 0; k < 8; k++) rec[22 + k] = bswap32(root[k]);
        rec[30] = 0; rec[31] = 0;
    }
    subchain_signal(a, tid == 0);
}

// ---- reduce stage: one CTA per range, records are 32 little-endian words ----
// rec[0] is_enabled, [1] fail, [2,3] start_block, [4,5] end_block, [6..14) start_header,
// [14..22) end_header, [22..30) data_merkle_root (header/root words hold raw digest bytes).
__global__ void reduce_subchains_kernel(uint32_t n_jobs, uint32_t B, const uint8_t *__restrict__ map_subchains,
                                        const uint64_t *__restrict__ start_blocks, const uint8_t *__restrict__ start_header,
                                        const uint64_t *__restrict__ end_blocks, const uint8_t *__restrict__ end_header,
                                        uint8_t *__restrict__ reduce_digests, uint8_t *__restrict__ reduce_nodes,
                                        uint8_t *__restrict__ data_commitments, uint32_t *__restrict__ fail,
                                        const uint32_t *wait_flags, uint32_t n_wait, uint32_t wait_step) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *src = smem, *dst = smem + 32 * (size_t)n_jobs;
    const size_t r = blockIdx.x;
    __shared__ uint32_t s_timeout;
    if (wait_flags) {
        // the records of this range come from every rank's map kernel (peer stores): acquire each rank's step flag
        // (subchain_signal) before reading them.  Bounded: a rank that never arrives turns into BSX_FAIL_EXCHANGE_TIMEOUT.
        if (threadIdx.x == 0) s_timeout = 0;
        __syncthreads();
        if (threadIdx.x < n_wait) {
            uint64_t t0, t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            while ((int32_t)(ld_acquire_sys(wait_flags + threadIdx.x) - wait_step) < 0) {
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 4000000000ull) { s_timeout = 1; break; }      // 4 s
                __nanosleep(64);
            }
        }
        __syncthreads();
    }
    const uint32_t *g = reinterpret_cast<const uint32_t *>(map_subchains + r * n_jobs * BSX_SUBCHAIN_BYTES);
    if (wait_flags) {   // peer-written data: plain (non-cached-readonly) loads after the acquire
        const volatile uint32_t *gv = g;
        for (uint32_t k = threadIdx.x; k < 32 * n_jobs; k += blockDim.x) src[k] = gv[k];
    } else {
        for (uint32_t k = threadIdx.x; k < 32 * n_jobs; k += blockDim.x) src[k] = g[k];
    }
    __syncthreads();
    uint32_t off = 0;
    for (uint32_t len = n_jobs; len > 1; len /= 2) {
        for (uint32_t i = threadIdx.x; i < len / 2; i += blockDim.x) {
            const uint32_t *L = src + 64 * i, *R = L + 32;
            uint32_t f = L[1] | R[1];
            const bool right_disabled = (R[0] == 0);
            bool linked = (L[4] == R[2]) && (L[5] == R[3]);
#pragma unroll
            for (int k = 0; k < 8; k++) linked = linked && (L[14 + k] == R[6 + k]);
            if (!(right_disabled || linked)) f |= BSX_FAIL_REDUCE_LINK;
            uint32_t lr[8], rr[8], p[8];
#pragma unroll
            for (int k = 0; k < 8; k++) { lr[k] = bswap32(L[22 + k]); rr[k] = bswap32(R[22 + k]); }
            tm_inner_hash(lr, rr, p);  // sha256(0x01 ‖ left.root ‖ right.root), builder.rs:357-364
            if (reduce_digests) store_digest_be(reduce_digests + 32 * (r * (n_jobs - 1) + off + i), p);
            uint32_t o[32];
            o[0] = L[0]; o[1] = f; o[2] = L[2]; o[3] = L[3];
            o[4] = right_disabled ? L[4] : R[4];
            o[5] = right_disabled ? L[5] : R[5];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                o[6 + k] = L[6 + k];
                o[14 + k] = right_disabled ? L[14 + k] : R[14 + k];
                o[22 + k] = right_disabled ? L[22 + k] : bswap32(p[k]);
            }
            o[30] = 0; o[31] = 0;
            uint4 *d4 = reinterpret_cast<uint4 *>(dst + 32 * i);
            uint4 *g4 = reduce_nodes ? reinterpret_cast<uint4 *>(reduce_nodes + BSX_SUBCHAIN_BYTES * (r * (n_jobs - 1) + off + i)) : nullptr;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                uint4 v = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
                d4[k] = v;
                if (g4) g4[k] = v;
            }
        }
        off += len / 2;
        uint32_t *t = src; src = dst; dst = t;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        uint32_t f = src[1];
        const uint64_t sb = start_blocks[r], eb = end_blocks[r];
        if (!(eb <= sb + (uint64_t)n_jobs * B)) f |= BSX_FAIL_RANGE;  // builder.rs:291-297
        const uint32_t *sh = reinterpret_cast<const uint32_t *>(start_header + 32 * r);
        const uint32_t *eh = reinterpret_cast<const uint32_t *>(end_header + 32 * r);
        bool ok = src[2] == (uint32_t)sb && src[3] == (uint32_t)(sb >> 32) && src[4] == (uint32_t)eb && src[5] == (uint32_t)(eb >> 32);
#pragma unroll
        for (int k = 0; k < 8; k++) ok = ok && src[6 + k] == sh[k] && src[14 + k] == eh[k];
        if (!ok) f |= BSX_FAIL_RESULT;  // builder.rs:398-406
        if (wait_flags && s_timeout) f |= BSX_FAIL_EXCHANGE_TIMEOUT;
        uint32_t *dc = reinterpret_cast<uint32_t *>(data_commitments + 32 * r);
#pragma unroll
        for (int k = 0; k < 8; k++) dc[k] = src[22 + k];
        if (fail) fail[r] = f;
    }
}

// =============================================================================================
// Split form of the map stage (default): 90 % of a job's hashing is its 2B inclusion proofs, which need no
// cooperation at all -- so they get their own kernel, one proof per thread, no shared memory, no barriers,
// full warps.  A second, much smaller kernel packs the tuple leaves / trees / link checks of G jobs per CTA so
// that the shrinking tree levels still fill warps.  (The fused one-CTA-per-job kernel above spent 27 % of its
// issue slots waiting at barriers and ran its tree levels at 16, 8, 4, 2, 1 active lanes; profiles/r01b.)
// =============================================================================================
template <int MIN_CTAS>
__global__ void __launch_bounds__(128, MIN_CTAS) subchain_proofs_kernel(SubchainArgs a, uint32_t B, uint32_t n_jobs) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n_jobs * 2 * B) return;
    const size_t job = t / (2 * B);
    const uint32_t r = (uint32_t)(t % (2 * B)), kind = r / B, i = r % B;   // B = 32: a warp is all-data_hash or all-last_block_id
    const size_t hdr = job * B + i;
    const uint8_t *leaf = kind ? a.lb_leaf + hdr * 72 : a.dh_leaf + hdr * 34;
    const uint8_t *aunts = (kind ? a.lb_aunts : a.dh_aunts) + hdr * 128;
    const uint32_t bits = kind ? 0x4u : 0x6u;  // path of leaf 4 / leaf 6, LSB first (builder.rs:166-169)
    uint8_t *o = a.digests + (job * (size_t)(20 * B - 1) + 18 * i + 9 * kind) * 32;
    uint32_t h[8];
    tm_leaf_hash([&](uint32_t k) -> uint8_t { return __ldg(leaf + k); }, kind ? 72u : 34u, h);
    store_digest_be(o, h);
#pragma unroll 1
    for (int l = 0; l < 4; l++) {
        uint32_t au[8], left[8], right[8];
        load_words_be(aunts + 32 * l, au);
        tm_inner_hash(h, au, left);
        tm_inner_hash(au, h, right);
        store_digest_be(o + 32 + 64 * l, left);
        store_digest_be(o + 64 + 64 * l, right);
        const bool sel = (bits >> l) & 1;
#pragma unroll
        for (int k = 0; k < 8; k++) h[k] = sel ? right[k] : left[k];
    }
}

// G jobs per CTA, one thread per header: link checks (builder.rs:174-232), tuple leaves (:133-139), the G
// data-commitment trees evaluated as one forest (tendermint.rs:165-204), subchain records.
// Proof roots are read back from the digests the proofs kernel wrote (level-3 "left" digest: bit 3 of both paths is 0).
template <int B, int G>
__global__ void __launch_bounds__(B * G < 32 ? 32 : B * G) subchain_commit_kernel(SubchainArgs a, uint32_t n_jobs) {
    constexpr int T = B * G;
    // 112 bytes of shared memory per thread (dynamic: the wide variants exceed the 48 KB static limit)
    extern __shared__ __align__(16) uint32_t s_commit[];
    uint32_t *s_dh_root = s_commit, *s_lb_root = s_dh_root + 8 * T, *s_A = s_lb_root + 8 * T, *s_Bf = s_A + 8 * T;
    uint32_t *s_fail = s_Bf + 4 * T + 8;
    const uint32_t tid = threadIdx.x, g = tid / B, i = tid % B;
    const size_t job = (size_t)blockIdx.x * G + g;
    const bool live = tid < T && job < n_jobs;
    if (tid < G) s_fail[tid] = 0;

    uint64_t batch_start = 0, batch_end = 0, global_end = 0;
    const uint8_t *g_end_hdr = nullptr;
    uint8_t *out = nullptr;
    if (live) {
        if (a.range_jobs) {
            const size_t r = job / a.range_jobs, j = job % a.range_jobs;
            batch_start = a.start_blocks[r] + (uint64_t)j * B;
            batch_end = batch_start + B;
            global_end = a.end_blocks[r];
            g_end_hdr = a.range_end_header + 32 * r;
        } else {
            batch_start = a.batch_start[job];
            batch_end = a.batch_end[job];
            global_end = a.global_end[job];
            g_end_hdr = a.global_end_header + 32 * job;
        }
        out = a.digests + job * (size_t)(20 * B - 1) * 32;
        load_words_be(out + 32 * (size_t)(18 * i + 7), s_dh_root + 8 * tid);
        load_words_be(out + 32 * (size_t)(18 * i + 9 + 7), s_lb_root + 8 * tid);
    }
    __syncthreads();

    const bool batch_enabled = batch_start < global_end;
    const uint64_t last = global_end - 1;          // last_block_to_process
    const uint64_t kk = last - batch_start;        // index of the last enabled iteration (if enabled)
    if (live) {
        uint32_t f = 0;
        const bool e_i = batch_enabled && (uint64_t)i <= kk;
        const bool is_last = (last == batch_start + i);
        const uint8_t *lbl = a.lb_leaf + (job * B + i) * 72, *dhl = a.dh_leaf + (job * B + i) * 34;
        uint32_t hh[8], cur[8];
        load_words_be(lbl + 2, hh);                // header hash of block curr_idx inside last_block_id
        if (i == 0) load_words_be(a.start_headers + 32 * job, cur);
        else {
#pragma unroll
            for (int k = 0; k < 8; k++) cur[k] = s_lb_root[8 * (tid - 1) + k];
        }
        if (e_i && !digest_eq(cur, hh)) f |= BSX_FAIL_PREV_HEADER;
        if (e_i && !digest_eq(s_dh_root + 8 * tid, hh)) f |= BSX_FAIL_DATA_HASH;
        if (is_last) {
            uint32_t ge[8];
            load_words_be(g_end_hdr, ge);
            if (!digest_eq(s_lb_root + 8 * tid, ge)) f |= BSX_FAIL_END_HEADER;
        }
        if (f) atomicOr(&s_fail[g], f);
        // tuple leaf: 0x00 ‖ 0^24 ‖ u64be(batch_start+i) ‖ data_hash_i   (data_hash = dh_leaf[2..34])
        uint32_t x[8], d[8], w[16];
        load_words_be(dhl + 2, x);
        const uint64_t hgt = batch_start + i;
#pragma unroll
        for (int k = 0; k < 6; k++) w[k] = 0;
        w[6] = (uint32_t)(hgt >> 40);
        w[7] = (uint32_t)(hgt >> 8);
        w[8] = ((uint32_t)hgt << 24) | (x[0] >> 8);
#pragma unroll
        for (int k = 1; k < 8; k++) w[8 + k] = __funnelshift_r(x[k], x[k - 1], 8);
        sha256_init(d);
        sha256_compress(d, w);
        sha256_tail65(d, x[7] & 0xffu);
        store_digest_be(out + 32 * (size_t)(18 * B + i), d);
#pragma unroll
        for (int k = 0; k < 8; k++) s_A[8 * tid + k] = d[k];
    } else if (tid < T) {
#pragma unroll
        for (int k = 0; k < 8; k++) s_A[8 * tid + k] = 0;
    }
    __syncthreads();

    // the forest: level arrays of G*B, G*B/2, ... , G nodes; pairs never straddle trees because B is a power of two
    uint32_t *src = s_A, *dst = s_Bf;
    uint32_t off = 0, shift = 0;
#pragma unroll 1
    for (uint32_t len = B; len > 1; len >>= 1) {          // len = nodes per tree on the source level
        const uint32_t half = len >> 1;
        if (tid < G * half) {
            const uint32_t tg = tid / half, j = tid % half;  // tree, node within the tree's next level
            const size_t tjob = (size_t)blockIdx.x * G + tg;
            uint32_t l[8], r[8], p[8];
            const uint4 *sp = reinterpret_cast<const uint4 *>(src + 16 * tid);
            const uint4 v0 = sp[0], v1 = sp[1], v2 = sp[2], v3 = sp[3];
            l[0] = v0.x; l[1] = v0.y; l[2] = v0.z; l[3] = v0.w; l[4] = v1.x; l[5] = v1.y; l[6] = v1.z; l[7] = v1.w;
            r[0] = v2.x; r[1] = v2.y; r[2] = v2.z; r[3] = v2.w; r[4] = v3.x; r[5] = v3.y; r[6] = v3.z; r[7] = v3.w;
            tm_inner_hash(l, r, p);
            uint64_t nb = 0;
            if (tjob < n_jobs) {
                uint64_t bs, be, ge;
                if (a.range_jobs) {
                    const size_t rr = tjob / a.range_jobs, jj = tjob % a.range_jobs;
                    bs = a.start_blocks[rr] + (uint64_t)jj * B; be = bs + B; ge = a.end_blocks[rr];
                } else { bs = a.batch_start[tjob]; be = a.batch_end[tjob]; ge = a.global_end[tjob]; }
                const uint64_t te = be < ge ? be : ge, eb = te < bs ? bs : te;
                nb = (eb - bs) & 0xffffffffull;
                store_digest_be(a.digests + (tjob * (size_t)(20 * B - 1) + 19 * B + off + j) * 32, p);
            }
            const bool both = ((uint64_t)(2 * j + 1) << shift) < nb;
            uint4 *o4 = reinterpret_cast<uint4 *>(dst + 8 * tid);
            o4[0] = both ? make_uint4(p[0], p[1], p[2], p[3]) : v0;
            o4[1] = both ? make_uint4(p[4], p[5], p[6], p[7]) : v1;
        }
        off += half;
        shift++;
        uint32_t *tmp = src; src = dst; dst = tmp;
        __syncthreads();
    }

    if (tid < G && (size_t)blockIdx.x * G + tid < n_jobs) {
        const size_t j2 = (size_t)blockIdx.x * G + tid;
        uint64_t bs, be, ge;
        if (a.range_jobs) {
            const size_t rr = j2 / a.range_jobs, jj = j2 % a.range_jobs;
            bs = a.start_blocks[rr] + (uint64_t)jj * B; be = bs + B; ge = a.end_blocks[rr];
        } else { bs = a.batch_start[j2]; be = a.batch_end[j2]; ge = a.global_end[j2]; }
        const bool en = bs < ge;
        const uint64_t k2 = ge - 1 - bs;
        const uint64_t te = be < ge ? be : ge, end_block = te < bs ? bs : te, nb_blocks = end_block - bs;
        uint32_t f = s_fail[tid];
        if (nb_blocks >> 32) f |= BSX_FAIL_END_LT_START;
        uint32_t cur[8], eh[8];
        if (en) {
            const uint64_t idx = k2 < (uint64_t)(B - 1) ? k2 : (uint64_t)(B - 1);
#pragma unroll
            for (int k = 0; k < 8; k++) cur[k] = s_lb_root[8 * (tid * B + idx) + k];
        } else {
            load_words_be(a.start_headers + 32 * j2, cur);
        }
        const bool e_B = en && (uint64_t)B <= k2;
        load_words_be(a.end_headers + 32 * j2, eh);
        if (e_B && !digest_eq(cur, eh)) f |= BSX_FAIL_BATCH_END_HEADER;
        uint32_t *rec = subchain_record(a, j2);
        rec[0] = en ? 1u : 0u;
        rec[1] = f;
        rec[2] = (uint32_t)bs; rec[3] = (uint32_t)(bs >> 32);
        rec[4] = (uint32_t)end_block; rec[5] = (uint32_t)(end_block >> 32);
        const uint32_t *sh = reinterpret_cast<const uint32_t *>(a.start_headers + 32 * j2);
#pragma unroll
        for (int k = 0; k < 8; k++) rec[6 + k] = sh[k];
#pragma unroll
        for (int k = 0; k < 8; k++) rec[14 + k] = bswap32(cur[k]);
#pragma unroll
        for (int k = 0; k < 8; k++) rec[22 + k] = bswap32(src[8 * tid + k]);   // tree g's root sits at src[g]
        rec[30] = 0; rec[31] = 0;
    }
    subchain_signal(a, tid < G);
}

template <int B, int G>
static int launch_commit(bsx_ctx *ctx, cudaStream_t st, uint32_t n_jobs, const SubchainArgs &a) {
    constexpr int T = B * G < 32 ? 32 : B * G;
    constexpr size_t smem = sizeof(uint32_t) * (28 * (B * G) + 8 + G);
    auto k = subchain_commit_kernel<B, G>;
    if (smem > 48 * 1024) {
        static const cudaError_t once = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (once != cudaSuccess) return bsx::fail(ctx, BSX_ERR_CUDA, "cudaFuncSetAttribute(commit kernel): %s%s", cudaGetErrorString(once));
    } else {
        BSX_PIN_CARVEOUT(k);
    }
    k<<<(n_jobs + G - 1) / G, T, smem, st>>>(a, n_jobs);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

template <int B>
static int launch_subchain_split(bsx_ctx *ctx, cudaStream_t st, uint32_t n_jobs, const SubchainArgs &a) {
    const size_t total = (size_t)n_jobs * 2 * B;
    const int occ = ctx->tun[BSX_TUN_PROOFS_OCC];
    BSX_PIN_CARVEOUT(subchain_proofs_kernel<8>); BSX_PIN_CARVEOUT(subchain_proofs_kernel<6>);
    if (occ >= 8) subchain_proofs_kernel<8><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(a, B, n_jobs);   // 64 registers
    else subchain_proofs_kernel<6><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(a, B, n_jobs);             // 80 registers
    BSX_LAUNCHED(ctx);
    // COMMIT_THREADS (tunable): threads per CTA of the commit kernel = jobs per CTA x B.  The tree levels run at T/2, T/4, ...
    // active threads, so a wider CTA keeps more of them whole warps (B = 32: 128 threads -> levels of 64 .. 4 threads,
    // three of five below a warp; 1024 threads -> 512 .. 32, all whole warps) at the price of fewer, fatter barriers.
    const int want = ctx->tun[BSX_TUN_COMMIT_THREADS];
    if ((B == 32 || B == 64) && want >= 256) {
        if (want >= 1024) return launch_commit<B, 1024 / B>(ctx, st, n_jobs, a);
        if (want >= 512) return launch_commit<B, 512 / B>(ctx, st, n_jobs, a);
        return launch_commit<B, 256 / B>(ctx, st, n_jobs, a);
    }
    return launch_commit<B, (B >= 128 ? 1 : 128 / B)>(ctx, st, n_jobs, a);
}

template <int B>
static int launch_subchain(bsx_ctx *ctx, cudaStream_t st, uint32_t n_jobs, const SubchainArgs &a, bool aligned) {
    constexpr uint32_t SZ_DHL = (B * 34 + 15) & ~15u, SZ_LBL = (B * 72 + 15) & ~15u, SZ_AUNT = B * 128;
    const size_t smem = SZ_DHL + SZ_LBL + 2 * SZ_AUNT + 4 * (8 * B + 8 * B + 8 * B + 4 * B + 4) + 16;
    const int threads = (2 * B < 32) ? 32 : 2 * B;
    constexpr bool CAN_TMA = (B * 34) % 16 == 0;
    if (CAN_TMA && aligned) {
        auto k = prove_subchain_kernel<B, CAN_TMA>;
        if (smem > 48 * 1024) BSX_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<n_jobs, threads, smem, st>>>(a);
    } else {
        auto k = prove_subchain_kernel<B, false>;
        if (smem > 48 * 1024) BSX_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<n_jobs, threads, smem, st>>>(a);
    }
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

static bool use_fused_map(const bsx_ctx *ctx) {  // SUBCHAIN_FUSED=1 selects the one-CTA-per-job kernel (A/B measurements)
    return ctx->tun[BSX_TUN_SUBCHAIN_FUSED] == 1;
}

static int dispatch_subchain(bsx_ctx *ctx, cudaStream_t st, uint32_t B, uint32_t n_jobs, const SubchainArgs &a) {
    if (!use_fused_map(ctx)) {
        switch (B) {
            case 1: return launch_subchain_split<1>(ctx, st, n_jobs, a);
            case 2: return launch_subchain_split<2>(ctx, st, n_jobs, a);
            case 4: return launch_subchain_split<4>(ctx, st, n_jobs, a);
            case 8: return launch_subchain_split<8>(ctx, st, n_jobs, a);
            case 16: return launch_subchain_split<16>(ctx, st, n_jobs, a);
            case 32: return launch_subchain_split<32>(ctx, st, n_jobs, a);
            case 64: return launch_subchain_split<64>(ctx, st, n_jobs, a);
            case 128: return launch_subchain_split<128>(ctx, st, n_jobs, a);
            case 256: return launch_subchain_split<256>(ctx, st, n_jobs, a);
            default: return fail(ctx, BSX_ERR_INVALID, "BATCH_SIZE must be a power of two in [1,256]%s%s");
        }
    }
    const uintptr_t al = reinterpret_cast<uintptr_t>(a.dh_leaf) | reinterpret_cast<uintptr_t>(a.dh_aunts) |
                         reinterpret_cast<uintptr_t>(a.lb_leaf) | reinterpret_cast<uintptr_t>(a.lb_aunts);
    const bool aligned = (al & 15) == 0;
    switch (B) {
        case 1: return launch_subchain<1>(ctx, st, n_jobs, a, aligned);
        case 2: return launch_subchain<2>(ctx, st, n_jobs, a, aligned);
        case 4: return launch_subchain<4>(ctx, st, n_jobs, a, aligned);
        case 8: return launch_subchain<8>(ctx, st, n_jobs, a, aligned);
        case 16: return launch_subchain<16>(ctx, st, n_jobs, a, aligned);
        case 32: return launch_subchain<32>(ctx, st, n_jobs, a, aligned);
        case 64: return launch_subchain<64>(ctx, st, n_jobs, a, aligned);
        case 128: return launch_subchain<128>(ctx, st, n_jobs, a, aligned);
        case 256: return launch_subchain<256>(ctx, st, n_jobs, a, aligned);
        default: return fail(ctx, BSX_ERR_INVALID, "BATCH_SIZE must be a power of two in [1,256]%s%s");
    }
}

}  // namespace bsx

using namespace bsx;

extern "C" int bsx_prove_subchain_batch_dev(bsx_ctx *ctx, void *stream, uint32_t B, uint32_t n_jobs,
                                            const uint8_t *dh_leaf, const uint8_t *dh_aunts, const uint8_t *lb_leaf,
                                            const uint8_t *lb_aunts, const uint8_t *start_headers,
                                            const uint8_t *end_headers, const uint64_t *batch_start,
                                            const uint64_t *batch_end, const uint64_t *global_end,
                                            const uint8_t *global_end_header, uint8_t *digests, uint8_t *subchains) {
    BSX_REQUIRE(ctx, ctx && dh_leaf && dh_aunts && lb_leaf && lb_aunts && start_headers && end_headers && batch_start &&
                         batch_end && global_end && global_end_header && digests && subchains);
    BSX_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(digests) | reinterpret_cast<uintptr_t>(subchains)) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(start_headers) & 3) == 0);
    if (n_jobs == 0) return BSX_OK;
    SubchainArgs a{dh_leaf, dh_aunts, lb_leaf, lb_aunts, start_headers, end_headers, batch_start, batch_end, global_end,
                   global_end_header, 0, nullptr, nullptr, nullptr, digests, subchains, BSX_SUBCHAIN_ARGS_NO_P2P};
    return dispatch_subchain(ctx, (cudaStream_t)stream, B, n_jobs, a);
}

// multi-GPU form: the same map jobs, but every subchain record is stored into the reducing rank's gathered array through
// peer memory (pointers from torch symmetric memory / CUDA IPC; NVLink).  peer_bases[w] = base of rank w's
// [ranges_per_owner, total_jobs, 128] array as mapped in THIS process.  The caller runs a cross-rank barrier before the reduce.
extern "C" int bsx_prove_subchain_batch_p2p_dev(bsx_ctx *ctx, void *stream, uint32_t B, uint32_t n_jobs,
                                                const uint8_t *dh_leaf, const uint8_t *dh_aunts, const uint8_t *lb_leaf,
                                                const uint8_t *lb_aunts, const uint8_t *start_headers,
                                                const uint8_t *end_headers, const uint64_t *batch_start,
                                                const uint64_t *batch_end, const uint64_t *global_end,
                                                const uint8_t *global_end_header, uint8_t *digests,
                                                const uint64_t *peer_bases, uint32_t n_peers, uint32_t rank,
                                                uint32_t jobs_per_rank, uint32_t total_jobs, uint32_t ranges_per_owner) {
    BSX_REQUIRE(ctx, ctx && dh_leaf && dh_aunts && lb_leaf && lb_aunts && start_headers && end_headers && batch_start &&
                         batch_end && global_end && global_end_header && digests && peer_bases);
    BSX_REQUIRE(ctx, n_peers >= 1 && n_peers <= BSX_MAX_PEERS && rank < n_peers && jobs_per_rank >= 1 && ranges_per_owner >= 1 &&
                         total_jobs == jobs_per_rank * n_peers && n_jobs % jobs_per_rank == 0 &&
                         (n_jobs / jobs_per_rank) == ranges_per_owner * n_peers);
    BSX_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(digests) & 15) == 0 && (reinterpret_cast<uintptr_t>(start_headers) & 3) == 0);
    if (n_jobs == 0) return BSX_OK;
    SubchainArgs a{dh_leaf, dh_aunts, lb_leaf, lb_aunts, start_headers, end_headers, batch_start, batch_end, global_end,
                   global_end_header, 0, nullptr, nullptr, nullptr, digests, nullptr, {nullptr}, n_peers, rank, jobs_per_rank,
                   total_jobs, ranges_per_owner, nullptr, {nullptr}, 0u};
    for (uint32_t w = 0; w < n_peers; w++) {
        BSX_REQUIRE(ctx, peer_bases[w] != 0 && (peer_bases[w] & 15) == 0);
        a.peers[w] = reinterpret_cast<uint8_t *>(peer_bases[w]);
    }
    return dispatch_subchain(ctx, (cudaStream_t)stream, B, n_jobs, a);
}

int bsx_reduce_subchains_wait_dev(bsx_ctx *ctx, void *stream, uint32_t n_ranges, uint32_t n_jobs, const uint8_t *map_subchains,
                                  const uint64_t *start_blocks, const uint8_t *start_header, const uint64_t *end_blocks,
                                  const uint8_t *end_header, uint32_t B, uint8_t *reduce_digests, uint8_t *reduce_nodes,
                                  uint8_t *data_commitments, uint32_t *fail, const uint32_t *wait_flags, uint32_t n_wait,
                                  uint32_t wait_step) {
    BSX_REQUIRE(ctx, ctx && map_subchains && start_blocks && start_header && end_blocks && end_header && data_commitments);
    BSX_REQUIRE(ctx, n_jobs >= 1 && n_jobs <= 1024 && (n_jobs & (n_jobs - 1)) == 0);
    BSX_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(map_subchains) | reinterpret_cast<uintptr_t>(reduce_digests) |
                       reinterpret_cast<uintptr_t>(reduce_nodes)) & 15) == 0 &&
                         ((reinterpret_cast<uintptr_t>(start_header) | reinterpret_cast<uintptr_t>(end_header) |
                           reinterpret_cast<uintptr_t>(data_commitments)) & 3) == 0);
    BSX_REQUIRE(ctx, !wait_flags || (n_wait >= 1 && n_wait <= BSX_MAX_PEERS && fail));
    if (n_ranges == 0) return BSX_OK;
    uint32_t threads = n_jobs / 2 < 32 ? 32 : n_jobs / 2;
    size_t smem = 4 * (32 * (size_t)n_jobs + 16 * (size_t)n_jobs);
    if (smem > 48 * 1024)
        BSX_CUDA(ctx, cudaFuncSetAttribute(reduce_subchains_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (smem <= 48 * 1024) BSX_PIN_CARVEOUT(reduce_subchains_kernel);
    reduce_subchains_kernel<<<n_ranges, threads, smem, (cudaStream_t)stream>>>(
        n_jobs, B, map_subchains, start_blocks, start_header, end_blocks, end_header, reduce_digests, reduce_nodes,
        data_commitments, fail, wait_flags, n_wait, wait_step);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

extern "C" int bsx_reduce_subchains_dev(bsx_ctx *ctx, void *stream, uint32_t n_ranges, uint32_t n_jobs,
                                        const uint8_t *map_subchains, const uint64_t *start_blocks,
                                        const uint8_t *start_header, const uint64_t *end_blocks,
                                        const uint8_t *end_header, uint32_t B, uint8_t *reduce_digests,
                                        uint8_t *reduce_nodes, uint8_t *data_commitments, uint32_t *fail) {
    return bsx_reduce_subchains_wait_dev(ctx, stream, n_ranges, n_jobs, map_subchains, start_blocks, start_header, end_blocks,
                                         end_header, B, reduce_digests, reduce_nodes, data_commitments, fail, nullptr, 0, 0);
}

// map stage of the sharded engine (k_shard.cu): peer stores + completion flags
int bsx_subchain_map_signal_dev(bsx_ctx *ctx, void *stream, uint32_t B, uint32_t n_jobs, const uint8_t *dh_leaf,
                                const uint8_t *dh_aunts, const uint8_t *lb_leaf, const uint8_t *lb_aunts,
                                const uint8_t *start_headers, const uint8_t *end_headers, const uint64_t *batch_start,
                                const uint64_t *batch_end, const uint64_t *global_end, const uint8_t *global_end_header,
                                uint8_t *digests, uint8_t *const *peer_bases, uint32_t n_peers, uint32_t rank,
                                uint32_t jobs_per_rank, uint32_t total_jobs, uint32_t ranges_per_owner, uint32_t *sig_done,
                                uint32_t *const *sig_flags, uint32_t sig_step) {
    BSX_REQUIRE(ctx, ctx && dh_leaf && dh_aunts && lb_leaf && lb_aunts && start_headers && end_headers && batch_start &&
                         batch_end && global_end && global_end_header && digests && peer_bases && sig_done && sig_flags);
    BSX_REQUIRE(ctx, n_peers >= 1 && n_peers <= BSX_MAX_PEERS && rank < n_peers && jobs_per_rank >= 1 && ranges_per_owner >= 1 &&
                         total_jobs == jobs_per_rank * n_peers && n_jobs % jobs_per_rank == 0 &&
                         (n_jobs / jobs_per_rank) == ranges_per_owner * n_peers);
    BSX_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(digests) & 15) == 0 && (reinterpret_cast<uintptr_t>(start_headers) & 3) == 0);
    if (n_jobs == 0) return BSX_OK;
    SubchainArgs a{dh_leaf, dh_aunts, lb_leaf, lb_aunts, start_headers, end_headers, batch_start, batch_end, global_end,
                   global_end_header, 0, nullptr, nullptr, nullptr, digests, nullptr, {nullptr}, n_peers, rank, jobs_per_rank,
                   total_jobs, ranges_per_owner, sig_done, {nullptr}, sig_step};
    for (uint32_t w = 0; w < n_peers; w++) {
        BSX_REQUIRE(ctx, peer_bases[w] && (reinterpret_cast<uintptr_t>(peer_bases[w]) & 15) == 0 && sig_flags[w]);
        a.peers[w] = peer_bases[w];
        a.sig_flags[w] = sig_flags[w];
    }
    return dispatch_subchain(ctx, (cudaStream_t)stream, B, n_jobs, a);
}

extern "C" int bsx_prove_data_commitment_dev(bsx_ctx *ctx, void *stream, uint32_t n_ranges, uint32_t n_jobs, uint32_t B,
                                             const uint8_t *dh_leaf, const uint8_t *dh_aunts, const uint8_t *lb_leaf,
                                             const uint8_t *lb_aunts, const uint8_t *start_headers,
                                             const uint8_t *end_headers, const uint64_t *start_blocks,
                                             const uint8_t *start_header, const uint64_t *end_blocks,
                                             const uint8_t *end_header, uint8_t *map_digests, uint8_t *map_subchains,
                                             uint8_t *reduce_digests, uint8_t *reduce_nodes, uint8_t *data_commitments,
                                             uint32_t *fail) {
    BSX_REQUIRE(ctx, ctx && dh_leaf && dh_aunts && lb_leaf && lb_aunts && start_headers && end_headers && start_blocks &&
                         start_header && end_blocks && end_header && map_digests && map_subchains && data_commitments);
    BSX_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(map_digests) | reinterpret_cast<uintptr_t>(map_subchains)) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(start_headers) & 3) == 0);
    BSX_REQUIRE(ctx, n_jobs >= 1 && (n_jobs & (n_jobs - 1)) == 0);
    if (n_ranges == 0) return BSX_OK;
    SubchainArgs a{dh_leaf, dh_aunts, lb_leaf, lb_aunts, start_headers, end_headers, nullptr, nullptr, nullptr, nullptr,
                   n_jobs, start_blocks, end_blocks, end_header, map_digests, map_subchains, BSX_SUBCHAIN_ARGS_NO_P2P};
    int rc = dispatch_subchain(ctx, (cudaStream_t)stream, B, n_ranges * n_jobs, a);
    if (rc) return rc;
    return bsx_reduce_subchains_dev(ctx, stream, n_ranges, n_jobs, map_subchains, start_blocks, start_header, end_blocks,
                                    end_header, B, reduce_digests, reduce_nodes, data_commitments, fail);
}

// ---- host-buffer entry points ----
namespace {
struct SubchainSizes {
    size_t dhl, aunt, lbl, hdr, dig, sub;
};
SubchainSizes subchain_sizes(uint32_t B, size_t jobs) {
    return {jobs * B * 34, jobs * B * 128, jobs * B * 72, jobs * 32, jobs * (size_t)(20 * B - 1) * 32, jobs * BSX_SUBCHAIN_BYTES};
}
}  // namespace

extern "C" int bsx_prove_subchain_batch(bsx_ctx *ctx, uint32_t B, uint32_t n_jobs, const uint8_t *dh_leaf,
                                        const uint8_t *dh_aunts, const uint8_t *lb_leaf, const uint8_t *lb_aunts,
                                        const uint8_t *start_headers, const uint8_t *end_headers,
                                        const uint64_t *batch_start, const uint64_t *batch_end,
                                        const uint64_t *global_end, const uint8_t *global_end_header, uint8_t *digests,
                                        uint8_t *subchains) {
    BSX_REQUIRE(ctx, ctx && dh_leaf && dh_aunts && lb_leaf && lb_aunts && start_headers && end_headers && batch_start &&
                         batch_end && global_end && global_end_header && digests && subchains);
    if (n_jobs == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    SubchainSizes s = subchain_sizes(B, n_jobs);
    int rc = ws_begin(ctx, ws_size(s.dhl) + 2 * ws_size(s.aunt) + ws_size(s.lbl) + 3 * ws_size(s.hdr) + 3 * ws_size(8 * (size_t)n_jobs) +
                               ws_size(s.dig) + ws_size(s.sub));
    if (rc) return rc;
    uint8_t *d_dhl = ws_take<uint8_t>(ctx, s.dhl), *d_dha = ws_take<uint8_t>(ctx, s.aunt);
    uint8_t *d_lbl = ws_take<uint8_t>(ctx, s.lbl), *d_lba = ws_take<uint8_t>(ctx, s.aunt);
    uint8_t *d_sh = ws_take<uint8_t>(ctx, s.hdr), *d_eh = ws_take<uint8_t>(ctx, s.hdr), *d_geh = ws_take<uint8_t>(ctx, s.hdr);
    uint64_t *d_bs = ws_take<uint64_t>(ctx, n_jobs), *d_be = ws_take<uint64_t>(ctx, n_jobs), *d_ge = ws_take<uint64_t>(ctx, n_jobs);
    uint8_t *d_dig = ws_take<uint8_t>(ctx, s.dig), *d_sub = ws_take<uint8_t>(ctx, s.sub);
    cudaStream_t st = ctx->stream;
    BSX_CUDA(ctx, cudaMemcpyAsync(d_dhl, dh_leaf, s.dhl, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_dha, dh_aunts, s.aunt, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_lbl, lb_leaf, s.lbl, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_lba, lb_aunts, s.aunt, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_sh, start_headers, s.hdr, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_eh, end_headers, s.hdr, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_geh, global_end_header, s.hdr, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_bs, batch_start, 8 * (size_t)n_jobs, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_be, batch_end, 8 * (size_t)n_jobs, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_ge, global_end, 8 * (size_t)n_jobs, cudaMemcpyHostToDevice, st));
    rc = bsx_prove_subchain_batch_dev(ctx, st, B, n_jobs, d_dhl, d_dha, d_lbl, d_lba, d_sh, d_eh, d_bs, d_be, d_ge, d_geh,
                                      d_dig, d_sub);
    if (rc) return rc;
    BSX_CUDA(ctx, cudaMemcpyAsync(digests, d_dig, s.dig, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(subchains, d_sub, s.sub, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaStreamSynchronize(st));
    return BSX_OK;
}

extern "C" int bsx_prove_data_commitment(bsx_ctx *ctx, uint32_t n_ranges, uint32_t n_jobs, uint32_t B,
                                         const uint8_t *dh_leaf, const uint8_t *dh_aunts, const uint8_t *lb_leaf,
                                         const uint8_t *lb_aunts, const uint8_t *start_headers,
                                         const uint8_t *end_headers, const uint64_t *start_blocks,
                                         const uint8_t *start_header, const uint64_t *end_blocks,
                                         const uint8_t *end_header, uint8_t *map_digests, uint8_t *map_subchains,
                                         uint8_t *reduce_digests, uint8_t *reduce_nodes, uint8_t *data_commitments,
                                         uint32_t *fail) {
    BSX_REQUIRE(ctx, ctx && dh_leaf && dh_aunts && lb_leaf && lb_aunts && start_headers && end_headers && start_blocks &&
                         start_header && end_blocks && end_header && data_commitments);
    BSX_REQUIRE(ctx, n_jobs >= 1 && (n_jobs & (n_jobs - 1)) == 0);
    if (n_ranges == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t jobs = (size_t)n_ranges * n_jobs, R = n_ranges;
    SubchainSizes s = subchain_sizes(B, jobs);
    const size_t s_rd = R * (n_jobs - 1) * 32, s_rn = R * (n_jobs - 1) * BSX_SUBCHAIN_BYTES;
    int rc = ws_begin(ctx, ws_size(s.dhl) + 2 * ws_size(s.aunt) + ws_size(s.lbl) + 2 * ws_size(s.hdr) + 3 * ws_size(32 * R) +
                               2 * ws_size(8 * R) + ws_size(s.dig) + ws_size(s.sub) + ws_size(s_rd) + ws_size(s_rn) + ws_size(4 * R));
    if (rc) return rc;
    uint8_t *d_dhl = ws_take<uint8_t>(ctx, s.dhl), *d_dha = ws_take<uint8_t>(ctx, s.aunt);
    uint8_t *d_lbl = ws_take<uint8_t>(ctx, s.lbl), *d_lba = ws_take<uint8_t>(ctx, s.aunt);
    uint8_t *d_sh = ws_take<uint8_t>(ctx, s.hdr), *d_eh = ws_take<uint8_t>(ctx, s.hdr);
    uint8_t *d_rsh = ws_take<uint8_t>(ctx, 32 * R), *d_reh = ws_take<uint8_t>(ctx, 32 * R), *d_dc = ws_take<uint8_t>(ctx, 32 * R);
    uint64_t *d_sb = ws_take<uint64_t>(ctx, R), *d_eb = ws_take<uint64_t>(ctx, R);
    uint8_t *d_dig = ws_take<uint8_t>(ctx, s.dig), *d_sub = ws_take<uint8_t>(ctx, s.sub);
    uint8_t *d_rd = ws_take<uint8_t>(ctx, s_rd), *d_rn = ws_take<uint8_t>(ctx, s_rn);
    uint32_t *d_fail = ws_take<uint32_t>(ctx, R);
    cudaStream_t st = ctx->stream;
    BSX_CUDA(ctx, cudaMemcpyAsync(d_dhl, dh_leaf, s.dhl, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_dha, dh_aunts, s.aunt, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_lbl, lb_leaf, s.lbl, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_lba, lb_aunts, s.aunt, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_sh, start_headers, s.hdr, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_eh, end_headers, s.hdr, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_rsh, start_header, 32 * R, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_reh, end_header, 32 * R, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_sb, start_blocks, 8 * R, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_eb, end_blocks, 8 * R, cudaMemcpyHostToDevice, st));
    rc = bsx_prove_data_commitment_dev(ctx, st, n_ranges, n_jobs, B, d_dhl, d_dha, d_lbl, d_lba, d_sh, d_eh, d_sb, d_rsh, d_eb,
                                       d_reh, d_dig, d_sub, d_rd, d_rn, d_dc, d_fail);
    if (rc) return rc;
    if (map_digests) BSX_CUDA(ctx, cudaMemcpyAsync(map_digests, d_dig, s.dig, cudaMemcpyDeviceToHost, st));
    if (map_subchains) BSX_CUDA(ctx, cudaMemcpyAsync(map_subchains, d_sub, s.sub, cudaMemcpyDeviceToHost, st));
    if (reduce_digests && s_rd) BSX_CUDA(ctx, cudaMemcpyAsync(reduce_digests, d_rd, s_rd, cudaMemcpyDeviceToHost, st));
    if (reduce_nodes && s_rn) BSX_CUDA(ctx, cudaMemcpyAsync(reduce_nodes, d_rn, s_rn, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(data_commitments, d_dc, 32 * R, cudaMemcpyDeviceToHost, st));
    if (fail) BSX_CUDA(ctx, cudaMemcpyAsync(fail, d_fail, 4 * R, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaStreamSynchronize(st));
    return BSX_OK;
}

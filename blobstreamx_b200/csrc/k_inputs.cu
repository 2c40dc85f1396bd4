// Input shaping on the device: the 14-leaf Tendermint tree of every header of a range and the Merkle
// inclusion proofs the map circuits consume, written straight into the layout of bsx_range_batch.
// Replaces the per-header CPU work of
//   generate_proofs_from_header / compute_hash_from_aunts      TX/input/tendermint_utils.rs:214-224, 276-336, 374-393
//   DataCommitmentInputs::get_data_commitment_inputs           BX/circuits/input.rs:149-271
//   the 32 DataCommitmentOffchainInputs hints of one range     BX/circuits/builder.rs:316-333
// which is 27 SHA-256 calls (41 compressions) per header on the host -- three orders of magnitude slower than the map
// kernels that consume the proofs.  The protobuf field encoders stay on the host (byte shuffling, no hashing); the
// kernel takes the 14 encoded fields of each header as one 512-byte record.
//
// One thread per header.  14 leaves split 8 | 6 -> (4 | 4) | (4 | 2) (split point = largest power of two below n,
// tendermint_utils.rs:338-349): levels l0[14], l1[7], l2[4] (l2[3] = l1[6] promoted), l3[2], root; leaves 0..11 sit at
// depth 4 and their aunts are l0[i^1], l1[(i>>1)^1], l2[(i>>2)^1], l3[(i>>3)^1].
#include "common.cuh"
#include "sha256.cuh"

namespace bsx {

struct HeaderLevels {
    uint32_t d[27][8];   // l0: 0..13, l1: 14..20, l2: 21..24, l3: 25..26
};

// record: bytes [0,14) = field lengths, [16, 16 + sum) = the encoded fields back to back
__device__ __forceinline__ void header_tree(const uint8_t *__restrict__ rec, HeaderLevels &L, uint32_t root[8], uint32_t off[15]) {
    uint32_t o = 16;
#pragma unroll 1
    for (int i = 0; i < 14; i++) {
        off[i] = o;
        const uint32_t len = rec[i];
        const uint8_t *p = rec + o;
        tm_leaf_hash([&](uint32_t k) -> uint8_t { return p[k]; }, len, L.d[i]);
        o += len;
    }
    off[14] = o;
#pragma unroll 1
    for (int i = 0; i < 7; i++) tm_inner_hash(L.d[2 * i], L.d[2 * i + 1], L.d[14 + i]);
#pragma unroll 1
    for (int i = 0; i < 3; i++) tm_inner_hash(L.d[14 + 2 * i], L.d[15 + 2 * i], L.d[21 + i]);
#pragma unroll
    for (int k = 0; k < 8; k++) L.d[24][k] = L.d[20][k];
    tm_inner_hash(L.d[21], L.d[22], L.d[25]);
    tm_inner_hash(L.d[23], L.d[24], L.d[26]);
    tm_inner_hash(L.d[25], L.d[26], root);
}
__device__ __forceinline__ void put_digest(uint8_t *p, const uint32_t d[8]) {   // any alignment >= 4
    uint32_t *q = reinterpret_cast<uint32_t *>(p);
#pragma unroll
    for (int k = 0; k < 8; k++) q[k] = bswap32(d[k]);
}
__device__ __forceinline__ void put_aunts(uint8_t *p, const HeaderLevels &L, int idx) {
    put_digest(p, L.d[idx ^ 1]);
    put_digest(p + 32, L.d[14 + ((idx >> 1) ^ 1)]);
    put_digest(p + 64, L.d[21 + ((idx >> 2) ^ 1)]);
    put_digest(p + 96, L.d[25 + ((idx >> 3) ^ 1)]);
}

// roots (and optionally all 27 level digests) of n headers
__global__ void __launch_bounds__(128) header_trees_kernel(const uint8_t *__restrict__ headers, uint32_t n, uint8_t *__restrict__ roots,
                                                           uint8_t *__restrict__ levels) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    HeaderLevels L;
    uint32_t root[8], off[15];
    header_tree(headers + (size_t)BSX_HEADER_LEAVES_BYTES * i, L, root, off);
    put_digest(roots + 32 * (size_t)i, root);
    if (levels)
        for (int k = 0; k < 27; k++) put_digest(levels + ((size_t)i * 27 + k) * 32, L.d[k]);
}

struct RangeInputsArgs {
    const uint8_t *headers;          // n_ranges * (J*B + 1) records: blocks start .. start + J*B
    const uint64_t *start_blocks, *end_blocks, *latest_blocks;   // latest_blocks == nullptr: the range ends at the chain tip
    uint8_t *dh_leaf, *dh_aunts, *lb_leaf, *lb_aunts, *start_headers, *end_headers, *start_header, *end_header;
    uint32_t *fail;
    uint32_t n_ranges, J, B;
};

// thread (r, o): header at block start_r + o.  The hint of job j (DataCommitmentOffchainInputs, BX/circuits/data_commitment.rs:
// 22-44) is called with (start + jB, start + (j+1)B) and clamps only to the last fetchable block `latest` (latest_block - 2,
// BX/circuits/input.rs:160-163), not to the range's end: with avail = min(latest - start, J*B)
//   data_hash proof (leaf 6, 34 B)      -> slot o      for o <  avail       (blocks [bs, req_end),  input.rs:165-181)
//   last_block_id proof (leaf 4, 72 B)  -> slot o - 1  for 1 <= o <= avail  (blocks (bs, req_end],  :183-197)
//   job j = o / B starts at root(o = jB) and ends at root(min((j+1)B, avail)) when jB < avail (:247-262); everything past
//   `avail` stays zero (:221-241).  The range's own start / end header hashes are taken at o = 0 and o = end - start.
__global__ void __launch_bounds__(128) range_inputs_kernel(RangeInputsArgs a) {
    const uint32_t per = a.J * a.B + 1;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)a.n_ranges * per) return;
    const uint32_t r = (uint32_t)(t / per), o = (uint32_t)(t % per);
    const uint64_t sb = a.start_blocks[r], eb = a.end_blocks[r], lt = a.latest_blocks ? a.latest_blocks[r] : eb;
    const uint64_t span = lt > sb ? lt - sb : 0;
    const uint32_t avail = span < (uint64_t)(per - 1) ? (uint32_t)span : per - 1;
    if (o > avail || avail == 0) return;
    const uint8_t *rec = a.headers + (size_t)BSX_HEADER_LEAVES_BYTES * t;
    HeaderLevels L;
    uint32_t root[8], off[15];
    header_tree(rec, L, root, off);
    const size_t slot0 = (size_t)r * (per - 1);
    if (o < avail) {
        uint8_t *leaf = a.dh_leaf + (slot0 + o) * 34;
        if (rec[6] != 34) atomicOr(a.fail + r, BSX_FAIL_INPUT_LEAF);
        for (uint32_t k = 0; k < 34; k++) leaf[k] = k < rec[6] ? rec[off[6] + k] : (uint8_t)0;
        put_aunts(a.dh_aunts + (slot0 + o) * 128, L, 6);
        if (o % a.B == 0) put_digest(a.start_headers + ((size_t)r * a.J + o / a.B) * 32, root);
    }
    if (o >= 1) {
        uint8_t *leaf = a.lb_leaf + (slot0 + o - 1) * 72;
        if (rec[4] != 72) atomicOr(a.fail + r, BSX_FAIL_INPUT_LEAF);
        for (uint32_t k = 0; k < 72; k++) leaf[k] = k < rec[4] ? rec[off[4] + k] : (uint8_t)0;
        put_aunts(a.lb_aunts + (slot0 + o - 1) * 128, L, 4);
        if (o % a.B == 0 || o == avail) put_digest(a.end_headers + ((size_t)r * a.J + (o - 1) / a.B) * 32, root);
    }
    if (eb > sb && eb - sb <= (uint64_t)avail) {
        if (o == 0) put_digest(a.start_header + 32 * (size_t)r, root);
        if ((uint64_t)o == eb - sb) put_digest(a.end_header + 32 * (size_t)r, root);
    }
}

}  // namespace bsx

using namespace bsx;

extern "C" int bsx_header_trees_dev(bsx_ctx *ctx, void *stream, const uint8_t *headers, uint32_t n, uint8_t *roots, uint8_t *levels) {
    BSX_REQUIRE(ctx, ctx && headers && roots);
    BSX_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(roots) | reinterpret_cast<uintptr_t>(levels)) & 3) == 0);
    if (n == 0) return BSX_OK;
    BSX_PIN_CARVEOUT(header_trees_kernel);
    header_trees_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(headers, n, roots, levels);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

extern "C" int bsx_header_range_inputs_dev(bsx_ctx *ctx, void *stream, uint32_t n_ranges, uint32_t n_jobs, uint32_t B,
                                           const uint8_t *headers, const uint64_t *start_blocks, const uint64_t *end_blocks,
                                           const uint64_t *latest_blocks, uint8_t *dh_leaf, uint8_t *dh_aunts, uint8_t *lb_leaf, uint8_t *lb_aunts,
                                           uint8_t *start_headers, uint8_t *end_headers, uint8_t *start_header, uint8_t *end_header,
                                           uint32_t *fail) {
    BSX_REQUIRE(ctx, ctx && headers && start_blocks && end_blocks && dh_leaf && dh_aunts && lb_leaf && lb_aunts && start_headers &&
                         end_headers && start_header && end_header && fail);
    BSX_REQUIRE(ctx, n_jobs >= 1 && B >= 1 && (uint64_t)n_jobs * B < (1u << 24));
    BSX_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(dh_aunts) | reinterpret_cast<uintptr_t>(lb_aunts) | reinterpret_cast<uintptr_t>(start_headers) |
                       reinterpret_cast<uintptr_t>(end_headers) | reinterpret_cast<uintptr_t>(start_header) |
                       reinterpret_cast<uintptr_t>(end_header)) & 3) == 0);
    if (n_ranges == 0) return BSX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t slots = (size_t)n_ranges * n_jobs * B, jobs = (size_t)n_ranges * n_jobs;
    // slots and job headers beyond the last fetchable block are zero (input.rs:221-241, 247-262)
    BSX_CUDA(ctx, cudaMemsetAsync(dh_leaf, 0, slots * 34, st));
    BSX_CUDA(ctx, cudaMemsetAsync(dh_aunts, 0, slots * 128, st));
    BSX_CUDA(ctx, cudaMemsetAsync(lb_leaf, 0, slots * 72, st));
    BSX_CUDA(ctx, cudaMemsetAsync(lb_aunts, 0, slots * 128, st));
    BSX_CUDA(ctx, cudaMemsetAsync(start_headers, 0, jobs * 32, st));
    BSX_CUDA(ctx, cudaMemsetAsync(end_headers, 0, jobs * 32, st));
    BSX_CUDA(ctx, cudaMemsetAsync(start_header, 0, (size_t)n_ranges * 32, st));
    BSX_CUDA(ctx, cudaMemsetAsync(end_header, 0, (size_t)n_ranges * 32, st));
    BSX_CUDA(ctx, cudaMemsetAsync(fail, 0, (size_t)n_ranges * 4, st));
    RangeInputsArgs a{headers, start_blocks, end_blocks, latest_blocks, dh_leaf, dh_aunts, lb_leaf, lb_aunts, start_headers, end_headers,
                      start_header, end_header, fail, n_ranges, n_jobs, B};
    const size_t threads = (size_t)n_ranges * ((size_t)n_jobs * B + 1);
    BSX_PIN_CARVEOUT(range_inputs_kernel);
    range_inputs_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(a);
    BSX_LAUNCHED(ctx);
    return BSX_OK;
}

// ---- host-buffer forms ----
extern "C" int bsx_header_trees(bsx_ctx *ctx, const uint8_t *headers, uint32_t n, uint8_t *roots, uint8_t *levels) {
    BSX_REQUIRE(ctx, ctx && headers && roots);
    if (n == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t N = n, s_in = N * BSX_HEADER_LEAVES_BYTES;
    int rc = ws_begin(ctx, ws_size(s_in) + ws_size(32 * N) + ws_size(27 * 32 * N));
    if (rc) return rc;
    uint8_t *d_in = ws_take<uint8_t>(ctx, s_in), *d_roots = ws_take<uint8_t>(ctx, 32 * N), *d_lv = ws_take<uint8_t>(ctx, 27 * 32 * N);
    cudaStream_t st = ctx->stream;
    BSX_CUDA(ctx, cudaMemcpyAsync(d_in, headers, s_in, cudaMemcpyHostToDevice, st));
    rc = bsx_header_trees_dev(ctx, st, d_in, n, d_roots, levels ? d_lv : nullptr);
    if (rc) return rc;
    BSX_CUDA(ctx, cudaMemcpyAsync(roots, d_roots, 32 * N, cudaMemcpyDeviceToHost, st));
    if (levels) BSX_CUDA(ctx, cudaMemcpyAsync(levels, d_lv, 27 * 32 * N, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaStreamSynchronize(st));
    return BSX_OK;
}

extern "C" int bsx_header_range_inputs(bsx_ctx *ctx, uint32_t n_ranges, uint32_t n_jobs, uint32_t B, const uint8_t *headers,
                                       const uint64_t *start_blocks, const uint64_t *end_blocks, const uint64_t *latest_blocks,
                                       uint8_t *dh_leaf, uint8_t *dh_aunts,
                                       uint8_t *lb_leaf, uint8_t *lb_aunts, uint8_t *start_headers, uint8_t *end_headers,
                                       uint8_t *start_header, uint8_t *end_header, uint32_t *fail) {
    BSX_REQUIRE(ctx, ctx && headers && start_blocks && end_blocks && dh_leaf && dh_aunts && lb_leaf && lb_aunts && start_headers &&
                         end_headers && start_header && end_header && fail);
    BSX_REQUIRE(ctx, n_jobs >= 1 && B >= 1 && (uint64_t)n_jobs * B < (1u << 24));
    if (n_ranges == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t R = n_ranges, slots = R * n_jobs * B, jobs = R * n_jobs, s_in = R * ((size_t)n_jobs * B + 1) * BSX_HEADER_LEAVES_BYTES;
    int rc = ws_begin(ctx, ws_size(s_in) + 3 * ws_size(8 * R) + ws_size(slots * 34) + 2 * ws_size(slots * 128) + ws_size(slots * 72) +
                               2 * ws_size(jobs * 32) + 2 * ws_size(32 * R) + ws_size(4 * R));
    if (rc) return rc;
    uint8_t *d_in = ws_take<uint8_t>(ctx, s_in);
    uint64_t *d_sb = ws_take<uint64_t>(ctx, R), *d_eb = ws_take<uint64_t>(ctx, R), *d_lt = ws_take<uint64_t>(ctx, R);
    uint8_t *d_dhl = ws_take<uint8_t>(ctx, slots * 34), *d_dha = ws_take<uint8_t>(ctx, slots * 128);
    uint8_t *d_lbl = ws_take<uint8_t>(ctx, slots * 72), *d_lba = ws_take<uint8_t>(ctx, slots * 128);
    uint8_t *d_sh = ws_take<uint8_t>(ctx, jobs * 32), *d_eh = ws_take<uint8_t>(ctx, jobs * 32);
    uint8_t *d_rsh = ws_take<uint8_t>(ctx, 32 * R), *d_reh = ws_take<uint8_t>(ctx, 32 * R);
    uint32_t *d_fail = ws_take<uint32_t>(ctx, R);
    cudaStream_t st = ctx->stream;
    BSX_CUDA(ctx, cudaMemcpyAsync(d_in, headers, s_in, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_sb, start_blocks, 8 * R, cudaMemcpyHostToDevice, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(d_eb, end_blocks, 8 * R, cudaMemcpyHostToDevice, st));
    if (latest_blocks) BSX_CUDA(ctx, cudaMemcpyAsync(d_lt, latest_blocks, 8 * R, cudaMemcpyHostToDevice, st));
    rc = bsx_header_range_inputs_dev(ctx, st, n_ranges, n_jobs, B, d_in, d_sb, d_eb, latest_blocks ? d_lt : nullptr, d_dhl, d_dha, d_lbl, d_lba, d_sh, d_eh, d_rsh, d_reh,
                                     d_fail);
    if (rc) return rc;
    BSX_CUDA(ctx, cudaMemcpyAsync(dh_leaf, d_dhl, slots * 34, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(dh_aunts, d_dha, slots * 128, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(lb_leaf, d_lbl, slots * 72, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(lb_aunts, d_lba, slots * 128, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(start_headers, d_sh, jobs * 32, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(end_headers, d_eh, jobs * 32, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(start_header, d_rsh, 32 * R, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(end_header, d_reh, 32 * R, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaMemcpyAsync(fail, d_fail, 4 * R, cudaMemcpyDeviceToHost, st));
    BSX_CUDA(ctx, cudaStreamSynchronize(st));
    return BSX_OK;
}

// header_range = skip + prove_data_commitment (CombinedSkipCircuit::define, BX/circuits/header_range.rs:32-59) for
// n independent ranges in one call.  The Ed25519/verify_skip kernels (FMA-pipe bound) and the map/reduce SHA-256
// kernels (ALU-pipe bound) are independent until the caller reads the results, so they are issued on two streams
// and joined with an event: the two halves overlap on the SMs instead of running back to back.
#include "common.cuh"

static int fork_join_begin(bsx_ctx *ctx, cudaStream_t main) {
    BSX_CUDA(ctx, cudaEventRecord(ctx->ev_fork, main));
    BSX_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    return BSX_OK;
}
static int fork_join_end(bsx_ctx *ctx, cudaStream_t main) {
    BSX_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->stream2));
    BSX_CUDA(ctx, cudaStreamWaitEvent(main, ctx->ev_join, 0));
    return BSX_OK;
}

extern "C" int bsx_header_range_dev(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, uint32_t n_jobs, uint32_t B,
                                    const bsx_skip_batch *s, const bsx_range_batch *m) {
    BSX_REQUIRE(ctx, ctx && s && m);
    BSX_REQUIRE(ctx, s->hdr && s->validators && s->skip && s->trusted_pubkeys && s->trusted_powers && s->trusted_byte_lengths &&
                         s->digests && s->ed_out && s->fail);
    if (n == 0) return BSX_OK;
    cudaStream_t main = (cudaStream_t)stream;
    int rc = fork_join_begin(ctx, main);
    if (rc) return rc;
    // high-priority stream: the Ed25519 kernel only (few fat CTAs, latency-bound)
    rc = bsx_verify_launch_ed(ctx, ctx->stream2, n, N, s->validators, s->ed_out);
    if (rc) return rc;
    // caller's stream: every SHA-256 kernel -- the skip schedule, then map and reduce
    rc = bsx_verify_launch_hash(ctx, main, 1, n, N, s->hdr, s->validators, s->skip, s->trusted_pubkeys, s->trusted_powers,
                                s->trusted_byte_lengths, nullptr, s->digests, nullptr, s->fail);
    if (rc) return rc;
    rc = bsx_prove_data_commitment_dev(ctx, main, n, n_jobs, B, m->dh_leaf, m->dh_aunts, m->lb_leaf, m->lb_aunts,
                                       m->start_headers, m->end_headers, m->start_blocks, m->start_header, m->end_blocks,
                                       m->end_header, m->map_digests, m->map_subchains, m->reduce_digests, m->reduce_nodes,
                                       m->data_commitments, m->fail);
    if (rc) return rc;
    rc = fork_join_end(ctx, main);
    if (rc) return rc;
    return bsx_verify_launch_flags(ctx, main, n, N, s->ed_out, s->fail);
}

namespace {
struct Stage {
    bsx_ctx *ctx;
    cudaStream_t st;
    int rc = BSX_OK;
    // device copy of a host input array
    template <typename T>
    const T *in(const T *h, size_t count) {
        T *d = bsx::ws_take<T>(ctx, count);
        if (rc == BSX_OK && count && cudaMemcpyAsync(d, h, count * sizeof(T), cudaMemcpyHostToDevice, st) != cudaSuccess)
            rc = bsx::fail(ctx, BSX_ERR_CUDA, "cudaMemcpyAsync H2D failed%s%s");
        return d;
    }
    template <typename T>
    T *out(size_t count) { return bsx::ws_take<T>(ctx, count); }
    template <typename T>
    void back(T *h, const T *d, size_t count) {
        if (rc == BSX_OK && h && count && cudaMemcpyAsync(h, d, count * sizeof(T), cudaMemcpyDeviceToHost, st) != cudaSuccess)
            rc = bsx::fail(ctx, BSX_ERR_CUDA, "cudaMemcpyAsync D2H failed%s%s");
    }
};
}  // namespace

// host buffers: H2D of the skip half on stream2 and of the map half on the ctx stream, kernels, D2H on the same
// streams (copies of one half overlap kernels of the other), one synchronize at the end.
extern "C" int bsx_header_range(bsx_ctx *ctx, uint32_t n, uint32_t N, uint32_t n_jobs, uint32_t B, const bsx_skip_batch *s,
                                const bsx_range_batch *m) {
    BSX_REQUIRE(ctx, ctx && s && m);
    BSX_REQUIRE(ctx, s->hdr && s->validators && s->skip && s->trusted_pubkeys && s->trusted_powers && s->trusted_byte_lengths &&
                         s->digests && s->ed_out && s->fail);
    BSX_REQUIRE(ctx, m->dh_leaf && m->dh_aunts && m->lb_leaf && m->lb_aunts && m->start_headers && m->end_headers &&
                         m->start_blocks && m->start_header && m->end_blocks && m->end_header && m->data_commitments && m->fail);
    BSX_REQUIRE(ctx, n_jobs >= 1 && (n_jobs & (n_jobs - 1)) == 0 && N >= 1);
    if (n == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    using bsx::ws_size;
    const size_t R = n, nN = R * N, jobs = R * n_jobs, D = bsx_verify_digest_count(1, N);
    const size_t s_dhl = jobs * B * 34, s_aunt = jobs * B * 128, s_lbl = jobs * B * 72, s_dig = jobs * (size_t)(20 * B - 1) * 32,
                 s_sub = jobs * BSX_SUBCHAIN_BYTES, s_rd = R * (n_jobs - 1) * 32, s_rn = R * (n_jobs - 1) * BSX_SUBCHAIN_BYTES;
    size_t total = ws_size(R * sizeof(bsx_header_in)) + ws_size(nN * BSX_VAL_IN_BYTES) + ws_size(R * sizeof(bsx_skip_in)) +
                   ws_size(32 * nN) + ws_size(8 * nN) + ws_size(4 * nN) + ws_size(R * D * 32) + ws_size(nN * BSX_SIG_OUT_BYTES) +
                   ws_size(4 * R) + ws_size(s_dhl) + 2 * ws_size(s_aunt) + ws_size(s_lbl) + 2 * ws_size(jobs * 32) +
                   2 * ws_size(8 * R) + 3 * ws_size(32 * R) + ws_size(s_dig) + ws_size(s_sub) + ws_size(s_rd) + ws_size(s_rn) +
                   ws_size(4 * R);
    int rc = bsx::ws_begin(ctx, total);
    if (rc) return rc;
    cudaStream_t main = ctx->stream;
    rc = fork_join_begin(ctx, main);
    if (rc) return rc;
    // ---- skip half on stream2 ----
    Stage a{ctx, ctx->stream2};
    auto *d_hdr = a.in(s->hdr, R);
    auto *d_val = a.in(s->validators, nN * BSX_VAL_IN_BYTES);
    auto *d_skip = a.in(s->skip, R);
    auto *d_tpk = a.in(s->trusted_pubkeys, 32 * nN);
    auto *d_tpw = a.in(s->trusted_powers, nN);
    auto *d_tbl = a.in(s->trusted_byte_lengths, nN);
    auto *d_sdig = a.out<uint8_t>(R * D * 32);
    auto *d_ed = a.out<uint8_t>(nN * BSX_SIG_OUT_BYTES);
    auto *d_sfail = a.out<uint32_t>(R);
    if (a.rc) return a.rc;
    rc = bsx_verify_skip_dev(ctx, ctx->stream2, n, N, d_hdr, d_val, d_skip, d_tpk, d_tpw, d_tbl, d_sdig, d_ed, d_sfail);
    if (rc) return rc;
    a.back(s->digests, d_sdig, R * D * 32);
    a.back(s->ed_out, d_ed, nN * BSX_SIG_OUT_BYTES);
    a.back(s->fail, d_sfail, R);
    if (a.rc) return a.rc;
    // ---- map + reduce half on the ctx stream ----
    Stage b{ctx, main};
    auto *d_dhl = b.in(m->dh_leaf, s_dhl);
    auto *d_dha = b.in(m->dh_aunts, s_aunt);
    auto *d_lbl = b.in(m->lb_leaf, s_lbl);
    auto *d_lba = b.in(m->lb_aunts, s_aunt);
    auto *d_sh = b.in(m->start_headers, jobs * 32);
    auto *d_eh = b.in(m->end_headers, jobs * 32);
    auto *d_sb = b.in(m->start_blocks, R);
    auto *d_eb = b.in(m->end_blocks, R);
    auto *d_rsh = b.in(m->start_header, 32 * R);
    auto *d_reh = b.in(m->end_header, 32 * R);
    auto *d_dig = b.out<uint8_t>(s_dig);
    auto *d_sub = b.out<uint8_t>(s_sub);
    auto *d_rd = b.out<uint8_t>(s_rd);
    auto *d_rn = b.out<uint8_t>(s_rn);
    auto *d_dc = b.out<uint8_t>(32 * R);
    auto *d_fail = b.out<uint32_t>(R);
    if (b.rc) return b.rc;
    rc = bsx_prove_data_commitment_dev(ctx, main, n, n_jobs, B, d_dhl, d_dha, d_lbl, d_lba, d_sh, d_eh, d_sb, d_rsh, d_eb, d_reh,
                                       d_dig, d_sub, d_rd, d_rn, d_dc, d_fail);
    if (rc) return rc;
    b.back(m->map_digests, d_dig, s_dig);
    b.back(m->map_subchains, d_sub, s_sub);
    b.back(m->reduce_digests, d_rd, s_rd);
    b.back(m->reduce_nodes, d_rn, s_rn);
    b.back(m->data_commitments, d_dc, 32 * R);
    b.back(m->fail, d_fail, R);
    if (b.rc) return b.rc;
    rc = fork_join_end(ctx, main);
    if (rc) return rc;
    BSX_CUDA(ctx, cudaStreamSynchronize(main));
    return BSX_OK;
}

// header_range = skip + prove_data_commitment (CombinedSkipCircuit::define, BX/circuits/header_range.rs:32-59) for
// n independent ranges in one call.  The Ed25519/verify_skip kernels (FMA-pipe bound) and the map/reduce SHA-256
// kernels (ALU-pipe bound) are independent until the caller reads the results, so they are issued on two streams
// and joined with an event: the two halves overlap on the SMs instead of running back to back.
#include "common.cuh"

#include <stdlib.h>

#include <algorithm>
#include <vector>

// side streams of the device-resident step: stream2 (high priority) and pipe[0]
static int fork_join_begin(bsx_ctx *ctx, cudaStream_t main) {
    BSX_CUDA(ctx, cudaEventRecord(ctx->ev_fork, main));
    BSX_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    BSX_CUDA(ctx, cudaStreamWaitEvent(ctx->pipe[0], ctx->ev_fork, 0));
    return BSX_OK;
}
static int fork_join_end(bsx_ctx *ctx, cudaStream_t main) {
    BSX_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->stream2));
    BSX_CUDA(ctx, cudaStreamWaitEvent(main, ctx->ev_join, 0));
    BSX_CUDA(ctx, cudaEventRecord(ctx->ev_pipe[0], ctx->pipe[0]));
    BSX_CUDA(ctx, cudaStreamWaitEvent(main, ctx->ev_pipe[0], 0));
    return BSX_OK;
}

extern "C" int bsx_header_range_dev(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, uint32_t n_jobs, uint32_t B,
                                    const bsx_skip_batch *s, const bsx_range_batch *m) {
    BSX_REQUIRE(ctx, ctx && s && m);
    BSX_REQUIRE(ctx, s->hdr && s->validators && s->skip && s->trusted_pubkeys && s->trusted_powers && s->trusted_byte_lengths &&
                         s->digests && s->ed_out && s->fail);
    if (n == 0) return BSX_OK;
    cudaStream_t main = (cudaStream_t)stream;
    // BSX_HR_TRACE=1 (diagnostic, synchronises): when each half of the step finishes, on stderr
    const bool trace = ctx->tun[BSX_TUN_HR_TRACE] != 0;
    cudaEvent_t tev[4] = {};
    if (trace) {
        for (cudaEvent_t &e : tev) cudaEventCreate(&e);
        cudaEventRecord(tev[0], main);
    }
    int rc = fork_join_begin(ctx, main);
    if (rc) return rc;
    // high-priority stream: the Ed25519 kernel only (few fat CTAs, latency-bound)
    rc = bsx_verify_launch_ed(ctx, ctx->stream2, n, N, s->validators, s->ed_out);
    if (rc) return rc;
    if (trace) cudaEventRecord(tev[1], ctx->stream2);
    // third stream: the skip schedule's SHA-256 kernel (one CTA per range, 490 dependent digests: latency-bound).  On the
    // caller's stream it held the map kernels back until it had squeezed past the Ed25519 wave (1.7 ms instead of 0.19).
    // Only when the Ed25519 batch fills whole waves (common.cuh); otherwise it stays first on the caller's stream.
    const int hash_side = ctx->tun[BSX_TUN_HR_HASH_STREAM];   // -1 by wave fill, 0 caller's stream, 1 own
    cudaStream_t hs = (hash_side >= 0 ? hash_side != 0 : bsx_ed_fills_waves(ctx, (uint64_t)n * N)) ? ctx->pipe[0] : main;
    rc = bsx_verify_launch_hash(ctx, hs, 1, n, N, s->hdr, s->validators, s->skip, s->trusted_pubkeys, s->trusted_powers,
                                s->trusted_byte_lengths, nullptr, s->digests, nullptr, s->fail);
    if (rc) return rc;
    if (trace) cudaEventRecord(tev[2], hs);
    // caller's stream: map and reduce
    rc = bsx_prove_data_commitment_dev(ctx, main, n, n_jobs, B, m->dh_leaf, m->dh_aunts, m->lb_leaf, m->lb_aunts,
                                       m->start_headers, m->end_headers, m->start_blocks, m->start_header, m->end_blocks,
                                       m->end_header, m->map_digests, m->map_subchains, m->reduce_digests, m->reduce_nodes,
                                       m->data_commitments, m->fail);
    if (rc) return rc;
    if (trace) cudaEventRecord(tev[3], main);
    rc = fork_join_end(ctx, main);
    if (rc) return rc;
    rc = bsx_verify_launch_flags(ctx, main, n, N, s->ed_out, s->fail);
    if (trace) {
        cudaStreamSynchronize(main);
        float t[3];
        for (int i = 0; i < 3; i++) cudaEventElapsedTime(&t[i], tev[0], tev[i + 1]);
        fprintf(stderr, "[bsx header_range] Ed25519 done %.3f ms, skip hashes done %.3f ms, map + reduce done %.3f ms\n", t[0], t[1], t[2]);
        for (cudaEvent_t &e : tev) cudaEventDestroy(e);
    }
    return rc;
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

static int ensure_chunk_events(bsx_ctx *ctx, uint32_t need) {
    if (need <= ctx->n_ev_chunk) return BSX_OK;
    cudaEvent_t *p = (cudaEvent_t *)realloc(ctx->ev_chunk, sizeof(cudaEvent_t) * need);
    if (!p) return bsx::fail(ctx, BSX_ERR_NOMEM, "out of host memory%s%s");
    ctx->ev_chunk = p;
    while (ctx->n_ev_chunk < need) {
        BSX_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_chunk[ctx->n_ev_chunk], cudaEventDisableTiming));
        ctx->n_ev_chunk++;
    }
    return BSX_OK;
}

static int header_range_host(bsx_ctx *ctx, uint32_t n, uint32_t N, uint32_t n_jobs, uint32_t B, const bsx_skip_batch *s,
                             const bsx_range_batch *m);

// host buffers.  The map half is PCIe-bound (371 KB in, 0.66 MB of digests out per range of 1024 headers against ~3 us
// of kernel time), so it is cut into chunks of ranges that flow through three streams by role -- H2D copies, kernels,
// D2H copies -- chained with one event per chunk and stage: the upload of chunk k+1, the kernels of chunk k and the
// download of chunk k-1 overlap (the two copy engines run full duplex).  The skip half (1 MB in, ~73 KB out per range,
// Ed25519-bound) runs on stream2 beside them; all kernels share one shared-memory carveout (common.cuh) so that the
// map kernels are co-resident with the long-running Ed25519 CTAs instead of waiting for the SMs to drain (measured:
// without it a 0.1 ms chunk took 1.7 ms and the D2H engine idled behind it).  BSX_PIPE_ED=2 runs the skip half after
// the map kernels instead (A/B knob), BSX_PIPE_TRACE=1 prints the per-chunk timeline.
extern "C" int bsx_header_range(bsx_ctx *ctx, uint32_t n, uint32_t N, uint32_t n_jobs, uint32_t B, const bsx_skip_batch *s,
                                const bsx_range_batch *m) {
    BSX_REQUIRE(ctx, ctx && s && m);
    BSX_REQUIRE(ctx, s->hdr && s->validators && s->skip && s->trusted_pubkeys && s->trusted_powers && s->trusted_byte_lengths &&
                         s->digests && s->ed_out && s->fail);
    BSX_REQUIRE(ctx, m->dh_leaf && m->dh_aunts && m->lb_leaf && m->lb_aunts && m->start_headers && m->end_headers &&
                         m->start_blocks && m->start_header && m->end_blocks && m->end_header && m->data_commitments && m->fail);
    BSX_REQUIRE(ctx, n_jobs >= 1 && (n_jobs & (n_jobs - 1)) == 0 && N >= 1 && N <= 4096);
    BSX_REQUIRE(ctx, B >= 1 && B <= 256 && (B & (B - 1)) == 0);   // before any size below is derived from it
    if (n == 0) return BSX_OK;
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    const int rc_all = header_range_host(ctx, n, N, n_jobs, B, s, m);
    if (rc_all != BSX_OK) {
        // copies against the caller's buffers may still be in flight on any of the five streams: the caller must be free
        // to release or reuse them once this returns, whatever the status
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->stream2);
        for (int i = 0; i < BSX_PIPE_STREAMS; i++) cudaStreamSynchronize(ctx->pipe[i]);
    }
    return rc_all;
}

static int header_range_host(bsx_ctx *ctx, uint32_t n, uint32_t N, uint32_t n_jobs, uint32_t B, const bsx_skip_batch *s,
                             const bsx_range_batch *m) {
    using bsx::ws_size;
    const size_t R = n, nN = R * N, D = bsx_verify_digest_count(1, N);
    // chunking of the map half: the download of the digests is the longest stage, so the first chunks are small (the
    // D2H engine starts early) and the size doubles up to n/8: for n = 256 the chunks are 8, 16, 32, 32, ... ranges
    const uint32_t forced = (uint32_t)ctx->tun[BSX_TUN_PIPE_CHUNK];
    const int ed_order = ctx->tun[BSX_TUN_PIPE_ED];   // 1 first, 2 last
    std::vector<uint32_t> chunk_at;   // first range of each chunk, then n
    {
        uint32_t big = forced ? forced : (n + 7) / 8, cur = forced ? forced : (n + 31) / 32;
        if (big < 4) big = 4;
        if (cur < 4) cur = 4;
        for (uint32_t r = 0; r < n;) {
            chunk_at.push_back(r);
            r += cur < big ? cur : big;
            cur *= 2;
        }
        chunk_at.push_back(n);
    }
    const uint32_t n_chunks = (uint32_t)chunk_at.size() - 1;
    uint32_t chunk = 0;               // largest chunk: sizes the per-chunk device buffers
    for (uint32_t k = 0; k < n_chunks; k++) chunk = std::max(chunk, std::min(chunk_at[k + 1], n) - chunk_at[k]);
    const bool ed_last = ed_order == 2;
    int rc = ensure_chunk_events(ctx, 2 * n_chunks);
    if (rc) return rc;
    // per-range sizes of the map half
    const size_t j1 = n_jobs, r_dhl = j1 * B * 34, r_aunt = j1 * B * 128, r_lbl = j1 * B * 72, r_dig = j1 * (size_t)(20 * B - 1) * 32,
                 r_sub = j1 * BSX_SUBCHAIN_BYTES, r_rd = (j1 - 1) * 32, r_rn = (j1 - 1) * BSX_SUBCHAIN_BYTES;
    const size_t c = chunk, jobs = R * j1;
    const size_t per_chunk = ws_size(c * r_dhl) + 2 * ws_size(c * r_aunt) + ws_size(c * r_lbl) + ws_size(c * r_dig);
    size_t total = ws_size(R * sizeof(bsx_header_in)) + ws_size(nN * BSX_VAL_IN_BYTES) + ws_size(R * sizeof(bsx_skip_in)) +
                   ws_size(32 * nN) + ws_size(8 * nN) + ws_size(4 * nN) + ws_size(R * D * 32) + ws_size(nN * BSX_SIG_OUT_BYTES) +
                   ws_size(4 * R) + n_chunks * per_chunk + 2 * ws_size(jobs * 32) + 2 * ws_size(8 * R) + 3 * ws_size(32 * R) +
                   ws_size(R * r_sub) + ws_size(R * r_rd + 16) + ws_size(R * r_rn + 16) + ws_size(4 * R);
    rc = bsx::ws_begin(ctx, total);
    if (rc) return rc;
    cudaStream_t main = ctx->stream, s_up = ctx->pipe[0], s_k = ctx->pipe[1], s_down = ctx->pipe[2];
    // every worker stream starts after whatever the caller queued on the ctx stream
    BSX_CUDA(ctx, cudaEventRecord(ctx->ev_fork, main));
    BSX_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    for (int i = 0; i < BSX_PIPE_STREAMS; i++) BSX_CUDA(ctx, cudaStreamWaitEvent(ctx->pipe[i], ctx->ev_fork, 0));
    const bool trace = ctx->tun[BSX_TUN_PIPE_TRACE] != 0;   // diagnostic: per-chunk timeline on stderr
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        tev.push_back(e);
    };
    mark(main);
    // ---- skip half on stream2 (upload now; kernels now or after the map kernels) ----
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
    auto run_skip = [&]() -> int {
        // co-run register budget: the chunks' map kernels must get onto the SMs while the Ed25519 wave is resident
        int r = bsx_verify_skip_corun_dev(ctx, ctx->stream2, n, N, d_hdr, d_val, d_skip, d_tpk, d_tpw, d_tbl, d_sdig, d_ed, d_sfail);
        if (r) return r;
        a.back(s->digests, d_sdig, R * D * 32);
        a.back(s->ed_out, d_ed, nN * BSX_SIG_OUT_BYTES);
        a.back(s->fail, d_sfail, R);
        return a.rc;
    };
    if (!ed_last) {
        rc = run_skip();
        if (rc) return rc;
    }
    // ---- map + reduce half: chunks of ranges through upload -> kernels -> download ----
    // the small per-job / per-range arrays go up once for the whole batch and come back once at the end; only the
    // proofs (in) and the digests (out) are moved per chunk, as four and one large copies
    Stage up{ctx, s_up}, down{ctx, s_down};
    auto *d_sh = up.in(m->start_headers, jobs * 32);
    auto *d_eh = up.in(m->end_headers, jobs * 32);
    auto *d_sb = up.in(m->start_blocks, R);
    auto *d_eb = up.in(m->end_blocks, R);
    auto *d_rsh = up.in(m->start_header, 32 * R);
    auto *d_reh = up.in(m->end_header, 32 * R);
    auto *d_sub = up.out<uint8_t>(R * r_sub);
    auto *d_rd = up.out<uint8_t>(R * r_rd + 16);
    auto *d_rn = up.out<uint8_t>(R * r_rn + 16);
    auto *d_dc = up.out<uint8_t>(32 * R);
    auto *d_fail = up.out<uint32_t>(R);
    for (uint32_t k = 0; k < n_chunks; k++) {
        const size_t r0 = chunk_at[k], cr = chunk_at[k + 1] - chunk_at[k], j0 = r0 * j1;
        auto *d_dhl = up.in(m->dh_leaf + r0 * r_dhl, cr * r_dhl);
        auto *d_dha = up.in(m->dh_aunts + r0 * r_aunt, cr * r_aunt);
        auto *d_lbl = up.in(m->lb_leaf + r0 * r_lbl, cr * r_lbl);
        auto *d_lba = up.in(m->lb_aunts + r0 * r_aunt, cr * r_aunt);
        auto *d_dig = up.out<uint8_t>(cr * r_dig);
        if (up.rc) return up.rc;
        BSX_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[2 * k], s_up));
        mark(s_up);
        BSX_CUDA(ctx, cudaStreamWaitEvent(s_k, ctx->ev_chunk[2 * k], 0));
        rc = bsx_prove_data_commitment_dev(ctx, s_k, (uint32_t)cr, n_jobs, B, d_dhl, d_dha, d_lbl, d_lba, d_sh + j0 * 32, d_eh + j0 * 32,
                                           d_sb + r0, d_rsh + 32 * r0, d_eb + r0, d_reh + 32 * r0, d_dig, d_sub + r0 * r_sub,
                                           d_rd + r0 * r_rd, d_rn + r0 * r_rn, d_dc + 32 * r0, d_fail + r0);
        if (rc) return rc;
        BSX_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[2 * k + 1], s_k));
        mark(s_k);
        BSX_CUDA(ctx, cudaStreamWaitEvent(s_down, ctx->ev_chunk[2 * k + 1], 0));
        down.back(m->map_digests ? m->map_digests + r0 * r_dig : nullptr, d_dig, cr * r_dig);
        if (down.rc) return down.rc;
        mark(s_down);
    }
    down.back(m->map_subchains, d_sub, R * r_sub);
    down.back(m->reduce_digests, d_rd, R * r_rd);
    down.back(m->reduce_nodes, d_rn, R * r_rn);
    down.back(m->data_commitments, d_dc, 32 * R);
    down.back(m->fail, d_fail, R);
    if (down.rc) return down.rc;
    if (ed_last) {
        BSX_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_chunk[2 * (n_chunks - 1) + 1], 0));
        rc = run_skip();
        if (rc) return rc;
    }
    mark(ctx->stream2);
    // ---- join everything on the ctx stream ----
    BSX_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->stream2));
    BSX_CUDA(ctx, cudaStreamWaitEvent(main, ctx->ev_join, 0));
    for (int i = 0; i < BSX_PIPE_STREAMS; i++) {
        BSX_CUDA(ctx, cudaEventRecord(ctx->ev_pipe[i], ctx->pipe[i]));
        BSX_CUDA(ctx, cudaStreamWaitEvent(main, ctx->ev_pipe[i], 0));
    }
    BSX_CUDA(ctx, cudaStreamSynchronize(main));
    if (trace) {
        float t[3];
        for (uint32_t k = 0; k < n_chunks; k++) {
            for (int i = 0; i < 3; i++) cudaEventElapsedTime(&t[i], tev[0], tev[1 + 3 * k + i]);
            fprintf(stderr, "[bsx pipe] chunk %u: h2d done %.3f ms, kernels done %.3f ms, d2h done %.3f ms\n", k, t[0], t[1], t[2]);
        }
        cudaEventElapsedTime(&t[0], tev[0], tev[1 + 3 * n_chunks]);
        fprintf(stderr, "[bsx pipe] skip half (%s) done %.3f ms\n", ed_last ? "after the map kernels" : "concurrent", t[0]);
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    return BSX_OK;
}

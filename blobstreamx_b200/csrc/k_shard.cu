// Sharded header_range map/reduce across the GPUs of one node, below Python: the C ABI a Rust host binds.
// Replaces the reference's sequential `LocalProver::batch_prove` loop (PX/backend/prover/local.rs:34-57) and the HTTP
// fan-out of `RemoteProver` (PX/backend/prover/remote.rs:98-153) under `MapReduceGenerator::run_once`
// (PX/frontend/mapreduce/generator.rs:86-151): map jobs are independent (:97-111), the reduce layers follow (:113-151).
//
// Partition: rank r of W owns jobs [r J/W, (r+1) J/W) of EVERY range in flight and reduces ranges
// [r R/W, (r+1) R/W).  The only exchange is the 128-byte MapReduceSubchainVariable record of each job:
//   * the kernel that computes a record stores it straight into the gathered [R/W, J, 128] array of the rank that reduces
//     its range -- peer memory over NVLink (cudaIpc handles between processes, plain pointers inside one process);
//   * the LAST CTA of that kernel publishes the step number into a flag word on every peer (st.release.sys), and the
//     reduce kernel of each rank acquires the W flags before it reads the records (ld.acquire.sys).
// No collective kernel and no barrier kernel: nothing has to find an SM slot between the resident Ed25519 CTAs, and a
// rank waits only for records it actually needs.  Two gathered arrays alternate by step parity: a rank can only start
// the map of step k+2 after its own reduce of step k+1, which waited for every rank's map of step k+1, which each
// rank's stream ordered after its reduce of step k -- so the array of step k is free again.
//
// Exchange buffer of a rank: records[2][R/W][J][128] then flags[2][BSX_MAX_PEERS] (u32 step numbers, one per source rank).
#include "common.cuh"

#include <new>

int bsx_subchain_map_signal_dev(bsx_ctx *ctx, void *stream, uint32_t B, uint32_t n_jobs, const uint8_t *dh_leaf,
                                const uint8_t *dh_aunts, const uint8_t *lb_leaf, const uint8_t *lb_aunts,
                                const uint8_t *start_headers, const uint8_t *end_headers, const uint64_t *batch_start,
                                const uint64_t *batch_end, const uint64_t *global_end, const uint8_t *global_end_header,
                                uint8_t *digests, uint8_t *const *peer_bases, uint32_t n_peers, uint32_t rank,
                                uint32_t jobs_per_rank, uint32_t total_jobs, uint32_t ranges_per_owner, uint32_t *sig_done,
                                uint32_t *const *sig_flags, uint32_t sig_step);
int bsx_reduce_subchains_wait_dev(bsx_ctx *ctx, void *stream, uint32_t n_ranges, uint32_t n_jobs, const uint8_t *map_subchains,
                                  const uint64_t *start_blocks, const uint8_t *start_header, const uint64_t *end_blocks,
                                  const uint8_t *end_header, uint32_t B, uint8_t *reduce_digests, uint8_t *reduce_nodes,
                                  uint8_t *data_commitments, uint32_t *fail, const uint32_t *wait_flags, uint32_t n_wait,
                                  uint32_t wait_step);

struct bsx_shard {
    bsx_ctx *ctx;
    uint32_t rank, world, n_ranges, n_jobs, B, per, own;   // per = jobs per range per rank, own = ranges this rank reduces
    size_t records_bytes;            // one gathered array
    size_t total_bytes;
    uint8_t *buf;                    // this rank's exchange buffer
    bool owns_buf;
    uint8_t *peer[BSX_MAX_PEERS];    // every rank's exchange buffer as mapped here (peer[rank] = buf)
    bool peer_ipc[BSX_MAX_PEERS];    // opened with cudaIpcOpenMemHandle (to be closed)
    uint32_t *done;                  // CTA counter of the map kernel (device)
    uint32_t step;                   // steps issued so far
};

static size_t flags_offset(const bsx_shard *sh) { return 2 * sh->records_bytes; }

extern "C" size_t bsx_shard_exchange_bytes(uint32_t world, uint32_t n_ranges, uint32_t n_jobs) {
    if (world == 0 || n_ranges % world) return 0;
    const size_t rec = (size_t)(n_ranges / world) * n_jobs * BSX_SUBCHAIN_BYTES;
    return 2 * rec + 2 * BSX_MAX_PEERS * sizeof(uint32_t);
}

extern "C" int bsx_shard_create(bsx_ctx *ctx, uint32_t rank, uint32_t world, uint32_t n_ranges, uint32_t n_jobs, uint32_t B,
                                void *exchange_buf, bsx_shard **out) {
    BSX_REQUIRE(ctx, ctx && out);
    *out = nullptr;
    BSX_REQUIRE(ctx, world >= 1 && world <= BSX_MAX_PEERS && rank < world);
    BSX_REQUIRE(ctx, n_jobs >= 1 && (n_jobs & (n_jobs - 1)) == 0 && n_jobs % world == 0 && n_ranges >= world && n_ranges % world == 0);
    BSX_REQUIRE(ctx, B >= 1 && B <= 256 && (B & (B - 1)) == 0);
    BSX_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(exchange_buf) & 127) == 0);
    BSX_CUDA(ctx, cudaSetDevice(ctx->device));
    bsx_shard *sh = new (std::nothrow) bsx_shard();
    if (!sh) return bsx::fail(ctx, BSX_ERR_NOMEM, "out of host memory%s%s");
    memset(sh, 0, sizeof *sh);
    sh->ctx = ctx; sh->rank = rank; sh->world = world; sh->n_ranges = n_ranges; sh->n_jobs = n_jobs; sh->B = B;
    sh->per = n_jobs / world; sh->own = n_ranges / world;
    sh->records_bytes = (size_t)sh->own * n_jobs * BSX_SUBCHAIN_BYTES;
    sh->total_bytes = bsx_shard_exchange_bytes(world, n_ranges, n_jobs);
    if (exchange_buf) {
        sh->buf = reinterpret_cast<uint8_t *>(exchange_buf);     // e.g. symmetric memory the caller set up
    } else {
        // cudaMalloc (not the stream-ordered pool): the allocation must be exportable with cudaIpcGetMemHandle
        if (cudaMalloc(&sh->buf, sh->total_bytes) != cudaSuccess) {
            delete sh;
            return bsx::fail(ctx, BSX_ERR_NOMEM, "cudaMalloc (exchange buffer)%s%s");
        }
        sh->owns_buf = true;
    }
    if (cudaMemset(sh->buf, 0, sh->total_bytes) != cudaSuccess || cudaMalloc(&sh->done, sizeof(uint32_t)) != cudaSuccess ||
        cudaMemset(sh->done, 0, sizeof(uint32_t)) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        if (sh->owns_buf) cudaFree(sh->buf);
        if (sh->done) cudaFree(sh->done);
        delete sh;
        return bsx::fail(ctx, BSX_ERR_CUDA, "exchange buffer setup failed%s%s");
    }
    sh->peer[rank] = sh->buf;
    *out = sh;
    return BSX_OK;
}

extern "C" void bsx_shard_destroy(bsx_shard *sh) {
    if (!sh) return;
    cudaSetDevice(sh->ctx->device);
    cudaDeviceSynchronize();
    for (uint32_t w = 0; w < sh->world; w++)
        if (sh->peer_ipc[w] && sh->peer[w]) cudaIpcCloseMemHandle(sh->peer[w]);
    if (sh->owns_buf && sh->buf) cudaFree(sh->buf);
    if (sh->done) cudaFree(sh->done);
    delete sh;
}

extern "C" int bsx_shard_exchange_buffer(bsx_shard *sh, void **dev_ptr, size_t *bytes) {
    if (!sh) return BSX_ERR_INVALID;
    if (dev_ptr) *dev_ptr = sh->buf;
    if (bytes) *bytes = sh->total_bytes;
    return BSX_OK;
}

extern "C" int bsx_shard_ipc_handle(bsx_shard *sh, uint8_t *handle) {
    if (!sh || !handle) return BSX_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == BSX_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    BSX_CUDA(sh->ctx, cudaSetDevice(sh->ctx->device));
    cudaIpcMemHandle_t h;
    BSX_CUDA(sh->ctx, cudaIpcGetMemHandle(&h, sh->buf));
    memcpy(handle, &h, sizeof h);
    return BSX_OK;
}

extern "C" int bsx_shard_open_peer(bsx_shard *sh, uint32_t peer, const uint8_t *handle) {
    if (!sh || !handle) return BSX_ERR_INVALID;
    BSX_REQUIRE(sh->ctx, peer < sh->world && peer != sh->rank && !sh->peer[peer]);
    BSX_CUDA(sh->ctx, cudaSetDevice(sh->ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void *p = nullptr;
    BSX_CUDA(sh->ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    sh->peer[peer] = reinterpret_cast<uint8_t *>(p);
    sh->peer_ipc[peer] = true;
    return BSX_OK;
}

extern "C" int bsx_shard_set_peer(bsx_shard *sh, uint32_t peer, void *dev_ptr) {
    if (!sh) return BSX_ERR_INVALID;
    BSX_REQUIRE(sh->ctx, peer < sh->world && peer != sh->rank && dev_ptr && (reinterpret_cast<uintptr_t>(dev_ptr) & 127) == 0);
    BSX_CUDA(sh->ctx, cudaSetDevice(sh->ctx->device));
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, dev_ptr) == cudaSuccess && at.type == cudaMemoryTypeDevice && at.device != sh->ctx->device) {
        const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);      // same process, another GPU
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            return bsx::fail(sh->ctx, BSX_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s%s", cudaGetErrorString(e));
        cudaGetLastError();
    }
    sh->peer[peer] = reinterpret_cast<uint8_t *>(dev_ptr);
    sh->peer_ipc[peer] = false;
    return BSX_OK;
}

extern "C" int bsx_shard_step_dev(bsx_shard *sh, void *stream, const bsx_shard_in *in, const bsx_shard_out *out) {
    if (!sh) return BSX_ERR_INVALID;
    bsx_ctx *ctx = sh->ctx;
    BSX_REQUIRE(ctx, in && out && out->map_digests && out->data_commitments && out->fail);
    for (uint32_t w = 0; w < sh->world; w++)
        if (!sh->peer[w]) return bsx::fail(ctx, BSX_ERR_INVALID, "bsx_shard_step_dev: peer buffer of a rank is not set%s%s");
    const uint32_t step = ++sh->step, par = step & 1;
    uint8_t *bases[BSX_MAX_PEERS];
    uint32_t *flags[BSX_MAX_PEERS];
    for (uint32_t w = 0; w < sh->world; w++) {
        bases[w] = sh->peer[w] + par * sh->records_bytes;
        flags[w] = reinterpret_cast<uint32_t *>(sh->peer[w] + flags_offset(sh)) + par * BSX_MAX_PEERS + sh->rank;
    }
    int rc = bsx_subchain_map_signal_dev(ctx, stream, sh->B, sh->n_ranges * sh->per, in->dh_leaf, in->dh_aunts, in->lb_leaf,
                                         in->lb_aunts, in->start_headers, in->end_headers, in->batch_start, in->batch_end,
                                         in->global_end, in->global_end_header, out->map_digests, bases, sh->world, sh->rank,
                                         sh->per, sh->n_jobs, sh->own, sh->done, flags, step);
    if (rc) return rc;
    const uint8_t *mine = sh->buf + par * sh->records_bytes;
    const uint32_t *wait = reinterpret_cast<const uint32_t *>(sh->buf + flags_offset(sh)) + par * BSX_MAX_PEERS;
    rc = bsx_reduce_subchains_wait_dev(ctx, stream, sh->own, sh->n_jobs, mine, in->start_blocks, in->start_header, in->end_blocks,
                                       in->end_header, sh->B, out->reduce_digests, out->reduce_nodes, out->data_commitments,
                                       out->fail, wait, sh->world, step);
    if (rc) return rc;
    if (out->map_subchains)   // optional copy of the gathered records of this rank's ranges (stream-ordered after the reduce)
        BSX_CUDA(ctx, cudaMemcpyAsync(out->map_subchains, mine, sh->records_bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return BSX_OK;
}

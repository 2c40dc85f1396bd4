/*
 * bsx.h -- C ABI of libbsx, the B200 (sm_100a) witness-generation library for the Blobstream X
 * header_range / next_header / data_commitment circuits.
 *
 * This is the drop-in boundary: a plonky2x `Hint` on the Rust side packs the queued
 * requests of one accelerator and calls ONE of these entry points instead of looping over
 * per-request hints.  Each entry point cites the reference interface it replaces
 * (PX = contracts/lib/succinctx/plonky2x/core/src, TX = contracts/lib/tendermintx/circuits,
 * BX = blobstreamx root; all under /root/reference).  INTEGRATION.md shows the Rust binding.
 *
 * Conventions
 *   - every function returns 0 (BSX_OK) or a negative bsx_status; bsx_last_error(ctx) gives text.
 *   - plain pointers + sizes only; the caller owns every buffer.
 *   - functions without a suffix take HOST pointers and are synchronous (H2D, kernels, D2H).
 *   - `_dev` twins take DEVICE pointers plus a cudaStream_t (passed as void*) and only enqueue
 *     work; no allocation, no sync.  They are what the host entry points call internally.
 *   - one bsx_ctx per GPU; a ctx is NOT thread-safe (the reference's witness loop is single
 *     threaded: PX/backend/circuit/witness.rs:144-199); use one ctx per calling thread.
 *   - digests are raw big-endian SHA bytes (what `Bytes32Variable` holds); curve field elements
 *     are 32 little-endian canonical bytes (= 16 u16 limbs of `FieldVariable`,
 *     PX/frontend/curta/field/variable.rs:38-103); integers are little-endian.
 *   - circuit assertions never abort: they are reported as bit masks (BSX_FAIL_* / BSX_VFAIL_*),
 *     0 meaning the reference circuit would have accepted the witness.
 */
#ifndef BSX_H
#define BSX_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSX_VERSION 100

typedef enum {
    BSX_OK = 0,
    BSX_ERR_INVALID = -1, /* bad argument */
    BSX_ERR_CUDA = -2,    /* CUDA runtime failure, see bsx_last_error */
    BSX_ERR_NOMEM = -3,
    BSX_ERR_NODEVICE = -4
} bsx_status;

typedef struct bsx_ctx bsx_ctx;

int bsx_init(int device, bsx_ctx **out);
void bsx_destroy(bsx_ctx *ctx);
const char *bsx_last_error(const bsx_ctx *ctx);
int bsx_version(void);
/* number of kernels this ctx has launched since init (bench.py's gpu_launches) */
uint64_t bsx_launch_count(const bsx_ctx *ctx);
/* block until everything enqueued on the ctx's own stream is done */
int bsx_sync(bsx_ctx *ctx);
/* Measurement knobs of one ctx (kernel-build and stream-arrangement choices; defaults = the measured best; the
 * environment variable BSX_<name> sets the default a new ctx starts with).  Names: ED_MODE, ED_QUAD_MAX, ED_INLINE,
 * ED_OCC, ED_REGS, ED_FP64, ED_KEYTAB, ED_PAIR, ED_RESIDENT, HR_HASH_STREAM, HR_TRACE, PIPE_CHUNK, PIPE_ED, PIPE_TRACE,
 * PROOFS_OCC, SUBCHAIN_FUSED, COMMIT_THREADS, ED_TRACE_LANES
 * (DESIGN.md section 4).  No reference counterpart: results are identical under every setting. */
int bsx_set_tunable(bsx_ctx *ctx, const char *name, int value);
int bsx_get_tunable(const bsx_ctx *ctx, const char *name, int *value);

/* ------------------------------------------------------------------------------------------
 * K1  batched SHA-256 / SHA-512 digests
 * replaces: one HashDigestHint::hint per request (PX/frontend/hash/curta/digest_hint.rs:30-38,
 *           emitted by PX/frontend/hash/curta/builder.rs:26-49); fixed requests pass the whole
 *           message, variable requests pass the first `len` bytes (digest_hint.rs:31-33).
 * msgs: concatenated messages; offsets[n+1]; digests: n*32 (n*64) raw bytes.
 * ------------------------------------------------------------------------------------------ */
int bsx_sha256_batch(bsx_ctx *ctx, const uint8_t *msgs, const uint32_t *offsets, uint32_t n, uint8_t *digests);
int bsx_sha512_batch(bsx_ctx *ctx, const uint8_t *msgs, const uint32_t *offsets, uint32_t n, uint8_t *digests);
int bsx_sha256_batch_dev(bsx_ctx *ctx, void *stream, const uint8_t *msgs, const uint32_t *offsets, uint32_t n,
                         uint8_t *digests);
int bsx_sha512_batch_dev(bsx_ctx *ctx, void *stream, const uint8_t *msgs, const uint32_t *offsets, uint32_t n,
                         uint8_t *digests);

/* ------------------------------------------------------------------------------------------
 * K3  Tendermint Merkle inclusion proofs, fixed-shape evaluation
 * replaces: get_root_from_merkle_proof / _hashed_leaf (PX/frontend/merkle/tendermint.rs:62-93):
 *           leaf hash, then per level BOTH inner(h,aunt) and inner(aunt,h); path bit selects.
 * leaves: n*leaf_len (or n*32 leaf digests when hashed_leaf); aunts: n*depth*32;
 * path_bits[n]: bit i = path_indices[i]; digests: n*(2*depth + !hashed_leaf)*32 in request
 * order; roots: n*32.  depth <= 32.
 * ------------------------------------------------------------------------------------------ */
int bsx_tm_merkle_proofs(bsx_ctx *ctx, const uint8_t *leaves, uint32_t leaf_len, const uint8_t *aunts, uint32_t depth,
                         const uint32_t *path_bits, uint32_t n, int hashed_leaf, uint8_t *digests, uint8_t *roots);
int bsx_tm_merkle_proofs_dev(bsx_ctx *ctx, void *stream, const uint8_t *leaves, uint32_t leaf_len,
                             const uint8_t *aunts, uint32_t depth, const uint32_t *path_bits, uint32_t n,
                             int hashed_leaf, uint8_t *digests, uint8_t *roots);

/* ------------------------------------------------------------------------------------------
 * K2  Tendermint Merkle tree over hashed leaves, fixed-shape evaluation
 * replaces: get_root_from_hashed_leaves<N> + hash_merkle_layer (tendermint.rs:124-204).
 * t trees of N leaf digests each (padded internally to P = 2^ceil(log2 N) with zero digests);
 * nb_enabled[t] (field element; >= P means all enabled); inner: t*(P-1)*32 raw inner hashes,
 * layer-major (EVERY pair is hashed); roots: t*32 after select.  N <= 4096.
 * ------------------------------------------------------------------------------------------ */
int bsx_tm_merkle_tree(bsx_ctx *ctx, const uint8_t *leaf_digests, uint32_t N, uint32_t t, const uint64_t *nb_enabled,
                       uint8_t *inner, uint8_t *roots);
int bsx_tm_merkle_tree_dev(bsx_ctx *ctx, void *stream, const uint8_t *leaf_digests, uint32_t N, uint32_t t,
                           const uint64_t *nb_enabled, uint8_t *inner, uint8_t *roots);

/* ------------------------------------------------------------------------------------------
 * get_data_commitment<N>  (BX/circuits/builder.rs:105-148 with encode_data_root_tuple :82-103)
 * t independent trees; data_hashes: t*N*32; start/end: t u64 each;
 * digests: t*(N + P-1)*32 = N leaf digests then P-1 inner (layer-major); roots: t*32;
 * fail[t]: BSX_FAIL_END_LT_START when end < start or end-start >= 2^32.
 * ------------------------------------------------------------------------------------------ */
int bsx_data_commitment_batch(bsx_ctx *ctx, const uint8_t *data_hashes, uint32_t N, uint32_t t,
                              const uint64_t *start_blocks, const uint64_t *end_blocks, uint8_t *digests,
                              uint8_t *roots, uint32_t *fail);
int bsx_data_commitment_batch_dev(bsx_ctx *ctx, void *stream, const uint8_t *data_hashes, uint32_t N, uint32_t t,
                                  const uint64_t *start_blocks, const uint64_t *end_blocks, uint8_t *digests,
                                  uint8_t *roots, uint32_t *fail);

/* ------------------------------------------------------------------------------------------
 * Attestation proofs: Merkle inclusion proofs of data-root tuples in a data commitment, as
 * BlobstreamX.verifyAttestation consumes them (BX/contracts/src/BlobstreamX.sol: BinaryMerkleProof{sideNodes, key,
 * numLeaves} over DataRootTuple{height, dataRoot}); sideNodes = the aunts of compute_hash_from_aunts
 * (TX/input/tendermint_utils.rs:225-273), leaf side first.
 * Query q asks for block q_height[q] of tree q_tree[q].  side_nodes: n_q * bsx_attestation_max_depth(N) * 32, the
 * first depth[q] entries are the proof, the rest zero; key[q] = height - start (0xFFFFFFFF if the height is outside
 * [start, end)); num_leaves[q] = end - start.  The _dev form gathers from the `digests` array of
 * bsx_data_commitment_batch_dev; the host form computes the commitments first (roots optional).
 * ------------------------------------------------------------------------------------------ */
uint32_t bsx_attestation_max_depth(uint32_t N);
int bsx_attestation_proofs(bsx_ctx *ctx, const uint8_t *data_hashes, uint32_t N, uint32_t t, const uint64_t *start_blocks,
                           const uint64_t *end_blocks, uint32_t n_q, const uint32_t *q_tree, const uint64_t *q_height,
                           uint8_t *side_nodes, uint32_t *depth, uint32_t *key, uint32_t *num_leaves, uint8_t *roots);
int bsx_attestation_proofs_dev(bsx_ctx *ctx, void *stream, const uint8_t *digests, uint32_t N, const uint64_t *start_blocks,
                               const uint64_t *end_blocks, uint32_t n_q, const uint32_t *q_tree, const uint64_t *q_height,
                               uint8_t *side_nodes, uint32_t *depth, uint32_t *key, uint32_t *num_leaves);

/* ------------------------------------------------------------------------------------------
 * prove_subchain<B>  -- the map circuit of header_range (BX/circuits/builder.rs:150-271)
 * n_jobs independent map jobs of B headers each (B in {1,2,4,...,256}).
 *   dh_leaf  n_jobs*B*34   protobuf data_hash leaves        dh_aunts n_jobs*B*4*32
 *   lb_leaf  n_jobs*B*72   protobuf last_block_id leaves    lb_aunts n_jobs*B*4*32
 *   start_headers/end_headers n_jobs*32 (DataCommitmentProofVariable.{start,end}_header)
 *   batch_start/batch_end/global_end: n_jobs u64;  global_end_header: n_jobs*32
 * outputs
 *   digests   n_jobs*(20B-1)*32 in Curta request order: per header 9 data_hash-proof
 *             digests then 9 last_block_id-proof digests; then B tuple-leaf digests; then B-1 inner.
 *   subchains n_jobs*BSX_SUBCHAIN_BYTES  MapReduceSubchainVariable (BX/circuits/vars.rs:29-36):
 *             [0] is_enabled, [4..8) fail mask, [8..16) start_block, [16..24) end_block,
 *             [24..56) start_header, [56..88) end_header, [88..120) data_merkle_root.
 * ------------------------------------------------------------------------------------------ */
#define BSX_SUBCHAIN_BYTES 128
#define BSX_FAIL_PREV_HEADER 1u       /* builder.rs:204-207 */
#define BSX_FAIL_DATA_HASH 2u         /* :209-212 */
#define BSX_FAIL_END_HEADER 4u        /* :214-219 */
#define BSX_FAIL_BATCH_END_HEADER 8u  /* :228-232 */
#define BSX_FAIL_END_LT_START 16u     /* :111-128 */
#define BSX_FAIL_REDUCE_LINK 32u      /* :349-355 */
#define BSX_FAIL_RANGE 64u            /* :291-297 */
#define BSX_FAIL_RESULT 128u          /* :398-406 */
#define BSX_FAIL_EXCHANGE_TIMEOUT 2048u /* sharded engine: a peer rank's subchain records did not arrive within 4 s */
#define BSX_FAIL_INPUT_LEAF 256u      /* input shaping: a proven header field does not have the circuit's fixed size */
int bsx_prove_subchain_batch(bsx_ctx *ctx, uint32_t B, uint32_t n_jobs, const uint8_t *dh_leaf,
                             const uint8_t *dh_aunts, const uint8_t *lb_leaf, const uint8_t *lb_aunts,
                             const uint8_t *start_headers, const uint8_t *end_headers, const uint64_t *batch_start,
                             const uint64_t *batch_end, const uint64_t *global_end, const uint8_t *global_end_header,
                             uint8_t *digests, uint8_t *subchains);
int bsx_prove_subchain_batch_dev(bsx_ctx *ctx, void *stream, uint32_t B, uint32_t n_jobs, const uint8_t *dh_leaf,
                                 const uint8_t *dh_aunts, const uint8_t *lb_leaf, const uint8_t *lb_aunts,
                                 const uint8_t *start_headers, const uint8_t *end_headers,
                                 const uint64_t *batch_start, const uint64_t *batch_end, const uint64_t *global_end,
                                 const uint8_t *global_end_header, uint8_t *digests, uint8_t *subchains);

/* ------------------------------------------------------------------------------------------
 * prove_data_commitment<NB_MAP_JOBS,B>  -- map + reduce of n_ranges independent header ranges
 * (BX/circuits/builder.rs:273-409; PX/frontend/mapreduce/generator.rs:86-151).  Job j of range r
 * covers [start+jB, start+(j+1)B); arrays are range-major then job-major.
 *   start_blocks/end_blocks: n_ranges u64; start_header/end_header: n_ranges*32 (public inputs)
 * outputs (per range): map_digests n_jobs*(20B-1)*32, map_subchains n_jobs*128,
 *   reduce_digests (n_jobs-1)*32 layer-major (the bit-level sha256 of builder.rs:363-364),
 *   reduce_nodes (n_jobs-1)*128, data_commitments 32, fail u32.  n_jobs must be a power of two.
 * ------------------------------------------------------------------------------------------ */
int bsx_prove_data_commitment(bsx_ctx *ctx, uint32_t n_ranges, uint32_t n_jobs, uint32_t B, const uint8_t *dh_leaf,
                              const uint8_t *dh_aunts, const uint8_t *lb_leaf, const uint8_t *lb_aunts,
                              const uint8_t *start_headers, const uint8_t *end_headers, const uint64_t *start_blocks,
                              const uint8_t *start_header, const uint64_t *end_blocks, const uint8_t *end_header,
                              uint8_t *map_digests, uint8_t *map_subchains, uint8_t *reduce_digests,
                              uint8_t *reduce_nodes, uint8_t *data_commitments, uint32_t *fail);
int bsx_prove_data_commitment_dev(bsx_ctx *ctx, void *stream, uint32_t n_ranges, uint32_t n_jobs, uint32_t B,
                                  const uint8_t *dh_leaf, const uint8_t *dh_aunts, const uint8_t *lb_leaf,
                                  const uint8_t *lb_aunts, const uint8_t *start_headers, const uint8_t *end_headers,
                                  const uint64_t *start_blocks, const uint8_t *start_header,
                                  const uint64_t *end_blocks, const uint8_t *end_header, uint8_t *map_digests,
                                  uint8_t *map_subchains, uint8_t *reduce_digests, uint8_t *reduce_nodes,
                                  uint8_t *data_commitments, uint32_t *fail);
/* reduce stage alone over already-computed map outputs (used after the multi-GPU all-gather):
 * map_subchains n_ranges*n_jobs*128 -> reduce_* / data_commitments / fail as above. */
/* Multi-GPU map stage: as bsx_prove_subchain_batch_dev, but each 128-byte subchain record is stored through
 * peer memory (NVLink) straight into the gathered [ranges_per_owner, total_jobs, 128] array of the rank that reduces its
 * range, so the exchange step of the map/reduce needs no collective kernel -- only a cross-rank barrier before
 * bsx_reduce_subchains_dev.  This rank holds `jobs_per_rank` jobs (global jobs rank*jobs_per_rank ...) of every range
 * in flight, n_jobs = jobs_per_rank * ranges_per_owner * n_peers, range r is reduced by rank r / ranges_per_owner.
 * peer_bases: host array of n_peers device addresses (rank w's array as mapped in this process, e.g. the buffer_ptrs of
 * torch symmetric memory or cudaIpcOpenMemHandle results); n_peers <= BSX_MAX_PEERS. */
#define BSX_MAX_PEERS 16
int bsx_prove_subchain_batch_p2p_dev(bsx_ctx *ctx, void *stream, uint32_t B, uint32_t n_jobs, const uint8_t *dh_leaf,
                                     const uint8_t *dh_aunts, const uint8_t *lb_leaf, const uint8_t *lb_aunts,
                                     const uint8_t *start_headers, const uint8_t *end_headers, const uint64_t *batch_start,
                                     const uint64_t *batch_end, const uint64_t *global_end, const uint8_t *global_end_header,
                                     uint8_t *digests, const uint64_t *peer_bases, uint32_t n_peers, uint32_t rank,
                                     uint32_t jobs_per_rank, uint32_t total_jobs, uint32_t ranges_per_owner);
int bsx_reduce_subchains_dev(bsx_ctx *ctx, void *stream, uint32_t n_ranges, uint32_t n_jobs,
                             const uint8_t *map_subchains, const uint64_t *start_blocks, const uint8_t *start_header,
                             const uint64_t *end_blocks, const uint8_t *end_header, uint32_t B, uint8_t *reduce_digests,
                             uint8_t *reduce_nodes, uint8_t *data_commitments, uint32_t *fail);

/* ------------------------------------------------------------------------------------------
 * Sharded map/reduce across the GPUs of one node
 * replaces: the sequential loop of LocalProver::batch_prove (PX/backend/prover/local.rs:34-57) / the HTTP fan-out of
 *           RemoteProver (PX/backend/prover/remote.rs:98-153) under MapReduceGenerator::run_once
 *           (PX/frontend/mapreduce/generator.rs:86-151; map jobs independent :97-111, reduce layers :113-151).
 * One bsx_shard per rank (= per GPU, on that rank's ctx).  `n_ranges` ranges are in flight per step over all ranks; rank r
 * of `world` owns jobs [r*n_jobs/world, (r+1)*n_jobs/world) of EVERY range and reduces ranges [r*n_ranges/world, ...).
 * The subchain records travel by peer-memory stores from the kernel that computes them, completion by one flag word per
 * (source rank, step parity) published by that kernel's last CTA and acquired by the reduce kernel: no collective and no
 * barrier launch on the data path.
 * Setup: create on every rank; exchange either bsx_shard_ipc_handle -> bsx_shard_open_peer (ranks in different processes)
 * or bsx_shard_exchange_buffer -> bsx_shard_set_peer (same process, or buffers the caller mapped itself, e.g. symmetric
 * memory passed as `exchange_buf`, whose size is bsx_shard_exchange_bytes); then every rank calls bsx_shard_step_dev once
 * per step, each on ONE stream of its own (steps of a rank must be stream-ordered).
 * fail[r] gains BSX_FAIL_EXCHANGE_TIMEOUT if a rank's records have not arrived after 4 s (never a hang).
 * ------------------------------------------------------------------------------------------ */
typedef struct bsx_shard bsx_shard;
#define BSX_IPC_HANDLE_BYTES 64
typedef struct {              /* device pointers */
    /* this rank's job slice of every range in flight: n_ranges * n_jobs/world jobs, range-major, then local job */
    const uint8_t *dh_leaf, *dh_aunts, *lb_leaf, *lb_aunts; /* per job: B*34, B*128, B*72, B*128 */
    const uint8_t *start_headers, *end_headers;             /* per job: 32 */
    const uint64_t *batch_start, *batch_end, *global_end;   /* per job */
    const uint8_t *global_end_header;                       /* per job: 32 */
    /* public inputs of the n_ranges/world ranges this rank reduces */
    const uint64_t *start_blocks, *end_blocks;
    const uint8_t *start_header, *end_header;               /* 32 each */
} bsx_shard_in;
typedef struct {              /* device pointers */
    uint8_t *map_digests;      /* n_ranges * n_jobs/world * (20B-1) * 32: the digests of this rank's jobs */
    uint8_t *map_subchains;    /* optional (may be NULL): n_ranges/world * n_jobs * 128, the gathered records of this rank's ranges */
    uint8_t *reduce_digests;   /* optional: n_ranges/world * (n_jobs-1) * 32 */
    uint8_t *reduce_nodes;     /* optional: n_ranges/world * (n_jobs-1) * 128 */
    uint8_t *data_commitments; /* n_ranges/world * 32 */
    uint32_t *fail;            /* n_ranges/world */
} bsx_shard_out;
size_t bsx_shard_exchange_bytes(uint32_t world, uint32_t n_ranges, uint32_t n_jobs);
int bsx_shard_create(bsx_ctx *ctx, uint32_t rank, uint32_t world, uint32_t n_ranges, uint32_t n_jobs, uint32_t B,
                     void *exchange_buf /* NULL: allocated here (cudaMalloc, IPC-exportable) */, bsx_shard **out);
void bsx_shard_destroy(bsx_shard *sh);
int bsx_shard_exchange_buffer(bsx_shard *sh, void **dev_ptr, size_t *bytes);
int bsx_shard_ipc_handle(bsx_shard *sh, uint8_t *handle /* BSX_IPC_HANDLE_BYTES */);
int bsx_shard_open_peer(bsx_shard *sh, uint32_t peer, const uint8_t *handle /* BSX_IPC_HANDLE_BYTES */);
int bsx_shard_set_peer(bsx_shard *sh, uint32_t peer, void *dev_ptr);
int bsx_shard_step_dev(bsx_shard *sh, void *stream, const bsx_shard_in *in, const bsx_shard_out *out);

/* ------------------------------------------------------------------------------------------
 * K4+K5  batched Ed25519 witness generation
 * replaces, per signature of curta_eddsa_verify_sigs[_conditional]
 *   (PX/frontend/ecc/curve25519/ed25519/eddsa.rs:72-127,131-204):
 *   HashDigestHint<SHA512>            PX/frontend/hash/sha/sha512/curta.rs:103-111
 *   BigUintDivRemGenerator            PX/frontend/uint/num/biguint/mod.rs:451-488
 *   7 x EcOpResultHint::hint          PX/frontend/ecc/curve25519/curta/result_hint.rs:21-50
 *     order: ScalarMul(s,G), Decompress(pk), IsValid, ScalarMul(h,A), Decompress(R), IsValid, Add
 * pks n*32 (compressed), sigs n*64 (R ‖ s little-endian), msgs n*msg_stride (MAX_MSG_LENGTH_BYTES),
 * msg_lens[n] (NULL = every message is msg_stride long), active[n] (NULL = all; 0 = run the lane on
 * DUMMY_PUBLIC_KEY / DUMMY_SIGNATURE / 32 zero bytes, eddsa.rs:28-42,102-115).
 * out: n * BSX_SIG_OUT_BYTES records (device pointers of the _dev forms: 8-byte aligned), all values little-endian canonical (x,y = 16 u16 limbs each):
 *   [0..64) sha512 digest   [64..96) h = LE512(digest) mod l   [96..136) div = LE512(digest) / l
 *   [136..200) s*G (x,y)    [200..264) A (x,y)   [264..296) A_root   [296..360) h*A
 *   [360..424) Rp (x,y)     [424..456) R_root    [456..520) Rp + h*A   [520..524) flags
 * flags: BSX_SIG_S_LT_L | BSX_SIG_A_OK | BSX_SIG_R_OK | BSX_SIG_EQ ; 0xF = the circuit accepts.
 * A point that fails to decompress (the reference panics) is reported as the identity with root 0.
 * Batches above 16 384 signatures (one thread per signature): public keys that repeat within the batch
 * (one validator set signing many ranges) are found on the device and h*A is taken from per-key window
 * tables, R is checked against s*G - h*A before it is decompressed; the records are the same bytes on
 * every path.  Such a call draws up to ~250 MB of stream-ordered pool memory (cudaMallocAsync) while it runs.
 * ------------------------------------------------------------------------------------------ */
#define BSX_SIG_OUT_BYTES 576
#define BSX_SIG_S_LT_L 1u
#define BSX_SIG_A_OK 2u
#define BSX_SIG_R_OK 4u
#define BSX_SIG_EQ 8u
int bsx_ed25519_batch(bsx_ctx *ctx, uint32_t n, const uint8_t *pks, const uint8_t *sigs, const uint8_t *msgs,
                      uint32_t msg_stride, const uint32_t *msg_lens, const uint8_t *active, uint8_t *out);
int bsx_ed25519_batch_dev(bsx_ctx *ctx, void *stream, uint32_t n, const uint8_t *pks, const uint8_t *sigs,
                          const uint8_t *msgs, uint32_t msg_stride, const uint32_t *msg_lens, const uint8_t *active,
                          uint8_t *out);

/* ------------------------------------------------------------------------------------------
 * verify_header / verify_skip / verify_step (+ prove_next_header_data_commitment)
 * replaces every Curta SHA-256 / SHA-512 / EC hint of
 *   TendermintVerify::verify_header   TX/builder/verify.rs:225-329   (n_digests = N + P-1 + 27)
 *   TendermintVerify::verify_skip     TX/builder/verify.rs:527-564   (9 + N + P-1 + the above)
 *   TendermintVerify::verify_step     TX/builder/verify.rs:468-505 followed by
 *   prove_next_header_data_commitment BX/circuits/builder.rs:411-443 (the above + 28)
 * for n independent instances with VALIDATOR_SET_SIZE_MAX = N (P = next power of two).
 * validators: n*N records of BSX_VAL_IN_BYTES (ValidatorVariable, TX/variables.rs:72-83):
 *   [0..32) pubkey  [32..96) signature R‖s  [96..220) message  [220..224) message_byte_length LE
 *   [224..232) voting_power LE  [232..236) validator_byte_length LE  [236] signed  [237] present_on_trusted_header
 * digests: SHA-256 digests in Curta request order;
 * ed_out: n*N Ed25519 records (BSX_SIG_OUT_BYTES); fail[n]: BSX_VFAIL_* mask, 0 = circuit accepts.
 * ------------------------------------------------------------------------------------------ */
#define BSX_VAL_IN_BYTES 240
typedef struct bsx_header_in {        /* one header to verify; 616 bytes, 8-byte aligned */
    uint8_t header[32];               /* expected header hash */
    uint64_t height, round;           /* block height; commit round */
    uint64_t nb_enabled;              /* nb_enabled_validators */
    uint8_t chain_id_enc[64];         /* ChainIdProofVariable.enc_chain_id_bytes (52 used), zero padded */
    uint32_t chain_id_enc_len;        /* enc_chain_id_byte_length */
    uint32_t height_enc_len;          /* HeightProofVariable.enc_height_byte_length */
    uint8_t chain_id_aunts[128];
    uint8_t height_aunts[128];
    uint8_t validators_hash_proof[168]; /* leaf[34] ‖ aunts[128] ‖ pad */
    uint8_t expected_chain_id[56];    /* circuit constant CHAIN_ID_BYTES */
    uint32_t expected_chain_id_len;
    uint32_t _pad;
} bsx_header_in;
typedef struct bsx_skip_in {          /* VerifySkipVariable extras (TX/variables.rs:98-111) */
    uint64_t trusted_block;
    uint64_t trusted_nb_enabled;
    uint32_t skip_max;
    uint32_t _pad;
    uint8_t trusted_header[32];
    uint8_t trusted_validators_hash_proof[168];
} bsx_skip_in;
typedef struct bsx_step_in {          /* VerifyStepVariable extras + DataCommitmentProofVariable<1> */
    uint64_t prev_block;
    uint8_t prev_header[32];
    uint8_t last_block_id_proof[200];          /* leaf[72] ‖ aunts[128], against the next header */
    uint8_t prev_next_validators_proof[168];   /* leaf[34] ‖ aunts[128] ‖ pad, against prev header */
    uint8_t data_hash_proof[168];              /* leaf[34] ‖ aunts[128] ‖ pad, against prev header */
} bsx_step_in;
#define BSX_VFAIL_SIG 1u
#define BSX_VFAIL_VALHASH 2u
#define BSX_VFAIL_VALHASH_PROOF 4u
#define BSX_VFAIL_THRESHOLD 8u
#define BSX_VFAIL_MESSAGE 16u
#define BSX_VFAIL_CHAIN_ID 32u
#define BSX_VFAIL_HEIGHT 64u
#define BSX_VFAIL_TRUSTED_PROOF 128u
#define BSX_VFAIL_TRUSTED_VALHASH 256u
#define BSX_VFAIL_TRUSTED_PRESENT 512u
#define BSX_VFAIL_TRUSTED_THRESHOLD 1024u
#define BSX_VFAIL_SKIP_DISTANCE 2048u
#define BSX_VFAIL_PREV_HEADER 4096u
#define BSX_VFAIL_NEXT_VALS 8192u
#define BSX_VFAIL_DATA_HASH_PROOF 16384u
/* mode 0 verify_header, 1 verify_skip, 2 next_header -> SHA-256 digests per instance */
uint32_t bsx_verify_digest_count(int mode, uint32_t N);
int bsx_verify_header(bsx_ctx *ctx, uint32_t n, uint32_t N, const bsx_header_in *hdr, const uint8_t *validators,
                      uint8_t *digests, uint8_t *ed_out, uint32_t *fail);
int bsx_verify_skip(bsx_ctx *ctx, uint32_t n, uint32_t N, const bsx_header_in *hdr, const uint8_t *validators,
                    const bsx_skip_in *skip, const uint8_t *trusted_pubkeys /*n*N*32*/,
                    const uint64_t *trusted_powers /*n*N*/, const uint32_t *trusted_byte_lengths /*n*N*/,
                    uint8_t *digests, uint8_t *ed_out, uint32_t *fail);
int bsx_next_header(bsx_ctx *ctx, uint32_t n, uint32_t N, const bsx_header_in *hdr, const uint8_t *validators,
                    const bsx_step_in *step, uint8_t *digests, uint8_t *ed_out, uint8_t *data_commitments /*n*32*/,
                    uint32_t *fail);
int bsx_verify_header_dev(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, const bsx_header_in *hdr,
                          const uint8_t *validators, uint8_t *digests, uint8_t *ed_out, uint32_t *fail);
int bsx_verify_skip_dev(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, const bsx_header_in *hdr,
                        const uint8_t *validators, const bsx_skip_in *skip, const uint8_t *trusted_pubkeys,
                        const uint64_t *trusted_powers, const uint32_t *trusted_byte_lengths, uint8_t *digests,
                        uint8_t *ed_out, uint32_t *fail);
int bsx_next_header_dev(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, const bsx_header_in *hdr,
                        const uint8_t *validators, const bsx_step_in *step, uint8_t *digests, uint8_t *ed_out,
                        uint8_t *data_commitments, uint32_t *fail);
/* Ed25519 over strided records (what the verify_* entry points run over the validator array) */
int bsx_ed25519_strided_dev(bsx_ctx *ctx, void *stream, uint32_t n, const uint8_t *pks, uint32_t pk_stride,
                            const uint8_t *sigs, uint32_t sig_stride, const uint8_t *msgs, uint32_t msg_stride,
                            uint32_t msg_max, const uint8_t *msg_lens, uint32_t len_stride, const uint8_t *active,
                            uint32_t active_stride, uint8_t *out);

/* ------------------------------------------------------------------------------------------
 * K6/K7  Goldilocks u32 gates: constraint evaluation and witness generators
 * replaces: Gate::eval_unfiltered_base_batch -> PackedEvaluableBase::eval_unfiltered_base_packed and the per-gate
 *   SimpleGenerator::run_once of the five vendored gates:
 *   U32ArithmeticGate{num_ops=p0}                 PX/frontend/uint/num/u32/gates/arithmetic_u32.rs:280-349, 383-431
 *   U32AddManyGate{num_addends=p0, num_ops=p1}    PX/frontend/uint/num/u32/gates/add_many_u32.rs:107-146, 340-391
 *   U32SubtractionGate{num_ops=p0}                PX/frontend/uint/num/u32/gates/subtraction_u32.rs:101-135, 305-350
 *   ComparisonGate{num_bits=p0, num_chunks=p1}    PX/frontend/uint/num/u32/gates/comparison.rs:118-195, 441-540
 *   U32RangeCheckGate{num_input_limbs=p0}         PX/frontend/uint/num/u32/gates/range_check_u32.rs:69-91, 202-224
 * wires: bsx_gate_num_wires x rows u64, WIRE-major (wires[w*rows + r], plonky2's EvaluationVarsBaseBatch view);
 * constraints: bsx_gate_num_constraints x rows, constraints[c*rows + r], canonical (< p = 2^64 - 2^32 + 1), in the
 * order the gate yields them.  bsx_gl_gate_witness fills the generator-owned wires of every row in place.
 * ------------------------------------------------------------------------------------------ */
#define BSX_GATE_U32_ARITHMETIC 0
#define BSX_GATE_U32_ADD_MANY 1
#define BSX_GATE_U32_SUBTRACTION 2
#define BSX_GATE_U32_COMPARISON 3
#define BSX_GATE_U32_RANGE_CHECK 4
uint32_t bsx_gate_num_wires(uint32_t gate, uint32_t p0, uint32_t p1);
uint32_t bsx_gate_num_constraints(uint32_t gate, uint32_t p0, uint32_t p1);
int bsx_gl_gate_eval(bsx_ctx *ctx, uint32_t gate, uint32_t p0, uint32_t p1, const uint64_t *wires, uint32_t rows,
                     uint64_t *constraints);
int bsx_gl_gate_eval_dev(bsx_ctx *ctx, void *stream, uint32_t gate, uint32_t p0, uint32_t p1, const uint64_t *wires,
                         uint32_t rows, uint64_t *constraints);
int bsx_gl_gate_witness(bsx_ctx *ctx, uint32_t gate, uint32_t p0, uint32_t p1, uint64_t *wires, uint32_t rows);
int bsx_gl_gate_witness_dev(bsx_ctx *ctx, void *stream, uint32_t gate, uint32_t p0, uint32_t p1, uint64_t *wires,
                            uint32_t rows);

/* ------------------------------------------------------------------------------------------
 * K8  Poseidon sponge over Goldilocks (width 12, rate 8, no padding, overwrite mode, 4 outputs)
 * replaces: plonky2 hash_n_to_hash_no_pad::<PoseidonPermutation> behind poseidon_hash / poseidon_hash_pair
 *   (PX/frontend/hash/poseidon/poseidon256.rs:61-86) and mapreduce_merkle_tree_root (PX/utils/poseidon/mod.rs:9-66).
 * in: concatenated field elements; offsets[n+1]; out: n x 4 canonical elements.
 * ------------------------------------------------------------------------------------------ */
int bsx_gl_poseidon_batch(bsx_ctx *ctx, const uint64_t *in, const uint32_t *offsets, uint32_t n, uint64_t *out);
int bsx_gl_poseidon_batch_dev(bsx_ctx *ctx, void *stream, const uint64_t *in, const uint32_t *offsets, uint32_t n,
                              uint64_t *out);

/* ------------------------------------------------------------------------------------------
 * header_range = skip + prove_data_commitment for n independent ranges in ONE call
 * replaces every accelerator hint of CombinedSkipCircuit::define (BX/circuits/header_range.rs:32-59):
 *   builder.skip(..)                    TX/skip.rs:29-58 -> verify_skip (TX/builder/verify.rs:527-564)
 *   builder.prove_data_commitment(..)   BX/circuits/builder.rs:273-409 (32 map jobs + reduce tree)
 * The two argument blocks are plain structs of pointers (repr(C) on the Rust side); every pointer has the
 * meaning and size of the like-named parameter of bsx_verify_skip / bsx_prove_data_commitment.
 * The skip half and the map/reduce half run concurrently on two streams of the ctx.
 * ------------------------------------------------------------------------------------------ */
typedef struct bsx_skip_batch {
    const bsx_header_in *hdr;           /* n */
    const uint8_t *validators;          /* n*N*BSX_VAL_IN_BYTES */
    const bsx_skip_in *skip;            /* n */
    const uint8_t *trusted_pubkeys;     /* n*N*32 */
    const uint64_t *trusted_powers;     /* n*N */
    const uint32_t *trusted_byte_lengths; /* n*N */
    uint8_t *digests;                   /* out: n*bsx_verify_digest_count(1,N)*32 */
    uint8_t *ed_out;                    /* out: n*N*BSX_SIG_OUT_BYTES */
    uint32_t *fail;                     /* out: n, BSX_VFAIL_* */
} bsx_skip_batch;
typedef struct bsx_range_batch {
    const uint8_t *dh_leaf, *dh_aunts, *lb_leaf, *lb_aunts; /* n*n_jobs*B*{34,128,72,128} */
    const uint8_t *start_headers, *end_headers;             /* n*n_jobs*32 */
    const uint64_t *start_blocks, *end_blocks;              /* n */
    const uint8_t *start_header, *end_header;               /* n*32 */
    uint8_t *map_digests, *map_subchains;                   /* out: n*n_jobs*(20B-1)*32, n*n_jobs*128 */
    uint8_t *reduce_digests, *reduce_nodes;                 /* out: n*(n_jobs-1)*32, n*(n_jobs-1)*128 */
    uint8_t *data_commitments;                              /* out: n*32 */
    uint32_t *fail;                                         /* out: n, BSX_FAIL_* */
} bsx_range_batch;
int bsx_header_range(bsx_ctx *ctx, uint32_t n, uint32_t N, uint32_t n_jobs, uint32_t B, const bsx_skip_batch *skip,
                     const bsx_range_batch *range);
int bsx_header_range_dev(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, uint32_t n_jobs, uint32_t B,
                         const bsx_skip_batch *skip, const bsx_range_batch *range);

/* ------------------------------------------------------------------------------------------
 * Prover inner loops over Goldilocks: transforms, coset LDE, Poseidon Merkle caps, quotient, FRI fold
 * replaces (restated, PARITY UNPINNED -- plonky2 0.2.1 is un-vendored): the inner loops of
 *   prove_with_partition_witness (call site PX/backend/circuit/build.rs:69-75; config standard_recursion_config,
 *   PX/frontend/builder/mod.rs:69: rate_bits 3, cap_height 4, 135 wires): PolynomialBatch::from_values (ifft + lde +
 *   coset fft + MerkleTree::new over the transposed, index-bit-reversed values), compute_quotient_polys (gate constraints
 *   reduced with powers of alpha, divided by Z_H on the coset), fri_committed_trees (reduce_with_powers fold).
 * Layouts: polynomials are poly-major (element i of polynomial p at base[p * stride + i]); transforms take natural order
 * and produce BIT-REVERSED order (position i holds the value at w^bitrev(i)) -- the order the Merkle leaves are hashed in,
 * so leaf i = position i.  An extension of rate 2^r: position i of the 2^(log_n+r) outputs holds the value at
 * shift * w_N^bitrev(i).  Digests are 4 canonical words.  All pointers are device pointers.
 * ------------------------------------------------------------------------------------------ */
uint64_t bsx_gl_root_of_unity(uint32_t log_n);   /* primitive 2^log_n-th root: POWER_OF_TWO_GENERATOR^(2^(32-log_n)) */
uint64_t bsx_gl_coset_shift(void);               /* MULTIPLICATIVE_GROUP_GENERATOR */
int bsx_gl_ntt_dev(bsx_ctx *ctx, void *stream, const uint64_t *in, uint64_t *out, uint32_t log_n, uint32_t n_polys,
                   size_t in_stride, size_t out_stride, int inverse, int natural_out, uint64_t *scratch);
int bsx_gl_lde_dev(bsx_ctx *ctx, void *stream, const uint64_t *coeffs, uint64_t *out, uint32_t log_n, uint32_t rate_bits,
                   uint32_t n_polys, size_t in_stride, size_t out_stride, uint64_t shift);
size_t bsx_gl_merkle_digest_words(uint32_t n_leaves, uint32_t cap_height);
int bsx_gl_merkle_caps_dev(bsx_ctx *ctx, void *stream, const uint64_t *data, size_t poly_stride, uint32_t width,
                           uint32_t n_leaves, uint32_t cap_height, uint64_t *digests);
int bsx_gl_quotient_tables_dev(bsx_ctx *ctx, void *stream, const uint64_t *alphas, uint32_t n_alphas, uint32_t n_constraints,
                               uint32_t log_n, uint32_t rate_bits, uint64_t shift, uint64_t *alpha_pows, uint64_t *zh_inv);
int bsx_gl_gate_quotient_dev(bsx_ctx *ctx, void *stream, uint32_t gate, uint32_t p0, uint32_t p1, const uint64_t *wires,
                             uint32_t rows, const uint64_t *alpha_pows, uint32_t n_alphas, const uint64_t *zh_inv,
                             uint32_t log_block, uint64_t *out);
int bsx_gl_fri_fold_dev(bsx_ctx *ctx, void *stream, const uint64_t *in, uint32_t n_in, uint32_t arity_bits, uint64_t beta0,
                        uint64_t beta1, uint64_t *out);

/* ------------------------------------------------------------------------------------------
 * Input shaping on the device
 * The 14-leaf Tendermint tree of every header of a range and the inclusion proofs the map circuits consume.
 * Replaces the per-header host work of generate_proofs_from_header / compute_hash_from_aunts
 * (TX/input/tendermint_utils.rs:214-224,276-336,374-393), DataCommitmentInputs::get_data_commitment_inputs
 * (BX/circuits/input.rs:149-271) and the DataCommitmentOffchainInputs hints of a range (BX/circuits/builder.rs:316-333):
 * 27 SHA-256 calls per header.  The protobuf field encoders stay on the host.
 *
 * header record (BSX_HEADER_LEAVES_BYTES): bytes [0,14) = lengths of the 14 encoded fields in header order
 *   (version, chain_id, height, time, last_block_id, last_commit_hash, data_hash, validators_hash,
 *   next_validators_hash, consensus_hash, app_hash, last_results_hash, evidence_hash, proposer_address),
 *   bytes [16, 16+sum) = the fields back to back; sum <= 496.
 * bsx_header_trees: roots[n*32] = header hashes; levels (optional) n*27*32 = l0[14] l1[7] l2[4] l3[2].
 * bsx_header_range_inputs: headers = n_ranges * (n_jobs*B + 1) records, record o of range r = block
 *   start_blocks[r] + o.  latest_blocks[r] = the last block whose header can be fetched (latest_block - 2 in the
 *   reference, input.rs:160-163); NULL = end_blocks (the range ends at the chain tip).  As in the reference, the hint of
 *   job j is asked for (start + jB, start + (j+1)B) and clamps ONLY to latest_blocks[r] -- a job past the range's end still
 *   gets real proofs and headers when the chain has those blocks (the circuit disables them); records beyond
 *   latest_blocks[r] are ignored and their slots are zero.  Outputs are exactly the input arrays of bsx_range_batch;
 *   fail[r] = BSX_FAIL_INPUT_LEAF if a data_hash / last_block_id field is not 34 / 72 bytes (the host shaper returns an
 *   error there).
 * ------------------------------------------------------------------------------------------ */
#define BSX_HEADER_LEAVES_BYTES 512
int bsx_header_trees(bsx_ctx *ctx, const uint8_t *headers, uint32_t n, uint8_t *roots, uint8_t *levels);
int bsx_header_trees_dev(bsx_ctx *ctx, void *stream, const uint8_t *headers, uint32_t n, uint8_t *roots, uint8_t *levels);
int bsx_header_range_inputs(bsx_ctx *ctx, uint32_t n_ranges, uint32_t n_jobs, uint32_t B, const uint8_t *headers,
                            const uint64_t *start_blocks, const uint64_t *end_blocks, const uint64_t *latest_blocks,
                            uint8_t *dh_leaf, uint8_t *dh_aunts, uint8_t *lb_leaf, uint8_t *lb_aunts, uint8_t *start_headers,
                            uint8_t *end_headers, uint8_t *start_header, uint8_t *end_header, uint32_t *fail);
int bsx_header_range_inputs_dev(bsx_ctx *ctx, void *stream, uint32_t n_ranges, uint32_t n_jobs, uint32_t B,
                                const uint8_t *headers, const uint64_t *start_blocks, const uint64_t *end_blocks,
                                const uint64_t *latest_blocks, uint8_t *dh_leaf, uint8_t *dh_aunts, uint8_t *lb_leaf, uint8_t *lb_aunts,
                                uint8_t *start_headers, uint8_t *end_headers, uint8_t *start_header, uint8_t *end_header,
                                uint32_t *fail);

/* ------------------------------------------------------------------------------------------
 * Input shaping on the device, continued: protobuf field encoders, sign-bytes, validator records
 *
 * bsx_encode_headers: decoded headers -> the header records bsx_header_trees / bsx_header_range_inputs take.
 *   replaces the 14 `encode_vec` calls of generate_proofs_from_header (TX/input/tendermint_utils.rs:374-393).
 *   proto3 rules: zero varints and empty byte strings are omitted; last_block_id encodes to nothing when absent.
 * bsx_validator_records: one commit + its validator set -> N ValidatorVariable records (BSX_VAL_IN_BYTES, above)
 *   and/or the ValidatorHashField arrays of verify_skip's trusted set.
 *   replaces get_signed_message_data / get_validator_data_from_block (TX/input/conversion.rs:20-140): slot i <
 *   n_signatures with block_id_flag = commit carries the validator's signature and the CanonicalVote sign-bytes
 *   (length-delimited, zero padded to 124; precommit, the commit's height / round / block id, the signature's own
 *   timestamp, chain id); other slots of the set carry DUMMY_SIGNATURE, a zero message and length 32; slots beyond
 *   the set carry DUMMY_PUBLIC_KEY, power 0 and validator_byte_length 46 -- and validator_hash_field_from_block
 *   (:142-184).  Validators are passed in the set's canonical order (as the RPC returns them).  The reference also
 *   re-verifies every signature on the host (:49-50); here that is the witness kernel's job.
 *   fail[c] = BSX_FAIL_INPUT_SIGN_BYTES when a message exceeds 124 bytes or the set exceeds N (the reference panics).
 * bsx_present_on_trusted: sets present_on_trusted_header (record byte 237) of the target records.
 *   replaces update_present_on_trusted_header (:186-240): the trusted set is walked in order while
 *   total_power * (1/3) > shared (f64 arithmetic, as there); a trusted validator found in the target set (by address)
 *   whose address appears in the target commit's signatures (commit or nil votes) is marked and its power added once
 *   per matching signature.  fail[c] = BSX_FAIL_INPUT_THRESHOLD when a third is not reached (the reference asserts).
 * ------------------------------------------------------------------------------------------ */
#define BSX_FAIL_INPUT_SIGN_BYTES 512u
#define BSX_FAIL_INPUT_THRESHOLD 1024u
typedef struct bsx_header_fields {    /* a decoded tendermint Header; 464 bytes, 8-byte aligned */
    uint64_t version_block, version_app, height;
    int64_t time_seconds;
    uint32_t time_nanos;
    uint32_t chain_id_len;            /* <= 56 (larger values are read as 56; tendermint allows 50) */
    uint8_t chain_id[56];
    uint32_t parts_total;             /* last_block_id.part_set_header.total */
    uint8_t has_last_block_id;        /* 0: the field encodes to nothing (first block) */
    uint8_t hash_len[9];              /* lengths of `hashes` (0 = absent / empty; proposer_address 20) */
    uint8_t _pad[2];
    uint8_t last_block_hash[32], parts_hash[32];
    uint8_t hashes[9][32];            /* last_commit_hash, data_hash, validators_hash, next_validators_hash, consensus_hash,
                                         app_hash, last_results_hash, evidence_hash, proposer_address */
} bsx_header_fields;
typedef struct bsx_commit_in {        /* Commit + chain id; 152 bytes, 8-byte aligned */
    uint64_t height;
    uint32_t round;
    uint32_t n_signatures;            /* commit.signatures.len() = validators in the set */
    uint8_t block_hash[32], parts_hash[32];   /* commit.block_id */
    uint32_t parts_total;
    uint32_t chain_id_len;            /* <= 56 (larger values are read as 56; tendermint allows 50) */
    uint8_t chain_id[56];
    uint8_t has_block_id;             /* 0: block id omitted from the vote (nil) */
    uint8_t _pad[7];
} bsx_commit_in;
typedef struct bsx_commit_sig_in {    /* validator i of the set + commit.signatures[i]; 160 bytes, 8-byte aligned */
    uint8_t pubkey[32];
    uint8_t signature[64];
    uint64_t voting_power;
    int64_t ts_seconds;               /* the signature's timestamp */
    uint32_t ts_nanos;
    uint8_t block_id_flag;            /* 1 absent, 2 commit, 3 nil */
    uint8_t _pad[3];
    uint8_t address[20];              /* the validator's address */
    uint8_t sig_address[20];          /* commit.signatures[i].validator_address (ignored when absent) */
} bsx_commit_sig_in;
int bsx_encode_headers(bsx_ctx *ctx, uint32_t n, const bsx_header_fields *fields, uint8_t *headers /* n*512 */);
int bsx_encode_headers_dev(bsx_ctx *ctx, void *stream, uint32_t n, const bsx_header_fields *fields, uint8_t *headers);
/* validators (n*N*240), and pubkeys (n*N*32) / powers (n*N) / byte_lengths (n*N) may each be NULL (all three or none) */
int bsx_validator_records(bsx_ctx *ctx, uint32_t n, uint32_t N, const bsx_commit_in *commits, const bsx_commit_sig_in *sigs,
                          uint8_t *validators, uint8_t *pubkeys, uint64_t *powers, uint32_t *byte_lengths, uint32_t *fail);
int bsx_validator_records_dev(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, const bsx_commit_in *commits,
                              const bsx_commit_sig_in *sigs, uint8_t *validators, uint8_t *pubkeys, uint64_t *powers,
                              uint32_t *byte_lengths, uint32_t *fail);
/* target_sigs / trusted_sigs: n*N slots; n_target[c] / n_trusted[c] validators in each set; validators: in/out */
int bsx_present_on_trusted(bsx_ctx *ctx, uint32_t n, uint32_t N, const bsx_commit_sig_in *target_sigs, const uint32_t *n_target,
                           const bsx_commit_sig_in *trusted_sigs, const uint32_t *n_trusted, uint8_t *validators, uint32_t *fail);
int bsx_present_on_trusted_dev(bsx_ctx *ctx, void *stream, uint32_t n, uint32_t N, const bsx_commit_sig_in *target_sigs,
                               const uint32_t *n_target, const bsx_commit_sig_in *trusted_sigs, const uint32_t *n_trusted,
                               uint8_t *validators, uint32_t *fail);

/* ------------------------------------------------------------------------------------------
 * Witness data formats
 * HashInputData: the STARK public-input layout of one SHA accelerator, built from its request list
 *   replaces get_hash_data (PX/frontend/hash/curta/mod.rs:95-192; stream order PX/frontend/hash/curta/data.rs:63-78)
 *   with the padding of PX/frontend/hash/sha/sha256/pad.rs:15-157 / sha512/pad.rs:13-58.
 *   sha512 = 0: 64-byte chunks, padded_chunks = u32 big-endian words (16 per chunk);
 *   sha512 = 1: 128-byte chunks, padded_chunks = u64 big-endian words (16 per chunk).
 *   bufs: concatenated request buffers (fixed request: the message; variable: the whole buffer), buf_offsets[n+1];
 *   lens[n]: message length of variable requests (ignored for fixed); kinds[n]: 0 fixed, 1 variable.
 *   outputs: padded_chunks (total*16 words), end_bits / digest_bits (one per chunk), digest_indices (one per request).
 *   Call with padded_chunks = NULL to obtain *total_chunks only.  bsx_hash_input_chunks gives one request's chunk count
 *   (a circuit constant), from which the caller of the _dev form builds chunk_offsets[n+1].
 * Field-element encodings: ByteVariable = 8 big-endian bit elements (PX/frontend/vars/byte.rs:49-66);
 *   SHA-256 digest = 8 big-endian u32 words, one element each (PX/frontend/hash/sha/sha256/curta.rs:81-92).
 * ------------------------------------------------------------------------------------------ */
/* SHA-256 execution trace: the per-round table a STARK over the SHA-256 accelerator commits to, one
 * polynomial per column (column c, row r at trace[c * 2^log_rows + r]), 64 rows per padded chunk, from the
 * padded_chunks / end_bits / digest_bits of bsx_hash_input_data (sha512 = 0).
 * replaces: the row-by-row fill of HashStark::prove (PX/frontend/hash/curta/stark.rs:107-133, TraceWriter::
 *   write_row_instructions).  The reference's column assignment is starkyx's (un-vendored): this layout is our own and
 *   PARITY IS UNPINNED; tests pin it by recomputing every digest from the trace columns.
 * Columns (32-bit words as 4 little-endian byte limbs, one field element per byte; t = row within the chunk):
 *   0 w_t | 4..35 a b c d e f g h entering round t | 36 rotr6(e) 40 rotr11(e) 44 rotr25(e) 48 Sigma1 | 52 e&f 56 ~e&g 60 ch
 *   64 rotr2(a) 68 rotr13(a) 72 rotr22(a) 76 Sigma0 | 80 a&b 84 a&c 88 b&c 92 maj | 96 temp1 (100 carry) | 101 temp2 (105 carry)
 *   106 a' (110 carry) | 111 e' (115 carry) | schedule step for w_{t+16}, t < 48: 116 w_{t+1} 120 rotr7 124 rotr18 128 shr3
 *   132 sigma0 | 136 w_{t+14} 140 rotr17 144 rotr19 148 shr10 152 sigma1 | 156 w_{t+9} 160 w_{t+16} (164 carry)
 *   165 first row of chunk 166 last row 167 end_bit 168 digest_bit | 169..174 bits of t | 175 K_t */
#define BSX_SHA256_TRACE_COLS 176
/* SHA-512 (the EdDSA accelerator): the same construction with 64-bit words as 8 byte limbs and 80 rows per 128-byte chunk.
 *   0 w_t | 8..71 a..h | 72 rotr14(e) 80 rotr18(e) 88 rotr41(e) 96 Sigma1 | 104 e&f 112 ~e&g 120 ch | 128 rotr28(a) 136 rotr34(a)
 *   144 rotr39(a) 152 Sigma0 | 160 a&b 168 a&c 176 b&c 184 maj | 192 temp1 (200 carry) | 201 temp2 (209) | 210 a' (218) | 219 e' (227)
 *   schedule step for w_{t+16}, t < 64: 228 w_{t+1} 236 rotr1 244 rotr8 252 shr7 260 sigma0 | 268 w_{t+14} 276 rotr19 284 rotr61
 *   292 shr6 300 sigma1 | 308 w_{t+9} 316 w_{t+16} (324 carry) | 325 first row 326 last row 327 end_bit 328 digest_bit
 *   329..335 bits of t | 336, 337 K_t as two 32-bit limbs */
#define BSX_SHA512_TRACE_COLS 338
int bsx_sha512_trace_dev(bsx_ctx *ctx, void *stream, const uint64_t *padded_chunks, const uint8_t *end_bits,
                         const uint8_t *digest_bits, uint32_t n_chunks, uint32_t log_rows, uint64_t *trace);
int bsx_sha256_trace_dev(bsx_ctx *ctx, void *stream, const uint32_t *padded_chunks, const uint8_t *end_bits,
                         const uint8_t *digest_bits, uint32_t n_chunks, uint32_t log_rows, uint64_t *trace);
/* n_circuits accelerators of the same shape (the map circuits of a range) in one launch: circuit c reads its chunks and
 * flags chunk_stride chunks after circuit c - 1's and writes table c (BSX_SHA256_TRACE_COLS * 2^log_rows elements each) */
int bsx_sha256_trace_batch_dev(bsx_ctx *ctx, void *stream, const uint32_t *padded_chunks, const uint8_t *end_bits,
                               const uint8_t *digest_bits, uint32_t n_chunks, uint32_t n_circuits, size_t chunk_stride,
                               uint32_t log_rows, uint64_t *trace);
/* Ed25519 scalar-multiplication execution trace (the EdDSA accelerator): 256 rows per ScalarMul operation
 * (two per signature: s*G and h*A), one scalar bit per row, an affine double-and-add whose 16 field operations per row
 * carry starkyx-style witnesses (16-bit limbs; quotient `carry`; witness polynomial of the division by x - 2^16).
 * replaces: the trace fill of Ed25519Stark::prove (PX/frontend/ecc/curve25519/curta/stark.rs:182-219, write_trace_instructions
 *   over chunks_par(256)); operations collected at stark.rs:93-124.  The AIR (`scalar_mul_batch`) is starkyx's, un-vendored:
 *   this column assignment is our own and PARITY IS UNPINNED; the CPU restatements are oracle/ed_trace.py (Python integers)
 *   and oracle/ed25519.c orc_ed25519_trace (C), pinned by
 *   re-checking every operation's polynomial identity, k * P against an independent scalar multiplication and the
 *   mocha-4 fixture signatures' s*G / h*A (tests/test_oracle_ed_trace.py).
 * Row 256 m + j = step j of multiplication m; temp = 2^j P, acc = (k mod 2^j) P; next row: temp' = dbl, acc' = bit ? sum : acc.
 *   0 bit j of k | 1 real row (0 on padding) | 2 j == 0 | 3 j == 255 | 4..19 temp.x 20..35 temp.y 36..51 acc.x 52..67 acc.y
 *   68 + 92 o, o = 0..15: result[16] carry[16] witness_low[30] witness_high[30] of field operation o, where
 *     o = 0..7 is sum = acc + temp and o = 8..15 is dbl = temp + temp, an addition (x1,y1) + (x2,y2) being
 *     0 xn = x1 y2 + x2 y1 | 1 yn = y1 y2 + x1 x2 | 2 m1 = x1 y1 | 3 m2 = x2 y2 | 4 f = m1 m2 | 5 df = d f |
 *     6 x3 = xn / (1 + df): df x3 + x3 - xn = carry p | 7 y3 = yn / (1 - df): df y3 + yn - y3 = carry p
 *   witness: w(x) = (lhs(x) - result(x) - carry(x) p(x)) / (x - 2^16), stored as w_k + 2^22 in two 16-bit halves.
 * Padding rows (beyond 256 * n_muls) are the rows of 0 * (0, 1) with column 1 = 0.
 * scalars: n_muls x 32 bytes little-endian; points: n_muls x 64 bytes (x, y canonical little-endian, on the curve);
 * scratch: bsx_ed25519_trace_scratch_bytes(n_muls) bytes of device memory, 16-byte aligned; results: n_muls x 64 bytes
 * k * P, or NULL.  The points MUST be on the curve (the reference only ever multiplies decompressed points): the chains double
 * with the dedicated doubling formulas, which use the curve equation, so for any other (x, y) the rows are meaningless. */
#define BSX_ED25519_TRACE_COLS 1540
size_t bsx_ed25519_trace_scratch_bytes(uint32_t n_muls);
/* The ScalarMul operands of n_sigs signatures, gathered on the device: scalars[2i] = s_i, points[2i] = G, scalars[2i+1] =
 * h_i, points[2i+1] = A_i, from the signature bytes (sigs + i * sig_stride: R ‖ s), the lane flags (active + i *
 * active_stride; NULL = all active; an inactive lane ran on the DUMMY signature, eddsa.rs:28-30, and takes its s) and the
 * witness records ed_out of bsx_ed25519_batch[_dev] / bsx_verify_skip (BSX_SIG_OUT_BYTES each; with the validator records of
 * bsx_verify_skip: sigs = validators + 32, active = validators + 236, strides BSX_VAL_IN_BYTES) -- the operations
 * Ed25519Stark::new collects (stark.rs:93-124), in the order of the EdDSA schedule (s*G, then h*A). */
int bsx_ed25519_trace_operands_dev(bsx_ctx *ctx, void *stream, uint32_t n_sigs, const uint8_t *sigs, uint32_t sig_stride,
                                   const uint8_t *active, uint32_t active_stride, const uint8_t *ed_out, uint8_t *scalars,
                                   uint8_t *points);
int bsx_ed25519_trace_dev(bsx_ctx *ctx, void *stream, const uint8_t *scalars, const uint8_t *points, uint32_t n_muls,
                          uint32_t log_rows, void *scratch, uint8_t *results, uint64_t *trace);
/* the two halves of bsx_ed25519_trace_dev, for callers that overlap batches (two streams, two scratch buffers): the
 * latency-bound multiplication chains (-> scratch), then the bandwidth-bound row expansion (scratch -> trace) */
int bsx_ed25519_trace_points_dev(bsx_ctx *ctx, void *stream, const uint8_t *scalars, const uint8_t *points, uint32_t n_muls,
                                 void *scratch);
int bsx_ed25519_trace_rows_dev(bsx_ctx *ctx, void *stream, const uint8_t *scalars, const uint8_t *points, uint32_t n_muls,
                               uint32_t log_rows, const void *scratch, uint8_t *results, uint64_t *trace);
uint32_t bsx_hash_input_chunks(int sha512, uint32_t buf_len, int variable);
int bsx_hash_input_data(bsx_ctx *ctx, int sha512, uint32_t n_req, const uint8_t *bufs, const uint32_t *buf_offsets,
                        const uint32_t *lens, const uint8_t *kinds, void *padded_chunks, uint8_t *end_bits,
                        uint8_t *digest_bits, uint32_t *digest_indices, uint32_t *total_chunks);
int bsx_hash_input_data_dev(bsx_ctx *ctx, void *stream, int sha512, uint32_t n_req, const uint8_t *bufs,
                            const uint32_t *buf_offsets, const uint32_t *lens, const uint8_t *kinds,
                            const uint32_t *chunk_offsets, uint32_t total_chunks, void *padded_chunks, uint8_t *end_bits,
                            uint8_t *digest_bits, uint32_t *digest_indices);
int bsx_witness_pack_bytes(bsx_ctx *ctx, const uint8_t *bytes, size_t n, uint64_t *elements /* n*8 */);
int bsx_witness_unpack_bytes(bsx_ctx *ctx, const uint64_t *elements, size_t n, uint8_t *bytes, uint32_t *not_bits);
int bsx_witness_pack_bytes_dev(bsx_ctx *ctx, void *stream, const uint8_t *bytes, size_t n, uint64_t *elements);
int bsx_witness_unpack_bytes_dev(bsx_ctx *ctx, void *stream, const uint64_t *elements, size_t n, uint8_t *bytes,
                                 uint32_t *not_bits);
int bsx_witness_pack_u32_be_dev(bsx_ctx *ctx, void *stream, const uint8_t *bytes, size_t n_words, uint64_t *elements);

#ifdef __cplusplus
}
#endif
#endif /* BSX_H */

/*
 * bsx_oracle.h -- CPU restatement of the Blobstream X witness-generation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (blobstreamx_b200/, include/)
 * may include, link or call this.  Allowed users: tests/, __graft_entry__.smoke(),
 * and bench.py's cpu_baseline / --impl reference legs (as the thing checked
 * against or timed as "the CPU path", never as the GPU product).
 *
 * Parity status: PINNED for SHA-256/512, Tendermint Merkle, data commitment,
 * validators hash, Ed25519 equation (fixtures + KATs of the reference, see
 * tests/golden/ and tests/test_oracle_golden.py); `decompress` root convention
 * pinned by audit prose only; Poseidon pinned by the single KAT at
 * PX/frontend/hash/poseidon/poseidon256.rs:172-178.
 *
 * PX = contracts/lib/succinctx/plonky2x/core/src, TX = contracts/lib/tendermintx/circuits,
 * BX = the blobstreamx tree itself (all under /root/reference).
 */
#ifndef BSX_ORACLE_H
#define BSX_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- SHA (starkyx SHAPure == FIPS 180-4; call sites PX/frontend/hash/sha/sha256/curta.rs:94-102,
 *      PX/frontend/hash/sha/sha512/curta.rs:103-111) ---- */
void orc_sha256(const uint8_t *msg, size_t len, uint8_t out[32]);
void orc_sha512(const uint8_t *msg, size_t len, uint8_t out[64]);
void orc_sha256_batch(const uint8_t *msgs, const uint32_t *offsets, uint32_t n, uint8_t *digests);
void orc_sha512_batch(const uint8_t *msgs, const uint32_t *offsets, uint32_t n, uint8_t *digests);
/* padded chunk layout of the circuit (PX/frontend/hash/sha/sha256/pad.rs:15-43 fixed,
 * :63-157 variable).  Returns number of 64-byte chunks written to out. */
uint32_t orc_sha256_pad_fixed(const uint8_t *msg, uint32_t len, uint8_t *out);
uint32_t orc_sha256_pad_variable(const uint8_t *buf, uint32_t buf_len, uint32_t len, uint8_t *out,
                                 uint32_t *last_chunk);
uint32_t orc_sha512_pad_variable(const uint8_t *buf, uint32_t buf_len, uint32_t len, uint8_t *out,
                                 uint32_t *last_chunk);

/* ---- Tendermint Merkle (PX/frontend/merkle/tendermint.rs:62-214, TX/input/tendermint_utils.rs) ---- */
void orc_leaf_hash(const uint8_t *leaf, size_t len, uint8_t out[32]);
void orc_inner_hash(const uint8_t l[32], const uint8_t r[32], uint8_t out[32]);
/* variable-shape tree root (TX/input/tendermint_utils.rs:276-349) over n byte-slices */
void orc_tm_root_from_slices(const uint8_t *items, const uint32_t *offsets, uint32_t n, uint8_t out[32]);
/* aunts (leaf side first) of leaf `index` in the same tree (TX/input/tendermint_utils.rs:276-336); returns the depth */
uint32_t orc_tm_aunts_from_slices(const uint8_t *items, const uint32_t *offsets, uint32_t n, uint32_t index, uint8_t *aunts,
                                  uint8_t root[32]);
/* inputs of the map circuits of one range from its encoded headers (BX/circuits/input.rs:149-271, builder.rs:316-333) */
int orc_header_range_inputs(uint32_t n_jobs, uint32_t B, const uint8_t *headers, uint64_t start, uint64_t end, uint64_t latest,
                            uint8_t *dh_leaf, uint8_t *dh_aunts, uint8_t *lb_leaf, uint8_t *lb_aunts, uint8_t *start_headers,
                            uint8_t *end_headers, uint8_t start_header[32], uint8_t end_header[32]);
/* off-chain input shaping: the 14 protobuf field encoders of a header (TX/input/tendermint_utils.rs:374-393),
 * CanonicalVote sign-bytes (TX/input/conversion.rs:34-39), validator records (:59-184), trusted-set walk (:186-240) */
uint32_t orc_encode_header_fields(uint64_t version_block, uint64_t version_app, const uint8_t *chain_id, uint32_t chain_id_len,
                                  uint64_t height, int64_t time_secs, uint32_t time_nanos, int has_last_block_id,
                                  const uint8_t last_block_hash[32], uint32_t parts_total, const uint8_t parts_hash[32],
                                  const uint8_t *hashes, const uint8_t hash_len[9], uint8_t lens[14], uint8_t *out);
uint32_t orc_vote_sign_bytes(const uint8_t *chain_id, uint32_t chain_id_len, uint64_t height, uint64_t round, int has_block_id,
                             const uint8_t block_hash[32], uint32_t parts_total, const uint8_t parts_hash[32], int64_t ts_secs,
                             uint32_t ts_nanos, uint8_t *out);
int orc_validator_records(uint32_t N, uint32_t n_sigs, const uint8_t *chain_id, uint32_t chain_id_len, uint64_t height, uint64_t round,
                          int has_block_id, const uint8_t block_hash[32], uint32_t parts_total, const uint8_t parts_hash[32],
                          const uint8_t *pubkeys, const uint8_t *signatures, const uint64_t *powers, const int64_t *ts_secs,
                          const uint32_t *ts_nanos, const uint8_t *flags, uint8_t *records, uint8_t *hf_pubkeys, uint64_t *hf_powers,
                          uint32_t *hf_lens);
int orc_present_on_trusted(uint32_t n_target, const uint8_t *tg_addr, const uint8_t *tg_sig_addr, const uint8_t *tg_flags,
                           const uint64_t *tg_powers, uint32_t n_trusted, const uint8_t *tr_addr, uint8_t *records);
/* fixed-shape proof: digests = [leaf?] + (left,right) per level, schedule order.
 * path_bits bit i = path_indices[i].  hashed_leaf!=0 => `leaf` is the 32-byte digest. */
void orc_tm_merkle_proof(const uint8_t *leaf, uint32_t leaf_len, const uint8_t *aunts, uint32_t depth,
                         uint32_t path_bits, int hashed_leaf, uint8_t *digests, uint8_t root[32]);
/* fixed-shape tree over N leaf digests, padded to P=2^ceil(log2 N); inner = (P-1) raw inner
 * hashes, layer-major (before select); root after select. nb_enabled is a field element. */
void orc_tm_merkle_tree(const uint8_t *leaf_digests, uint32_t N, uint64_t nb_enabled, uint8_t *inner,
                        uint8_t root[32]);

/* ---- Blobstream circuits (BX/circuits/builder.rs) ---- */
void orc_encode_data_root_tuple(const uint8_t data_hash[32], uint64_t height, uint8_t out[64]);
/* get_data_commitment<B>: digests = B leaf digests then P-1 inner; returns root */
void orc_get_data_commitment(const uint8_t *data_hashes, uint32_t B, uint64_t start, uint64_t end,
                             uint8_t *digests, uint8_t root[32], uint32_t *fail);

#define ORC_SUBCHAIN_BYTES 128
/* subchain record: [0] is_enabled, [4..8) fail mask LE, [8..16) start_block LE, [16..24) end_block LE,
 * [24..56) start_header, [56..88) end_header, [88..120) data_merkle_root, rest zero. */
#define ORC_FAIL_PREV_HEADER 1u
#define ORC_FAIL_DATA_HASH 2u
#define ORC_FAIL_END_HEADER 4u
#define ORC_FAIL_BATCH_END_HEADER 8u
#define ORC_FAIL_END_LT_START 16u
#define ORC_FAIL_REDUCE_LINK 32u
#define ORC_FAIL_RANGE 64u
#define ORC_FAIL_RESULT 128u

/* prove_subchain<B> (BX/circuits/builder.rs:150-271).  digests: (20B-1)*32 bytes, schedule order A.7. */
void orc_prove_subchain(uint32_t B, const uint8_t *dh_leaf /*B*34*/, const uint8_t *dh_aunts /*B*128*/,
                        const uint8_t *lb_leaf /*B*72*/, const uint8_t *lb_aunts /*B*128*/,
                        const uint8_t start_header[32], const uint8_t end_header[32],
                        uint64_t batch_start, uint64_t batch_end, uint64_t global_end,
                        const uint8_t global_end_header[32], uint8_t *digests, uint8_t *subchain);
/* reduce closure (BX/circuits/builder.rs:337-395): out record + the raw sha256 digest */
void orc_reduce_subchain(const uint8_t *left, const uint8_t *right, uint8_t *out, uint8_t digest[32]);
/* map (n_jobs) + full reduce tree; job arrays are concatenated.  reduce_digests: (n_jobs-1)*32,
 * reduce_nodes: (n_jobs-1)*128 layer-major.  OpenMP over jobs when threads>1. */
void orc_prove_data_commitment(uint32_t n_jobs, uint32_t B, const uint8_t *dh_leaf, const uint8_t *dh_aunts,
                               const uint8_t *lb_leaf, const uint8_t *lb_aunts, const uint8_t *start_headers,
                               const uint8_t *end_headers, uint64_t start_block, const uint8_t start_header[32],
                               uint64_t end_block, const uint8_t end_header[32], uint8_t *map_digests,
                               uint8_t *map_subchains, uint8_t *reduce_digests, uint8_t *reduce_nodes,
                               uint8_t data_commitment[32], uint32_t *fail, int threads);

/* reduce tree alone over map outputs (what every rank runs after the all-gather) */
void orc_reduce_subchains(uint32_t n_jobs, uint32_t B, const uint8_t *map_subchains, uint64_t start_block,
                          const uint8_t start_header[32], uint64_t end_block, const uint8_t end_header[32],
                          uint8_t *reduce_digests, uint8_t *reduce_nodes, uint8_t data_commitment[32], uint32_t *fail);

/* ---- tendermintx gadgets (TX/builder/{validator,shared,verify}.rs) ---- */
/* marshal_int64_varint: always 9 output bytes; returns significant length */
uint32_t orc_marshal_int64_varint(uint64_t v, uint8_t out[9]);
/* marshal_tendermint_validator: 46 bytes, returns significant length (38..47 -> capped 46) */
uint32_t orc_marshal_validator(const uint8_t pubkey[32], uint64_t power, uint8_t out[46]);
/* hash_validator_set<N>: digests = N leaf digests + (P-1) inner; returns root */
void orc_hash_validator_set(uint32_t N, const uint8_t *pubkeys, const uint64_t *powers,
                            const uint32_t *byte_lengths, uint64_t nb_enabled, uint8_t *digests,
                            uint8_t root[32]);

/* ---- Ed25519 (PX/frontend/ecc/curve25519/ed25519/eddsa.rs:131-204; result hints
 *      PX/frontend/ecc/curve25519/curta/result_hint.rs:21-50) ---- */
#define ORC_SIG_OUT_BYTES 576
/* per-signature witness record (all little-endian canonical field elements):
 * [0..64) sha512 digest, [64..96) h = LE512(digest) mod l, [96..136) div = LE512(digest) / l (40 B),
 * [136..200) sG (x,y), [200..264) A (x,y), [264..296) A_root, [296..360) hA, [360..424) Rp,
 * [424..456) R_root, [456..520) sum = Rp + hA, [520..524) flags LE, rest zero.
 * flags: bit0 s<l, bit1 A decompress ok, bit2 R decompress ok, bit3 sG==sum (verified). */
void orc_ed25519_witness(const uint8_t pk[32], const uint8_t sig[64], const uint8_t *msg, uint32_t msg_len,
                         uint8_t *out);
void orc_ed25519_batch(uint32_t n, const uint8_t *pks, const uint8_t *sigs, const uint8_t *msgs /*n*124*/,
                       const uint32_t *msg_lens, const uint8_t *active, uint8_t *out, int threads);
/* primitive EC hints */
int orc_ed25519_decompress(const uint8_t in[32], uint8_t xy[64], uint8_t root[32]);
void orc_ed25519_scalar_mul(const uint8_t scalar[32], const uint8_t xy[64], uint8_t out[64]);
void orc_ed25519_add(const uint8_t a[64], const uint8_t b[64], uint8_t out[64]);
void orc_ed25519_affine_double_and_add(const uint8_t scalar[32], const uint8_t xy[64], uint8_t out[64]);

/* ---- verify_header / verify_skip / verify_step schedules (TX/builder/verify.rs) ---- */
/* Validator record layout (input), VAL_IN_BYTES each:
 * [0..32) pubkey, [32..96) signature R||s, [96..220) message (124), [220..224) message_byte_length LE,
 * [224..232) voting_power LE, [232..236) validator_byte_length LE, [236] signed, [237] present_on_trusted */
#define ORC_VAL_IN_BYTES 240
#define ORC_HASHPROOF_BYTES (34 + 128) /* leaf[34] + 4 aunts */
#define ORC_BLOCKIDPROOF_BYTES (72 + 128)
typedef struct {
    uint32_t n_validators;        /* VALIDATOR_SET_SIZE_MAX */
    const uint8_t *validators;    /* n * ORC_VAL_IN_BYTES */
    uint64_t nb_enabled;
    const uint8_t *header;        /* 32 */
    uint64_t height;
    uint64_t round;
    const uint8_t *chain_id_enc;  /* 64-byte buffer: protobuf-encoded chain id, zero padded */
    uint32_t chain_id_enc_len;
    const uint8_t *chain_id_aunts;   /* 128 */
    const uint8_t *height_aunts;     /* 128 */
    uint32_t height_enc_len;         /* enc_height_byte_length */
    const uint8_t *validators_hash_proof; /* ORC_HASHPROOF_BYTES */
    const uint8_t *expected_chain_id;
    uint32_t expected_chain_id_len;
} orc_verify_header_in;

/* verify_header: sha256 digests (254 for N=100) in schedule order A.4, ed records n*ORC_SIG_OUT_BYTES.
 * returns fail mask (0 = all circuit assertions hold). */
uint32_t orc_verify_header(const orc_verify_header_in *in, uint8_t *sha256_digests, uint8_t *ed_out, int threads);

typedef struct {
    orc_verify_header_in target;
    uint64_t trusted_block;
    const uint8_t *trusted_header; /* 32 */
    uint32_t skip_max;
    const uint8_t *trusted_validators_hash_proof; /* ORC_HASHPROOF_BYTES */
    const uint8_t *trusted_pubkeys;               /* n*32 */
    const uint64_t *trusted_powers;               /* n */
    const uint32_t *trusted_byte_lengths;         /* n */
    uint64_t trusted_nb_enabled;
} orc_verify_skip_in;
/* verify_skip: sha256 digests 490 (A.5) */
uint32_t orc_verify_skip(const orc_verify_skip_in *in, uint8_t *sha256_digests, uint8_t *ed_out, int threads);

typedef struct {
    orc_verify_header_in next;
    uint64_t prev_block;
    const uint8_t *prev_header; /* 32 */
    const uint8_t *last_block_id_proof;        /* ORC_BLOCKIDPROOF_BYTES, against next header */
    const uint8_t *prev_next_validators_proof; /* ORC_HASHPROOF_BYTES, against prev header */
    const uint8_t *data_hash_proof;            /* ORC_HASHPROOF_BYTES, against prev header */
} orc_verify_step_in;
/* verify_step + prove_next_header_data_commitment: sha256 digests 282 (A.6) */
uint32_t orc_next_header(const orc_verify_step_in *in, uint8_t *sha256_digests, uint8_t *ed_out,
                         uint8_t data_commitment[32], int threads);

#define ORC_VFAIL_SIG 1u
#define ORC_VFAIL_VALHASH 2u
#define ORC_VFAIL_VALHASH_PROOF 4u
#define ORC_VFAIL_THRESHOLD 8u
#define ORC_VFAIL_MESSAGE 16u
#define ORC_VFAIL_CHAIN_ID 32u
#define ORC_VFAIL_HEIGHT 64u
#define ORC_VFAIL_TRUSTED_PROOF 128u
#define ORC_VFAIL_TRUSTED_VALHASH 256u
#define ORC_VFAIL_TRUSTED_PRESENT 512u
#define ORC_VFAIL_TRUSTED_THRESHOLD 1024u
#define ORC_VFAIL_SKIP_DISTANCE 2048u
#define ORC_VFAIL_PREV_HEADER 4096u
#define ORC_VFAIL_NEXT_VALS 8192u
#define ORC_VFAIL_DATA_HASH_PROOF 16384u

/* ---- HashInputData layout (PX/frontend/hash/curta/mod.rs:95-192) ---- */
/* requests: kind 0 fixed / 1 variable; returns n_chunks; padded_chunks as big-endian-decoded
 * u32 words (16/chunk); end_bits, digest_bits one per chunk; digest_indices one per request. */
uint32_t orc_sha256_hash_input_data(uint32_t n_req, const uint8_t *bufs, const uint32_t *buf_offsets,
                                    const uint32_t *lens, const uint8_t *kinds, uint32_t *padded_chunks,
                                    uint8_t *end_bits, uint8_t *digest_bits, uint32_t *digest_indices);

uint32_t orc_sha512_hash_input_data(uint32_t n_req, const uint8_t *bufs, const uint32_t *buf_offsets,
                                    const uint32_t *lens, const uint8_t *kinds, uint64_t *padded_chunks,
                                    uint8_t *end_bits, uint8_t *digest_bits, uint32_t *digest_indices);

/* ---- Goldilocks / gates / Poseidon (PX/frontend/uint/num/u32/gates/*.rs; plonky2 0.2.1) ---- */
#define ORC_GATE_U32_ARITHMETIC 0
#define ORC_GATE_U32_ADD_MANY 1
#define ORC_GATE_U32_SUBTRACTION 2
#define ORC_GATE_U32_COMPARISON 3
#define ORC_GATE_U32_RANGE_CHECK 4
uint32_t orc_gate_num_constraints(uint32_t gate, uint32_t p0, uint32_t p1);
uint32_t orc_gate_num_wires(uint32_t gate, uint32_t p0, uint32_t p1);
/* wires wire-major [w*rows+r]; constraints [c*rows+r] */
int orc_gate_eval(uint32_t gate, uint32_t p0, uint32_t p1, const uint64_t *wires, uint32_t rows,
                  uint64_t *constraints, int threads);
/* fill dependent wires of a row from its inputs (SimpleGenerator::run_once) */
int orc_gate_witness(uint32_t gate, uint32_t p0, uint32_t p1, uint64_t *wires, uint32_t rows, int threads);
void orc_poseidon_permute(uint64_t state[12]);
void orc_poseidon_hash_no_pad(const uint64_t *in, uint32_t n, uint64_t out[4]);
void orc_poseidon_batch(const uint64_t *in, const uint32_t *offsets, uint32_t n, uint64_t *out, int threads);
const uint64_t *orc_poseidon_round_constants(void); /* 360 */

/* ---- prover inner loops over Goldilocks (oracle/plonk.c; plonky2 0.2.1 un-vendored: PARITY UNPINNED) ---- */
uint64_t orc_gl_root_of_unity(uint32_t log_n);
uint64_t orc_gl_coset_shift(void);
void orc_gl_ntt(uint64_t *x, uint32_t log_n, int inverse);
void orc_gl_lde(const uint64_t *coeffs, uint32_t log_n, uint32_t rate_bits, uint64_t shift, uint64_t *out);
void orc_gl_merkle(const uint64_t *leaves, uint32_t width, uint32_t n_leaves, uint32_t cap_height, uint64_t *digests);
void orc_gl_quotient_combine(const uint64_t *constraints, uint32_t n_constraints, uint32_t rows, const uint64_t *alphas,
                             uint32_t n_alphas, const uint64_t *zh_inv, uint64_t *out);
/* SHA-256 execution trace (oracle/trace.c): columns documented in include/bsx.h (BSX_SHA256_TRACE_COLS = 176) */
void orc_sha256_trace(const uint32_t *chunks, const uint8_t *end_bits, const uint8_t *digest_bits, uint32_t n_chunks,
                      uint32_t log_rows, uint64_t *trace);
void orc_sha512_trace(const uint64_t *chunks, const uint8_t *end_bits, const uint8_t *digest_bits, uint32_t n_chunks,
                      uint32_t log_rows, uint64_t *trace);
/* Ed25519 scalar-multiplication trace (oracle/ed25519.c, restated a second time in Python: oracle/ed_trace.py): columns in
 * include/bsx.h (BSX_ED25519_TRACE_COLS = 1540); returns 1 (0 = an identity failed to hold, which cannot happen for curve points) */
int orc_ed25519_trace(const uint8_t *scalars, const uint8_t *points, uint32_t n_muls, uint32_t log_rows, uint8_t *results,
                      uint64_t *trace, int threads);
void orc_gl_fri_fold(const uint64_t *in, uint32_t n_in, uint32_t arity_bits, uint64_t beta0, uint64_t beta1, uint64_t *out);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif

/*
 * oracle/verify.c -- request schedules and assertion logic of the tendermintx light-client
 * gadgets: verify_header (TX/builder/verify.rs:225-329), verify_trusted_validators (:356-433),
 * verify_skip (:527-564), verify_step (:468-505) and Blobstream's
 * prove_next_header_data_commitment (BX/circuits/builder.rs:411-443).
 * The SHA-256 digests are emitted in the exact order the circuit enqueues Curta requests
 * (SURVEY Appendix A.4-A.6).  TEST INFRASTRUCTURE ONLY (see bsx_oracle.h).
 */
#include "bsx_oracle.h"
#include <stdlib.h>
#include <string.h>

static uint64_t get64(const uint8_t *p) { uint64_t v = 0; for (int i = 7; i >= 0; i--) v = (v << 8) | p[i]; return v; }
static uint32_t get32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

/* running-AND enabled mask used all over the circuits (e.g. verify.rs:296-303) */
static int enabled_at(uint32_t i, uint64_t nb_enabled) { return (uint64_t)i < nb_enabled; }

/* TX/builder/voting.rs:31-79 */
static int voting_threshold(uint32_t n, const uint8_t *validators, uint64_t nb_enabled, uint64_t num,
                            uint64_t den, const uint8_t *include) {
    uint64_t total = 0, acc = 0;
    for (uint32_t i = 0; i < n; i++) {
        uint64_t p = get64(validators + ORC_VAL_IN_BYTES * (size_t)i + 224);
        if (enabled_at(i, nb_enabled)) total += p;
        if (include[i]) acc += p;
    }
    return acc * den > total * num;
}

/* TX/builder/validator.rs:80-183 */
static int validator_message_ok(const uint8_t *v, int is_enabled, const uint8_t header[32], uint64_t height,
                                uint64_t round) {
    const uint8_t *msg = v + 96;
    int signed_ = v[236] != 0;
    const uint8_t *hash_at = (round == 0) ? msg + 16 : msg + 25;
    int hash_in_message = memcmp(hash_at, header, 32) == 0;
    int is_precommit = msg[1] == 8 && msg[2] == 2;
    uint8_t le[8];
    for (int i = 0; i < 8; i++) le[i] = (uint8_t)(height >> (8 * i));
    int height_ok = memcmp(le, msg + 4, 8) == 0;
    for (int i = 0; i < 8; i++) le[i] = (uint8_t)(round >> (8 * i));
    int round_ok = (round == 0) ? 1 : memcmp(le, msg + 13, 8) == 0;
    int valid = signed_ && is_enabled && hash_in_message && is_precommit && height_ok && round_ok;
    return signed_ == valid; /* assert signed == signed*enabled*... */
}

static uint32_t compute_validators_hash_fields(uint32_t n, const uint8_t *pubkeys, const uint64_t *powers,
                                               const uint32_t *blens, uint64_t nb_enabled, uint8_t *digests,
                                               uint8_t root[32]) {
    orc_hash_validator_set(n, pubkeys, powers, blens, nb_enabled, digests, root);
    uint32_t P = 1;
    while (P < n) P *= 2;
    return n + P - 1;
}

/* verify_header, TX/builder/verify.rs:225-329 */
uint32_t orc_verify_header(const orc_verify_header_in *in, uint8_t *sha256_digests, uint8_t *ed_out, int threads) {
    uint32_t n = in->n_validators, fail = 0;
    uint8_t *d = sha256_digests;
    /* (1) EdDSA batch, :239-251 */
    uint8_t *pks = (uint8_t *)malloc((size_t)n * 32), *sigs = (uint8_t *)malloc((size_t)n * 64);
    uint8_t *msgs = (uint8_t *)malloc((size_t)n * 124), *act = (uint8_t *)malloc(n);
    uint32_t *lens = (uint32_t *)malloc((size_t)n * 4), *blens = (uint32_t *)malloc((size_t)n * 4);
    uint64_t *powers = (uint64_t *)malloc((size_t)n * 8);
    for (uint32_t i = 0; i < n; i++) {
        const uint8_t *v = in->validators + ORC_VAL_IN_BYTES * (size_t)i;
        memcpy(pks + 32 * (size_t)i, v, 32);
        memcpy(sigs + 64 * (size_t)i, v + 32, 64);
        memcpy(msgs + 124 * (size_t)i, v + 96, 124);
        lens[i] = get32(v + 220);
        powers[i] = get64(v + 224);
        blens[i] = get32(v + 232);
        act[i] = v[236];
    }
    orc_ed25519_batch(n, pks, sigs, msgs, lens, act, ed_out, threads);
    for (uint32_t i = 0; i < n; i++)
        if ((ed_out[ORC_SIG_OUT_BYTES * (size_t)i + 520] & 0xf) != 0xf) fail |= ORC_VFAIL_SIG;
    /* (2) validators hash, :253-267 */
    uint8_t vh[32];
    d += 32 * (size_t)compute_validators_hash_fields(n, pks, powers, blens, in->nb_enabled, d, vh);
    if (memcmp(in->validators_hash_proof + 2, vh, 32) != 0) fail |= ORC_VFAIL_VALHASH;
    /* marshal_int64_varint asserts bit 63 of its argument is zero (TX/builder/shared.rs:77-80): voting powers here, the
     * height in verify_block_height (:178); verify_non_negative_round asserts the round's sign bit (validator.rs:73-78) */
    for (uint32_t i = 0; i < n; i++)
        if (powers[i] >> 63) fail |= ORC_VFAIL_VALHASH;
    if (in->height >> 63) fail |= ORC_VFAIL_HEIGHT;
    if (in->round >> 63) fail |= ORC_VFAIL_MESSAGE;
    /* (3) validators-hash inclusion proof, :269-277 ; VALIDATORS_HASH_INDEX = 7 */
    uint8_t root[32];
    orc_tm_merkle_proof(in->validators_hash_proof, 34, in->validators_hash_proof + 34, 4, 7, 0, d, root);
    d += 9 * 32;
    if (memcmp(root, in->header, 32) != 0) fail |= ORC_VFAIL_VALHASH_PROOF;
    /* (4) 2/3 threshold over `signed`, :279-288 */
    if (!voting_threshold(n, in->validators, in->nb_enabled, 2, 3, act)) fail |= ORC_VFAIL_THRESHOLD;
    /* (5) per-validator message checks, :290-312 */
    for (uint32_t i = 0; i < n; i++)
        if (!validator_message_ok(in->validators + ORC_VAL_IN_BYTES * (size_t)i, enabled_at(i, in->nb_enabled),
                                  in->header, in->height, in->round))
            fail |= ORC_VFAIL_MESSAGE;
    /* (6) chain id, :181-223 ; CHAIN_ID_INDEX = 1 */
    {
        uint8_t buf[64];
        memset(buf, 0, 64);
        buf[0] = 0x00;
        memcpy(buf + 1, in->chain_id_enc, 52);
        uint8_t lh[32];
        orc_sha256(buf, 1 + in->chain_id_enc_len, lh);
        memcpy(d, lh, 32);
        d += 32;
        orc_tm_merkle_proof(lh, 32, in->chain_id_aunts, 4, 1, 1, d, root);
        d += 8 * 32;
        if (memcmp(root, in->header, 32) != 0) fail |= ORC_VFAIL_CHAIN_ID;
        if (memcmp(in->chain_id_enc + 2, in->expected_chain_id, in->expected_chain_id_len) != 0)
            fail |= ORC_VFAIL_CHAIN_ID;
    }
    /* (7) height, TX/builder/shared.rs:169-207 ; BLOCK_HEIGHT_INDEX = 2 */
    {
        uint8_t buf[64];
        memset(buf, 0, 64);
        buf[0] = 0x00;
        buf[1] = 0x08;
        orc_marshal_int64_varint(in->height, buf + 2); /* height_proof.height == expected height */
        uint8_t lh[32];
        orc_sha256(buf, 1 + in->height_enc_len, lh);
        memcpy(d, lh, 32);
        d += 32;
        orc_tm_merkle_proof(lh, 32, in->height_aunts, 4, 2, 1, d, root);
        d += 8 * 32;
        if (memcmp(root, in->header, 32) != 0) fail |= ORC_VFAIL_HEIGHT;
    }
    free(pks); free(sigs); free(msgs); free(act); free(lens); free(blens); free(powers);
    return fail;
}

/* verify_skip, TX/builder/verify.rs:527-564 (+ verify_trusted_validators :356-433) */
uint32_t orc_verify_skip(const orc_verify_skip_in *in, uint8_t *sha256_digests, uint8_t *ed_out, int threads) {
    uint32_t n = in->target.n_validators, fail = 0;
    uint8_t *d = sha256_digests;
    /* verify_skip_distance :507-525 */
    uint64_t target_block = in->target.height;
    if (!(target_block > in->trusted_block + 1)) fail |= ORC_VFAIL_SKIP_DISTANCE;
    if (!(target_block <= in->trusted_block + in->skip_max)) fail |= ORC_VFAIL_SKIP_DISTANCE;
    /* trusted validators-hash proof against trusted header */
    uint8_t root[32];
    orc_tm_merkle_proof(in->trusted_validators_hash_proof, 34, in->trusted_validators_hash_proof + 34, 4, 7, 0, d, root);
    d += 9 * 32;
    if (memcmp(root, in->trusted_header, 32) != 0) fail |= ORC_VFAIL_TRUSTED_PROOF;
    uint8_t vh[32];
    d += 32 * (size_t)compute_validators_hash_fields(n, in->trusted_pubkeys, in->trusted_powers,
                                                     in->trusted_byte_lengths, in->trusted_nb_enabled, d, vh);
    if (memcmp(vh, in->trusted_validators_hash_proof + 2, 32) != 0) fail |= ORC_VFAIL_TRUSTED_VALHASH;
    for (uint32_t i = 0; i < n; i++)
        if (in->trusted_powers[i] >> 63) fail |= ORC_VFAIL_TRUSTED_VALHASH; /* marshal_int64_varint, shared.rs:77-80 */
    /* present_on_trusted_header => signed ; and really present (O(N^2) pubkey match) */
    uint8_t *present = (uint8_t *)malloc(n);
    for (uint32_t i = 0; i < n; i++) {
        const uint8_t *v = in->target.validators + ORC_VAL_IN_BYTES * (size_t)i;
        present[i] = v[237];
        if (v[237] && !v[236]) fail |= ORC_VFAIL_TRUSTED_PRESENT;
        int found = 0;
        for (uint32_t j = 0; j < n; j++)
            if (memcmp(v, in->trusted_pubkeys + 32 * (size_t)j, 32) == 0) found = 1;
        if (v[237] && !found) fail |= ORC_VFAIL_TRUSTED_PRESENT;
    }
    if (!voting_threshold(n, in->target.validators, in->target.nb_enabled, 1, 3, present))
        fail |= ORC_VFAIL_TRUSTED_THRESHOLD;
    free(present);
    fail |= orc_verify_header(&in->target, d, ed_out, threads);
    return fail;
}

/* verify_step (:468-505) then prove_next_header_data_commitment (BX/circuits/builder.rs:411-443) */
uint32_t orc_next_header(const orc_verify_step_in *in, uint8_t *sha256_digests, uint8_t *ed_out,
                         uint8_t data_commitment[32], int threads) {
    uint32_t fail = 0;
    uint8_t *d = sha256_digests;
    uint32_t n = in->next.n_validators, P = 1;
    while (P < n) P *= 2;
    fail |= orc_verify_header(&in->next, d, ed_out, threads);
    d += 32 * (size_t)(n + P - 1 + 9 + 9 + 9);
    uint8_t root[32];
    /* verify_prev_header_in_header :137-154 ; LAST_BLOCK_ID_INDEX = 4 */
    orc_tm_merkle_proof(in->last_block_id_proof, 72, in->last_block_id_proof + 72, 4, 4, 0, d, root);
    d += 9 * 32;
    if (memcmp(root, in->next.header, 32) != 0) fail |= ORC_VFAIL_PREV_HEADER;
    if (memcmp(in->last_block_id_proof + 2, in->prev_header, 32) != 0) fail |= ORC_VFAIL_PREV_HEADER;
    /* verify_prev_header_next_validators_hash :156-179 ; NEXT_VALIDATORS_HASH_INDEX = 8 */
    orc_tm_merkle_proof(in->prev_next_validators_proof, 34, in->prev_next_validators_proof + 34, 4, 8, 0, d, root);
    d += 9 * 32;
    if (memcmp(root, in->prev_header, 32) != 0) fail |= ORC_VFAIL_NEXT_VALS;
    if (memcmp(in->next.validators_hash_proof + 2, in->prev_next_validators_proof + 2, 32) != 0)
        fail |= ORC_VFAIL_NEXT_VALS;
    /* prove_next_header_data_commitment ; DATA_HASH_INDEX = 6 */
    orc_tm_merkle_proof(in->data_hash_proof, 34, in->data_hash_proof + 34, 4, 6, 0, d, root);
    d += 9 * 32;
    if (memcmp(root, in->prev_header, 32) != 0) fail |= ORC_VFAIL_DATA_HASH_PROOF;
    uint8_t tup[64];
    orc_encode_data_root_tuple(in->data_hash_proof + 2, in->prev_block, tup);
    orc_leaf_hash(tup, 64, d);
    memcpy(data_commitment, d, 32);
    return fail;
}

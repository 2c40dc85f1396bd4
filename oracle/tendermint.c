/*
 * oracle/tendermint.c -- Tendermint (RFC-6962 style) Merkle hashing, both the off-circuit
 * variable-shape tree (TX/input/tendermint_utils.rs:276-372) and the fixed-shape in-circuit
 * evaluation (PX/frontend/merkle/tendermint.rs:62-214), plus the Blobstream data-commitment
 * gadgets (BX/circuits/builder.rs:82-444).  TEST INFRASTRUCTURE ONLY (see bsx_oracle.h).
 */
#include "bsx_oracle.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* PX/frontend/merkle/tendermint.rs:95-106 ; TX/input/tendermint_utils.rs:358-364 */
void orc_leaf_hash(const uint8_t *leaf, size_t len, uint8_t out[32]) {
    uint8_t buf[256];
    uint8_t *p = len + 1 <= sizeof buf ? buf : (uint8_t *)malloc(len + 1);
    p[0] = 0x00;
    if (len) memcpy(p + 1, leaf, len);
    orc_sha256(p, len + 1, out);
    if (p != buf) free(p);
}

/* PX/frontend/merkle/tendermint.rs:108-122 ; TX/input/tendermint_utils.rs:366-372 */
void orc_inner_hash(const uint8_t l[32], const uint8_t r[32], uint8_t out[32]) {
    uint8_t buf[65];
    buf[0] = 0x01;
    memcpy(buf + 1, l, 32);
    memcpy(buf + 33, r, 32);
    orc_sha256(buf, 65, out);
}

/* TX/input/tendermint_utils.rs:338-349 */
static uint32_t split_point(uint32_t n) {
    uint32_t k = 1;
    while (k * 2 <= n) k *= 2; /* largest power of two <= n */
    return (k == n) ? k >> 1 : k;
}

static void root_rec(const uint8_t *items, const uint32_t *off, uint32_t lo, uint32_t hi, uint8_t out[32]) {
    uint32_t n = hi - lo;
    if (n == 0) { orc_sha256((const uint8_t *)"", 0, out); return; }
    if (n == 1) { orc_leaf_hash(items + off[lo], off[lo + 1] - off[lo], out); return; }
    uint32_t k = split_point(n);
    uint8_t l[32], r[32];
    root_rec(items, off, lo, lo + k, l);
    root_rec(items, off, lo + k, hi, r);
    orc_inner_hash(l, r, out);
}

void orc_tm_root_from_slices(const uint8_t *items, const uint32_t *offsets, uint32_t n, uint8_t out[32]) {
    root_rec(items, offsets, 0, n, out);
}

/* TX/input/tendermint_utils.rs:276-336 (proofs_from_byte_slices / trails_from_byte_slices, flatten_aunts): the aunts of
 * leaf `index` in the variable-shape tree over items [lo, hi), leaf-side first.  Returns the subtree root in `out` and
 * appends to aunts[*depth].  Generic in n -- the product hard-codes the 14-leaf shape, this does not. */
static void aunts_rec(const uint8_t *items, const uint32_t *offsets, uint32_t lo, uint32_t hi, uint32_t index,
                      uint8_t *aunts, uint32_t *depth, uint8_t out[32]) {
    if (hi - lo == 1) {
        orc_leaf_hash(items + offsets[lo], offsets[lo + 1] - offsets[lo], out);
        return;
    }
    uint32_t n = hi - lo, k = 1;
    while (k * 2 < n) k *= 2; /* get_split_point: largest power of two strictly below n (:338-349) */
    uint8_t l[32], r[32];
    if (index < lo + k) {
        aunts_rec(items, offsets, lo, lo + k, index, aunts, depth, l);
        root_rec(items, offsets, lo + k, hi, r);
        memcpy(aunts + 32 * (*depth)++, r, 32);
    } else {
        root_rec(items, offsets, lo, lo + k, l);
        aunts_rec(items, offsets, lo + k, hi, index, aunts, depth, r);
        memcpy(aunts + 32 * (*depth)++, l, 32);
    }
    orc_inner_hash(l, r, out);
}
uint32_t orc_tm_aunts_from_slices(const uint8_t *items, const uint32_t *offsets, uint32_t n, uint32_t index, uint8_t *aunts,
                                  uint8_t root[32]) {
    uint32_t depth = 0;
    aunts_rec(items, offsets, 0, n, index, aunts, &depth, root);
    return depth;
}

/* BX/circuits/input.rs:149-271 + BX/circuits/builder.rs:316-333: the inputs of the n_jobs map circuits of one range from
 * the encoded headers of blocks start .. start + n_jobs*B (records as include/bsx.h BSX_HEADER_LEAVES_BYTES: 14 lengths,
 * then at byte 16 the fields).  The hint of job j is called with (bs, be) = (start + jB, start + (j+1)B) -- NOT with the
 * range's end -- and clamps only to the last block it can fetch, `latest` (= latest_block - 2, input.rs:160-163):
 * req_end = min(be, latest); data_hash proofs (leaf 6) for blocks [bs, req_end), last_block_id proofs (leaf 4) for
 * (bs, req_end]; slots past req_end are zero, and so is a job with bs >= latest (:243-262).  So a job beyond the range's
 * end still carries real proofs whenever the chain has those blocks (its slots are disabled inside the circuit).
 * Returns 0, or 1 if a proven field does not have the circuit's fixed size (the reference errors out). */
int orc_header_range_inputs(uint32_t n_jobs, uint32_t B, const uint8_t *headers, uint64_t start, uint64_t end, uint64_t latest,
                            uint8_t *dh_leaf, uint8_t *dh_aunts, uint8_t *lb_leaf, uint8_t *lb_aunts, uint8_t *start_headers,
                            uint8_t *end_headers, uint8_t start_header[32], uint8_t end_header[32]) {
    const size_t slots = (size_t)n_jobs * B;
    int bad = 0;
    memset(dh_leaf, 0, slots * 34); memset(dh_aunts, 0, slots * 128);
    memset(lb_leaf, 0, slots * 72); memset(lb_aunts, 0, slots * 128);
    memset(start_headers, 0, (size_t)n_jobs * 32); memset(end_headers, 0, (size_t)n_jobs * 32);
    memset(start_header, 0, 32); memset(end_header, 0, 32);
    if (latest > start + slots) latest = start + slots; /* records exist for blocks start .. start + n_jobs*B only */
    for (uint32_t j = 0; j < n_jobs; j++) {
        const uint64_t bs = start + (uint64_t)j * B, be = bs + B;
        if (bs >= latest) continue; /* nothing to fetch: start >= request_end, all zero */
        const uint64_t req_end = be < latest ? be : latest;
        uint32_t k_dh = 0, k_lb = 0;
        for (uint64_t i = bs; i <= req_end; i++) {
            const uint8_t *rec = headers + (size_t)(i - start) * 512;
            uint32_t offs[15];
            offs[0] = 0;
            for (int f = 0; f < 14; f++) offs[f + 1] = offs[f] + rec[f];
            uint8_t aunts[8 * 32], root[32];
            if (i < req_end) {
                uint8_t *slot = dh_leaf + ((size_t)j * B + k_dh) * 34;
                if (rec[6] != 34) bad = 1;
                memcpy(slot, rec + 16 + offs[6], rec[6] < 34 ? rec[6] : 34);
                uint32_t d = orc_tm_aunts_from_slices(rec + 16, offs, 14, 6, aunts, root);
                if (d != 4) bad = 1;
                memcpy(dh_aunts + ((size_t)j * B + k_dh) * 128, aunts, 128);
                k_dh++;
            }
            if (i > bs) {
                uint8_t *slot = lb_leaf + ((size_t)j * B + k_lb) * 72;
                if (rec[4] != 72) bad = 1;
                memcpy(slot, rec + 16 + offs[4], rec[4] < 72 ? rec[4] : 72);
                uint32_t d = orc_tm_aunts_from_slices(rec + 16, offs, 14, 4, aunts, root);
                if (d != 4) bad = 1;
                memcpy(lb_aunts + ((size_t)j * B + k_lb) * 128, aunts, 128);
                k_lb++;
            }
            if (i == bs || i == req_end) {
                orc_tm_root_from_slices(rec + 16, offs, 14, root);
                if (i == bs) memcpy(start_headers + 32 * (size_t)j, root, 32);
                if (i == req_end) memcpy(end_headers + 32 * (size_t)j, root, 32);
            }
        }
    }
    /* the range's own public inputs: header hashes at `start` and `end` (evm inputs of the circuit, not part of the hint) */
    if (end > start && end <= latest) {
        for (int which = 0; which < 2; which++) {
            const uint8_t *rec = headers + (size_t)((which ? end : start) - start) * 512;
            uint32_t offs[15];
            offs[0] = 0;
            for (int f = 0; f < 14; f++) offs[f + 1] = offs[f] + rec[f];
            orc_tm_root_from_slices(rec + 16, offs, 14, which ? end_header : start_header);
        }
    }
    return bad;
}

/* PX/frontend/merkle/tendermint.rs:62-93: leaf hash, then per level BOTH inner(h,aunt) and
 * inner(aunt,h) are requested (left form first) and the path bit selects. */
void orc_tm_merkle_proof(const uint8_t *leaf, uint32_t leaf_len, const uint8_t *aunts, uint32_t depth,
                         uint32_t path_bits, int hashed_leaf, uint8_t *digests, uint8_t root[32]) {
    uint8_t h[32];
    uint8_t *d = digests;
    if (hashed_leaf) {
        memcpy(h, leaf, 32);
    } else {
        orc_leaf_hash(leaf, leaf_len, h);
        memcpy(d, h, 32);
        d += 32;
    }
    for (uint32_t i = 0; i < depth; i++) {
        const uint8_t *aunt = aunts + 32 * i;
        orc_inner_hash(h, aunt, d);      /* left_hash_pair  */
        orc_inner_hash(aunt, h, d + 32); /* right_hash_pair */
        memcpy(h, ((path_bits >> i) & 1) ? d + 32 : d, 32);
        d += 64;
    }
    memcpy(root, h, 32);
}

/* PX/frontend/merkle/tendermint.rs:124-153,165-204 */
void orc_tm_merkle_tree(const uint8_t *leaf_digests, uint32_t N, uint64_t nb_enabled, uint8_t *inner,
                        uint8_t root[32]) {
    uint32_t P = 1;
    while (P < N) P *= 2;
    uint8_t *nodes = (uint8_t *)calloc(P, 32);
    uint8_t *en = (uint8_t *)malloc(P);
    memcpy(nodes, leaf_digests, (size_t)N * 32);
    int is_enabled = 1;
    for (uint32_t i = 0; i < P; i++) { /* :184-194 running AND of (i != nb_enabled) */
        if ((uint64_t)i == nb_enabled) is_enabled = 0;
        en[i] = (uint8_t)is_enabled;
    }
    uint8_t *w = inner;
    for (uint32_t len = P; len > 1; len /= 2) {
        for (uint32_t i = 0; i < len; i += 2) {
            uint8_t ih[32];
            orc_inner_hash(nodes + 32 * (size_t)i, nodes + 32 * (size_t)(i + 1), ih);
            if (w) { memcpy(w, ih, 32); w += 32; }
            int both = en[i] && en[i + 1];
            int none = !en[i] && !en[i + 1];
            /* select(both_enabled, inner, left) ; enabled' = !both_disabled */
            if (both) memcpy(nodes + 32 * (size_t)(i / 2), ih, 32);
            else memmove(nodes + 32 * (size_t)(i / 2), nodes + 32 * (size_t)i, 32);
            en[i / 2] = (uint8_t)!none;
        }
    }
    memcpy(root, nodes, 32);
    free(nodes);
    free(en);
}

/* BX/circuits/builder.rs:82-103 : 24 zero bytes ‖ u64 BE height ‖ data_hash */
void orc_encode_data_root_tuple(const uint8_t data_hash[32], uint64_t height, uint8_t out[64]) {
    memset(out, 0, 24);
    for (int i = 0; i < 8; i++) out[24 + i] = (uint8_t)(height >> (56 - 8 * i));
    memcpy(out + 32, data_hash, 32);
}

/* BX/circuits/builder.rs:105-148 */
void orc_get_data_commitment(const uint8_t *data_hashes, uint32_t B, uint64_t start, uint64_t end,
                             uint8_t *digests, uint8_t root[32], uint32_t *fail) {
    if (end < start && fail) *fail |= ORC_FAIL_END_LT_START;
    uint64_t nb = end - start; /* wrapping, like the u64 gadget */
    if ((nb >> 32) && fail) *fail |= ORC_FAIL_END_LT_START;
    uint64_t nb_enabled = nb & 0xffffffffu; /* limbs[0] */
    for (uint32_t i = 0; i < B; i++) {
        uint8_t tup[64];
        orc_encode_data_root_tuple(data_hashes + 32 * (size_t)i, start + i, tup);
        orc_leaf_hash(tup, 64, digests + 32 * (size_t)i);
    }
    orc_tm_merkle_tree(digests, B, nb_enabled, digests + 32 * (size_t)B, root);
}

static void put64(uint8_t *p, uint64_t v) { for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (8 * i)); }
static uint64_t get64(const uint8_t *p) { uint64_t v = 0; for (int i = 7; i >= 0; i--) v = (v << 8) | p[i]; return v; }
static void put32(uint8_t *p, uint32_t v) { for (int i = 0; i < 4; i++) p[i] = (uint8_t)(v >> (8 * i)); }
static uint32_t get32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

/* BX/circuits/builder.rs:150-271 */
void orc_prove_subchain(uint32_t B, const uint8_t *dh_leaf, const uint8_t *dh_aunts, const uint8_t *lb_leaf,
                        const uint8_t *lb_aunts, const uint8_t start_header[32], const uint8_t end_header[32],
                        uint64_t batch_start, uint64_t batch_end, uint64_t global_end,
                        const uint8_t global_end_header[32], uint8_t *digests, uint8_t *subchain) {
    const uint32_t data_hash_path = 0x6;     /* [0,1,1,0] LSB-first : leaf 6 */
    const uint32_t last_block_id_path = 0x4; /* [0,0,1,0]           : leaf 4 */
    uint32_t fail = 0;
    int is_batch_enabled = batch_start < global_end;
    int curr_enabled = is_batch_enabled;
    uint8_t curr_header[32];
    memcpy(curr_header, start_header, 32);
    uint64_t last_block_to_process = global_end - 1;
    uint8_t *d = digests;
    uint8_t *data_hashes = (uint8_t *)malloc((size_t)B * 32);
    for (uint32_t i = 0; i < B; i++) {
        uint64_t curr_idx = batch_start + i;
        int curr_disabled = !curr_enabled;
        int is_last = (last_block_to_process == curr_idx);
        uint8_t dh_root[32], lb_root[32];
        orc_tm_merkle_proof(dh_leaf + 34 * (size_t)i, 34, dh_aunts + 128 * (size_t)i, 4, data_hash_path, 0, d, dh_root);
        d += 9 * 32;
        orc_tm_merkle_proof(lb_leaf + 72 * (size_t)i, 72, lb_aunts + 128 * (size_t)i, 4, last_block_id_path, 0, d, lb_root);
        d += 9 * 32;
        const uint8_t *header_hash = lb_leaf + 72 * (size_t)i + 2;
        if (!(curr_disabled || memcmp(curr_header, header_hash, 32) == 0)) fail |= ORC_FAIL_PREV_HEADER;
        if (!(curr_disabled || memcmp(dh_root, header_hash, 32) == 0)) fail |= ORC_FAIL_DATA_HASH;
        if (!(!is_last || memcmp(lb_root, global_end_header, 32) == 0)) fail |= ORC_FAIL_END_HEADER;
        if (curr_enabled) memcpy(curr_header, lb_root, 32);
        curr_enabled = curr_enabled && !is_last;
        memcpy(data_hashes + 32 * (size_t)i, dh_leaf + 34 * (size_t)i + 2, 32);
    }
    if (!(!curr_enabled || memcmp(curr_header, end_header, 32) == 0)) fail |= ORC_FAIL_BATCH_END_HEADER;
    uint64_t temp_end = (batch_end < global_end) ? batch_end : global_end;
    uint64_t end_block_num = (temp_end < batch_start) ? batch_start : temp_end;
    uint8_t root[32];
    orc_get_data_commitment(data_hashes, B, batch_start, end_block_num, d, root, &fail);
    free(data_hashes);
    memset(subchain, 0, ORC_SUBCHAIN_BYTES);
    subchain[0] = (uint8_t)is_batch_enabled;
    put32(subchain + 4, fail);
    put64(subchain + 8, batch_start);
    put64(subchain + 16, end_block_num);
    memcpy(subchain + 24, start_header, 32);
    memcpy(subchain + 56, curr_header, 32);
    memcpy(subchain + 88, root, 32);
}

/* BX/circuits/builder.rs:337-395 ; the hash is the bit-level sha256 gadget (:363-364) */
void orc_reduce_subchain(const uint8_t *left, const uint8_t *right, uint8_t *out, uint8_t digest[32]) {
    uint32_t fail = get32(left + 4) | get32(right + 4);
    int right_disabled = (right[0] == 0);
    int headers_linked = memcmp(left + 56, right + 24, 32) == 0;
    int blocks_linked = get64(left + 16) == get64(right + 8);
    if (!(right_disabled || (headers_linked && blocks_linked))) fail |= ORC_FAIL_REDUCE_LINK;
    orc_inner_hash(left + 88, right + 88, digest);
    memset(out, 0, ORC_SUBCHAIN_BYTES);
    out[0] = left[0];
    put32(out + 4, fail);
    memcpy(out + 8, left + 8, 8);                               /* start_block */
    memcpy(out + 16, right_disabled ? left + 16 : right + 16, 8); /* end_block */
    memcpy(out + 24, left + 24, 32);                            /* start_header */
    memcpy(out + 56, right_disabled ? left + 56 : right + 56, 32);
    memcpy(out + 88, right_disabled ? left + 88 : digest, 32);
}

/* BX/circuits/builder.rs:273-409 with PX/frontend/mapreduce/generator.rs:86-151 (map all jobs,
 * then reduce layer by layer). */
void orc_prove_data_commitment(uint32_t n_jobs, uint32_t B, const uint8_t *dh_leaf, const uint8_t *dh_aunts,
                               const uint8_t *lb_leaf, const uint8_t *lb_aunts, const uint8_t *start_headers,
                               const uint8_t *end_headers, uint64_t start_block, const uint8_t start_header[32],
                               uint64_t end_block, const uint8_t end_header[32], uint8_t *map_digests,
                               uint8_t *map_subchains, uint8_t *reduce_digests, uint8_t *reduce_nodes,
                               uint8_t data_commitment[32], uint32_t *fail, int threads) {
    uint32_t f = 0;
    size_t dig_per_job = (size_t)(20 * B - 1) * 32;
    (void)threads;
#pragma omp parallel for schedule(dynamic) num_threads(threads > 0 ? threads : 1)
    for (uint32_t j = 0; j < n_jobs; j++) {
        uint64_t bs = start_block + (uint64_t)j * B;
        uint64_t be = bs + B; /* last_block + 1 */
        orc_prove_subchain(B, dh_leaf + (size_t)j * B * 34, dh_aunts + (size_t)j * B * 128,
                           lb_leaf + (size_t)j * B * 72, lb_aunts + (size_t)j * B * 128,
                           start_headers + 32 * (size_t)j, end_headers + 32 * (size_t)j, bs, be, end_block,
                           end_header, map_digests + dig_per_job * j, map_subchains + ORC_SUBCHAIN_BYTES * (size_t)j);
    }
    uint32_t fr = 0;
    orc_reduce_subchains(n_jobs, B, map_subchains, start_block, start_header, end_block, end_header, reduce_digests,
                         reduce_nodes, data_commitment, &fr);
    if (fail) *fail = f | fr;
}

/* the reduce tree alone (BX/circuits/builder.rs:337-409) over already-computed map outputs */
void orc_reduce_subchains(uint32_t n_jobs, uint32_t B, const uint8_t *map_subchains, uint64_t start_block,
                          const uint8_t start_header[32], uint64_t end_block, const uint8_t end_header[32],
                          uint8_t *reduce_digests, uint8_t *reduce_nodes, uint8_t data_commitment[32], uint32_t *fail) {
    uint32_t f = 0;
    if (!(end_block <= start_block + (uint64_t)n_jobs * B)) f |= ORC_FAIL_RANGE;
    uint8_t *cur = (uint8_t *)malloc((size_t)n_jobs * ORC_SUBCHAIN_BYTES);
    memcpy(cur, map_subchains, (size_t)n_jobs * ORC_SUBCHAIN_BYTES);
    uint8_t *rd = reduce_digests, *rn = reduce_nodes;
    for (uint32_t len = n_jobs; len > 1; len /= 2) {
        for (uint32_t i = 0; i < len; i += 2) {
            uint8_t node[ORC_SUBCHAIN_BYTES], dg[32];
            orc_reduce_subchain(cur + ORC_SUBCHAIN_BYTES * (size_t)i, cur + ORC_SUBCHAIN_BYTES * (size_t)(i + 1), node, dg);
            memcpy(cur + ORC_SUBCHAIN_BYTES * (size_t)(i / 2), node, ORC_SUBCHAIN_BYTES);
            if (rd) { memcpy(rd, dg, 32); rd += 32; }
            if (rn) { memcpy(rn, node, ORC_SUBCHAIN_BYTES); rn += ORC_SUBCHAIN_BYTES; }
        }
    }
    f |= get32(cur + 4);
    /* :398-406 result must match the public inputs */
    if (get64(cur + 8) != start_block || memcmp(cur + 24, start_header, 32) != 0 ||
        get64(cur + 16) != end_block || memcmp(cur + 56, end_header, 32) != 0)
        f |= ORC_FAIL_RESULT;
    memcpy(data_commitment, cur + 88, 32);
    free(cur);
    if (fail) *fail = f;
}

/* TX/builder/shared.rs:67-156 : 9 bytes always; byte i = septet i, MSB set iff i < index of the
 * last non-zero septet. */
uint32_t orc_marshal_int64_varint(uint64_t v, uint8_t out[9]) {
    uint32_t last = 0;
    uint8_t sept[9];
    for (int i = 0; i < 9; i++) {
        sept[i] = (uint8_t)((v >> (7 * i)) & 0x7f);
        if (sept[i]) last = (uint32_t)i;
    }
    for (uint32_t i = 0; i < 9; i++) out[i] = (uint8_t)(sept[i] | ((i < last) ? 0x80 : 0));
    return last + 1;
}

/* TX/builder/validator.rs:185-207 : 0a 22 0a 20 ‖ pk ‖ 10 ‖ varint[9] = 46 bytes */
uint32_t orc_marshal_validator(const uint8_t pubkey[32], uint64_t power, uint8_t out[46]) {
    out[0] = 10; out[1] = 34; out[2] = 10; out[3] = 32;
    memcpy(out + 4, pubkey, 32);
    out[36] = 16;
    uint32_t n = orc_marshal_int64_varint(power, out + 37);
    return 37 + n;
}

/* TX/builder/validator.rs:209-252 */
void orc_hash_validator_set(uint32_t N, const uint8_t *pubkeys, const uint64_t *powers,
                            const uint32_t *byte_lengths, uint64_t nb_enabled, uint8_t *digests,
                            uint8_t root[32]) {
    for (uint32_t i = 0; i < N; i++) {
        uint8_t buf[64];
        memset(buf, 0, sizeof buf);
        buf[0] = 0x00;
        orc_marshal_validator(pubkeys + 32 * (size_t)i, powers[i], buf + 1);
        /* curta_sha256_variable over the 64-byte buffer: the hint hashes the first len bytes
         * (PX/frontend/hash/curta/digest_hint.rs:31-33) */
        orc_sha256(buf, 1 + byte_lengths[i], digests + 32 * (size_t)i);
    }
    orc_tm_merkle_tree(digests, N, nb_enabled, digests + 32 * (size_t)N, root);
}

/* ------------------------------------------------------------------------------------------------
 * Off-chain input shaping: protobuf encoders, sign-bytes, validator records
 * ------------------------------------------------------------------------------------------------ */
typedef struct { uint8_t *p; uint32_t n; } pbw;
static void pb_byte(pbw *w, uint8_t b) { w->p[w->n++] = b; }
static void pb_uvarint(pbw *w, uint64_t v) {
    while (v >= 0x80) { pb_byte(w, (uint8_t)(v | 0x80)); v >>= 7; }
    pb_byte(w, (uint8_t)v);
}
static void pb_varint_field(pbw *w, uint8_t tag, uint64_t v) {      /* proto3: default value omitted */
    if (!v) return;
    pb_byte(w, tag); pb_uvarint(w, v);
}
static void pb_bytes_field(pbw *w, uint8_t tag, const uint8_t *s, uint32_t len) {
    if (!len) return;
    pb_byte(w, tag); pb_uvarint(w, len);
    memcpy(w->p + w->n, s, len); w->n += len;
}
/* BlockId {hash = 1, part_set_header = 2 {total = 1, hash = 2}}; CanonicalBlockId uses the same numbers */
static uint32_t enc_block_id(const uint8_t hash[32], uint32_t parts_total, const uint8_t parts_hash[32], uint8_t *out) {
    uint8_t psh[48];
    pbw p = {psh, 0};
    pb_varint_field(&p, 0x08, parts_total);
    pb_bytes_field(&p, 0x12, parts_hash, 32);
    pbw w = {out, 0};
    pb_bytes_field(&w, 0x0A, hash, 32);
    pb_bytes_field(&w, 0x12, psh, p.n);
    return w.n;
}
static uint32_t enc_timestamp(int64_t secs, uint32_t nanos, uint8_t *out) {
    pbw w = {out, 0};
    pb_varint_field(&w, 0x08, (uint64_t)secs);
    pb_varint_field(&w, 0x10, nanos);
    return w.n;
}

/* the 14 encode_vec calls of generate_proofs_from_header (TX/input/tendermint_utils.rs:374-393); fields back to back
 * into out (<= 496 bytes), their lengths into lens; hashes = last_commit_hash, data_hash, validators_hash,
 * next_validators_hash, consensus_hash, app_hash, last_results_hash, evidence_hash, proposer_address */
uint32_t orc_encode_header_fields(uint64_t version_block, uint64_t version_app, const uint8_t *chain_id, uint32_t chain_id_len,
                                  uint64_t height, int64_t time_secs, uint32_t time_nanos, int has_last_block_id,
                                  const uint8_t last_block_hash[32], uint32_t parts_total, const uint8_t parts_hash[32],
                                  const uint8_t *hashes, const uint8_t hash_len[9], uint8_t lens[14], uint8_t *out) {
    pbw w = {out, 0};
    uint32_t at = 0;
#define ORC_CLOSE(f) do { lens[f] = (uint8_t)(w.n - at); at = w.n; } while (0)
    pb_varint_field(&w, 0x08, version_block);
    pb_varint_field(&w, 0x10, version_app);
    ORC_CLOSE(0);
    pb_bytes_field(&w, 0x0A, chain_id, chain_id_len);
    ORC_CLOSE(1);
    pb_varint_field(&w, 0x08, height);
    ORC_CLOSE(2);
    w.n += enc_timestamp(time_secs, time_nanos, w.p + w.n);
    ORC_CLOSE(3);
    if (has_last_block_id) w.n += enc_block_id(last_block_hash, parts_total, parts_hash, w.p + w.n);
    ORC_CLOSE(4);
    for (int k = 0; k < 9; k++) {
        pb_bytes_field(&w, 0x0A, hashes + 32 * k, hash_len[k]);
        ORC_CLOSE(5 + k);
    }
#undef ORC_CLOSE
    return w.n;
}

/* SignedVote::sign_bytes of a precommit (TX/input/conversion.rs:34-39): length-delimited CanonicalVote
 * {type = 1, height = 2 sfixed64, round = 3 sfixed64, block_id = 4, timestamp = 5, chain_id = 6}; out >= 192 bytes */
uint32_t orc_vote_sign_bytes(const uint8_t *chain_id, uint32_t chain_id_len, uint64_t height, uint64_t round, int has_block_id,
                             const uint8_t block_hash[32], uint32_t parts_total, const uint8_t parts_hash[32], int64_t ts_secs,
                             uint32_t ts_nanos, uint8_t *out) {
    uint8_t body[192], tmp[96];
    pbw b = {body, 0};
    pb_varint_field(&b, 0x08, 2);
    if (height) { pb_byte(&b, 0x11); put64(b.p + b.n, height); b.n += 8; }
    if (round) { pb_byte(&b, 0x19); put64(b.p + b.n, round); b.n += 8; }
    if (has_block_id) pb_bytes_field(&b, 0x22, tmp, enc_block_id(block_hash, parts_total, parts_hash, tmp));
    uint32_t tl = enc_timestamp(ts_secs, ts_nanos, tmp);
    pb_byte(&b, 0x2A); pb_uvarint(&b, tl);            /* Some(timestamp): written even when empty */
    memcpy(b.p + b.n, tmp, tl); b.n += tl;
    pb_bytes_field(&b, 0x32, chain_id, chain_id_len);
    pbw w = {out, 0};
    pb_uvarint(&w, b.n);
    memcpy(w.p + w.n, body, b.n);
    return w.n + b.n;
}

static const uint8_t ORC_DUMMY_PK[32] = {138, 136, 227, 221, 116, 9, 241, 149, 253, 82, 219, 45, 60, 186, 93, 114,
                                         202, 103, 9, 191, 29, 148, 18, 27, 243, 116, 136, 1, 180, 15, 111, 92};
static const uint8_t ORC_DUMMY_SIG[64] = {55, 20, 104, 158, 84, 120, 194, 17, 6, 237, 157, 164, 85, 88, 158, 137,
                                          187, 119, 187, 240, 159, 73, 80, 63, 133, 162, 74, 91, 48, 53, 6, 138,
                                          1, 41, 22, 121, 249, 46, 198, 145, 155, 102, 3, 210, 168, 135, 173, 55,
                                          252, 72, 45, 126, 169, 178, 191, 7, 153, 67, 112, 90, 150, 33, 140, 7};

/* get_validator_data_from_block (TX/input/conversion.rs:59-140) + validator_hash_field_from_block (:142-184) for one
 * commit.  records: N * 240 bytes in the layout of include/bsx.h (ValidatorVariable); hf_*: N entries each.
 * Either output group may be NULL.  Returns non-zero when a message exceeds 124 bytes or n_sigs > N (panics there). */
int orc_validator_records(uint32_t N, uint32_t n_sigs, const uint8_t *chain_id, uint32_t chain_id_len, uint64_t height, uint64_t round,
                          int has_block_id, const uint8_t block_hash[32], uint32_t parts_total, const uint8_t parts_hash[32],
                          const uint8_t *pubkeys, const uint8_t *signatures, const uint64_t *powers, const int64_t *ts_secs,
                          const uint32_t *ts_nanos, const uint8_t *flags, uint8_t *records, uint8_t *hf_pubkeys, uint64_t *hf_powers,
                          uint32_t *hf_lens) {
    int bad = n_sigs > N;
    for (uint32_t i = 0; i < N; i++) {
        const int in_set = i < n_sigs, is_commit = in_set && flags[i] == 2;
        const uint8_t *pk = in_set ? pubkeys + 32 * (size_t)i : ORC_DUMMY_PK;
        const uint64_t power = in_set ? powers[i] : 0;
        uint8_t vb[48];
        uint32_t vlen = 46;                                       /* VALIDATOR_BYTE_LENGTH_MAX */
        if (in_set) {                                             /* validator.hash_bytes(): 0a 22 0a 20 pk 10 varint */
            pbw v = {vb, 0};
            pb_byte(&v, 0x0A); pb_byte(&v, 0x22); pb_byte(&v, 0x0A); pb_byte(&v, 0x20);
            memcpy(v.p + v.n, pk, 32); v.n += 32;
            pb_byte(&v, 0x10); pb_uvarint(&v, power);
            vlen = v.n;
        }
        if (hf_pubkeys) {
            memcpy(hf_pubkeys + 32 * (size_t)i, pk, 32);
            hf_powers[i] = power;
            hf_lens[i] = vlen;
        }
        if (!records) continue;
        uint8_t *r = records + 240 * (size_t)i;
        memset(r, 0, 240);
        memcpy(r, pk, 32);
        uint32_t msg_len = 32;
        if (is_commit) {
            uint8_t msg[200];
            memcpy(r + 32, signatures + 64 * (size_t)i, 64);
            msg_len = orc_vote_sign_bytes(chain_id, chain_id_len, height, round, has_block_id, block_hash, parts_total, parts_hash,
                                          ts_secs[i], ts_nanos[i], msg);
            if (msg_len > 124) { bad = 1; msg_len = 0; }
            memcpy(r + 96, msg, msg_len);
        } else {
            memcpy(r + 32, ORC_DUMMY_SIG, 64);
        }
        put32(r + 220, msg_len);
        put64(r + 224, power);
        put32(r + 232, vlen);
        r[236] = (uint8_t)is_commit;
    }
    return bad;
}

/* update_present_on_trusted_header (TX/input/conversion.rs:186-240) for one commit; addresses are 20 bytes.
 * Sets records[i*240 + 237]; returns non-zero when a third of the target power is not reached (asserts there). */
int orc_present_on_trusted(uint32_t n_target, const uint8_t *tg_addr, const uint8_t *tg_sig_addr, const uint8_t *tg_flags,
                           const uint64_t *tg_powers, uint32_t n_trusted, const uint8_t *tr_addr, uint8_t *records) {
    uint64_t total = 0, shared = 0;
    for (uint32_t i = 0; i < n_target; i++) total += tg_powers[i];
    const double threshold = 1.0 / 3.0;
    uint32_t s = 0;
    while ((double)total * threshold > (double)shared && s < n_trusted) {
        const uint8_t *a = tr_addr + 20 * (size_t)s;
        for (uint32_t i = 0; i < n_target; i++) {
            if (memcmp(tg_addr + 20 * (size_t)i, a, 20)) continue;
            for (uint32_t j = 0; j < n_target; j++) {             /* sig.validator_address(): Some for commit and nil votes */
                if ((tg_flags[j] == 2 || tg_flags[j] == 3) && !memcmp(tg_sig_addr + 20 * (size_t)j, a, 20)) {
                    shared += tg_powers[i];
                    records[240 * (size_t)i + 237] = 1;
                }
            }
            break;
        }
        s++;
    }
    return (double)total * threshold > (double)shared;
}

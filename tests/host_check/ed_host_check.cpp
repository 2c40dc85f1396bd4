// TEST INFRASTRUCTURE: compiles the kernel-side Ed25519 arithmetic (blobstreamx_b200/csrc/ed25519.cuh)
// for the HOST with plain g++, so its logic can be checked against the oracle on a machine without a
// GPU.  Never linked into libbsx.so; the product has no CPU path.
#include "../../blobstreamx_b200/csrc/ed25519.cuh"
#include "../../blobstreamx_b200/csrc/ed_trace.cuh"

#include <fenv.h>
#include <stdlib.h>
#include <string.h>

using namespace bsx::ed;

static ge_niels_slot *g_table = nullptr;

static void build_table() {
    if (g_table) return;
    g_table = (ge_niels_slot *)aligned_alloc(16, sizeof(ge_niels_slot) * BSX_ED_BASE_WINDOWS * BSX_ED_BASE_ENTRIES);
    for (int w = 0; w < BSX_ED_BASE_WINDOWS; w++)
        for (int d = 1; d <= BSX_ED_BASE_ENTRIES; d++) ge_niels_store(g_table + w * BSX_ED_BASE_ENTRIES + d - 1, ge_base_table_entry(w, d));
}

extern "C" {
void hc_ed25519_witness(const uint8_t *pk, const uint8_t *sig, const uint8_t *digest, uint8_t *out) {
    build_table();
    ed25519_witness_core(pk, sig, digest, g_table, out);
}
void hc_fe_mul(const uint8_t *a, const uint8_t *b, uint8_t *out) { fe_tobytes(out, fe_mul(fe_frombytes(a), fe_frombytes(b))); }
void hc_fe_sq(const uint8_t *a, uint8_t *out, int twice) {
    fe f = fe_frombytes(a);
    fe_tobytes(out, twice ? fe_sq2(f) : fe_sq(f));
}
// raw-limb entry points: operands at the extreme limb bounds the point formulas produce
void hc_fe_mul_limbs(const int32_t *f, const int32_t *g, uint8_t *out) {
    fe a, b;
    for (int i = 0; i < 10; i++) { a.v[i] = f[i]; b.v[i] = g[i]; }
    fe_tobytes(out, fe_mul(a, b));
}
void hc_fe_sq_limbs(const int32_t *f, uint8_t *out, int twice) {
    fe a;
    for (int i = 0; i < 10; i++) a.v[i] = f[i];
    fe_tobytes(out, twice ? fe_sq2(a) : fe_sq(a));
}
void hc_fe_tighten_limbs(const int32_t *f, int32_t *out) {
    fe a;
    for (int i = 0; i < 10; i++) a.v[i] = f[i];
    a = fe_tighten(a);
    for (int i = 0; i < 10; i++) out[i] = a.v[i];
}
void hc_fe_invert(const uint8_t *a, uint8_t *out) { fe_tobytes(out, fe_invert(fe_frombytes(a))); }
void hc_fe_addsub_mul(const uint8_t *a, const uint8_t *b, uint8_t *out) {
    // worst-case limb growth the point formulas rely on: (a+b)*(a-b)
    fe x = fe_frombytes(a), y = fe_frombytes(b);
    fe_tobytes(out, fe_mul(fe_add(x, y), fe_sub(x, y)));
}
void hc_divrem_l(const uint8_t *digest, uint8_t *rem, uint8_t *div) { sc_divrem_l(digest, rem, div); }
void hc_scalarmult(const uint8_t *s, const uint8_t *xy, uint8_t *out, int base) {
    build_table();
    ge_p3 r = base ? ge_scalarmult_base(s, g_table) : ge_scalarmult(s, ge_from_affine(fe_frombytes(xy), fe_frombytes(xy + 32)));
    fe zi = fe_invert(r.Z);
    fe_tobytes(out, fe_mul(r.X, zi));
    fe_tobytes(out + 32, fe_mul(r.Y, zi));
}
int hc_decompress(const uint8_t *in, uint8_t *xy, uint8_t *root) {
    fe x, y;
    return ge_decompress(in, x, y, xy, xy + 32, root) ? 1 : 0;
}

// ---- the FP64-pipe field (fe51d.cuh): the host stands in for __fma_rz with fma() under round-toward-zero ----
struct RoundTowardZero {
    int old;
    RoundTowardZero() : old(fegetround()) { fesetround(FE_TOWARDZERO); }
    ~RoundTowardZero() { fesetround(old); }
};
void hc_fp64_ed25519_witness(const uint8_t *pk, const uint8_t *sig, const uint8_t *digest, uint8_t *out) {
    build_table();
    RoundTowardZero rz;
    bsx::edd::ed25519_witness_core(pk, sig, digest, g_table, out);
}
// raw limbs (integer-valued doubles) -> product limbs, for the bound tests
void hc_fed_mul(const double *f, const double *g, double *o, int twice) {
    RoundTowardZero rz;
    bsx::edd::fed a, b;
    for (int i = 0; i < 5; i++) { a.v[i] = f[i]; b.v[i] = g[i]; }
    const bsx::edd::fed r = twice ? bsx::edd::fed_mul2(a, b) : bsx::edd::fed_mul(a, b);
    for (int i = 0; i < 5; i++) o[i] = r.v[i];
}
void hc_fed_sq(const double *f, double *o, int twice) {
    RoundTowardZero rz;
    bsx::edd::fed a;
    for (int i = 0; i < 5; i++) a.v[i] = f[i];
    const bsx::edd::fed r = twice ? bsx::edd::fed_sq2(a) : bsx::edd::fed_sq(a);
    for (int i = 0; i < 5; i++) o[i] = r.v[i];
}
// bytes -> integer limbs -> double limbs -> integer limbs -> bytes
void hc_fed_roundtrip(const uint8_t *a, uint8_t *out) {
    RoundTowardZero rz;
    fe_tobytes(out, bsx::edd::fe_from_fed(bsx::edd::fed_from_fe(fe_frombytes(a))));
}
void hc_fed_invert(const uint8_t *a, uint8_t *out) {
    RoundTowardZero rz;
    fe_tobytes(out, bsx::edd::fe_from_fed(bsx::edd::fed_invert(bsx::edd::fed_from_fe(fe_frombytes(a)))));
}
void hc_fp64_scalarmult(const uint8_t *s, const uint8_t *xy, uint8_t *out, int base) {
    build_table();
    RoundTowardZero rz;
    using namespace bsx::edd;
    ged_p3 r = base ? ged_scalarmult_base(s, g_table)
                    : ged_scalarmult(s, ged_from_affine(fed_from_fe(fe_frombytes(xy)), fed_from_fe(fe_frombytes(xy + 32))));
    fed zi = fed_invert(r.Z);
    fe_tobytes(out, fe_from_fed(fed_mul(r.X, zi)));
    fe_tobytes(out + 32, fe_from_fed(fed_mul(r.Y, zi)));
}
// the per-key table path: bases and window tables of pk built here, then the keyed record
void hc_fp64_ed25519_witness_keyed(const uint8_t *pk, const uint8_t *sig, const uint8_t *digest, uint8_t *out) {
    build_table();
    RoundTowardZero rz;
    static uint8_t rec[BSX_ED_KEYREC_BYTES];
    static double bases[BSX_ED_KEY_WINDOWS * 20], tab[BSX_ED_KEY_WINDOWS * BSX_ED_KEY_ENTRIES * 20];
    bsx::edd::ed25519_key_bases(pk, rec, bases);
    for (int w = 0; w < BSX_ED_KEY_WINDOWS; w++) bsx::edd::ed25519_key_window(bases + 20 * w, tab + 20 * BSX_ED_KEY_ENTRIES * w);
    bsx::edd::ed25519_witness_core_keyed(sig, digest, g_table, rec, tab, out);
}
int hc_fp64_decompress(const uint8_t *in, uint8_t *xy, uint8_t *root) {
    RoundTowardZero rz;
    bsx::edd::fed x, y;
    return bsx::edd::ged_decompress(in, x, y, xy, xy + 32, root) ? 1 : 0;
}
// the Ed25519 scalar-multiplication trace (ed_trace.cuh): chain + every row, exactly as the two kernels call the cores
void hc_ed25519_trace(const uint8_t *scalars, const uint8_t *points, uint32_t n_muls, uint32_t log_rows, uint8_t *results, uint64_t *trace) {
    using namespace bsx::edt;
    const size_t n_rows = (size_t)1 << log_rows;
    int32_t *chain = (int32_t *)malloc((size_t)(n_muls + 1) * 256 * EDT_CHAIN_WORDS * 4);
    uint32_t *aff = (uint32_t *)malloc((size_t)(n_muls + 1) * 256 * EDT_AFF_WORDS * 4);
    for (uint32_t m = 0; m < n_muls; m++)
        edt_forward_core<false>(scalars + (size_t)m * 32, points + (size_t)m * 64, chain + (size_t)m * 256 * EDT_CHAIN_WORDS);
    for (uint32_t g = 0; g < n_muls * (256 / EDT_GROUP); g++)
        edt_affine_core(chain + (size_t)g * EDT_GROUP * EDT_CHAIN_WORDS, aff + (size_t)g * EDT_GROUP * EDT_AFF_WORDS);
    for (size_t row = 0; row < n_rows; row++) {
        const uint32_t m = (uint32_t)(row >> 8), j = (uint32_t)row & 255;
        const bool real = m < n_muls;
        const EdtSink sink{trace + row, n_rows, trace + row, n_rows};
        edt_row_core(real, j, real ? scalars + (size_t)m * 32 : nullptr, real ? points + (size_t)m * 64 : nullptr,
                     real ? aff + (size_t)m * 256 * EDT_AFF_WORDS : nullptr, sink, real && results ? results + (size_t)m * 64 : nullptr);
    }
    free(chain); free(aff);
}
// one witnessed field operation of the trace (ed_trace.cuh edt_op_rt) on arbitrary canonical operands: 92 values out
// kind 0 mul a1 b1, 1 inner a1 b1 + a2 b2, 2 / 3 division with the given result res (16-bit limbs, 16 each)
void hc_edt_op(int kind, const uint32_t *a1, const uint32_t *b1, const uint32_t *a2, const uint32_t *b2, uint32_t *res, uint64_t *out92) {
    bsx::edt::edt_op_rt(kind, a1, b1, a2, b2, res, out92, 1, out92, 1);
}
}


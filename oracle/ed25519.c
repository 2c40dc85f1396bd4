/*
 * oracle/ed25519.c -- Ed25519 witness values as the plonky2x EC hints produce them:
 *   EcOpResultHint::{Add, ScalarMul, Decompress}  PX/frontend/ecc/curve25519/curta/result_hint.rs:21-50
 *   verification schedule                        PX/frontend/ecc/curve25519/ed25519/eddsa.rs:161-203
 *   h = LE512(SHA512(R‖A‖M)) div/rem l           PX/frontend/uint/num/biguint/mod.rs:38-62,481-488
 * The arithmetic itself lives in starkyx@ad8eb4ba (`AffinePoint` add / scalar-mul, `decompress`),
 * which is NOT vendored.  Affine results in canonical form are mathematically unique, so any
 * correct implementation is bit-exact; `decompress` returns (point, root) with root = the EVEN
 * square root of (y^2-1)/(d y^2+1) and x = root (sign 0) or p-root (sign 1), as pinned by
 * audits/Curta_Plonky2x_Audit_Report_KALOS.md:1089-1132 (prose only; no in-tree KAT prints root).
 * Field elements: 5 x 51-bit limbs with unsigned __int128 products (deliberately a different
 * representation from the CUDA kernel's 10 x 25.5).  TEST INFRASTRUCTURE ONLY.
 */
#include "bsx_oracle.h"
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t v[5]; } fe;
#define M51 ((1ULL << 51) - 1)

static void fe_0(fe *r) { memset(r, 0, sizeof *r); }
static void fe_1(fe *r) { fe_0(r); r->v[0] = 1; }

static void fe_frombytes(fe *r, const uint8_t s[32]) {
    uint64_t w[4];
    for (int i = 0; i < 4; i++) { w[i] = 0; for (int j = 7; j >= 0; j--) w[i] = (w[i] << 8) | s[8 * i + j]; }
    r->v[0] = w[0] & M51;
    r->v[1] = ((w[0] >> 51) | (w[1] << 13)) & M51;
    r->v[2] = ((w[1] >> 38) | (w[2] << 26)) & M51;
    r->v[3] = ((w[2] >> 25) | (w[3] << 39)) & M51;
    r->v[4] = (w[3] >> 12) & M51; /* drops bit 255 */
}

static void fe_carry(fe *r) {
    uint64_t c;
    for (int k = 0; k < 2; k++) {
        c = r->v[0] >> 51; r->v[0] &= M51; r->v[1] += c;
        c = r->v[1] >> 51; r->v[1] &= M51; r->v[2] += c;
        c = r->v[2] >> 51; r->v[2] &= M51; r->v[3] += c;
        c = r->v[3] >> 51; r->v[3] &= M51; r->v[4] += c;
        c = r->v[4] >> 51; r->v[4] &= M51; r->v[0] += 19 * c;
    }
}

/* canonical little-endian bytes in [0,p) */
static void fe_tobytes(uint8_t s[32], const fe *a) {
    fe t = *a;
    fe_carry(&t);
    /* now t < 2^255 + small; subtract p if t >= p */
    uint64_t q = (t.v[0] + 19) >> 51;
    q = (t.v[1] + q) >> 51; q = (t.v[2] + q) >> 51; q = (t.v[3] + q) >> 51; q = (t.v[4] + q) >> 51;
    t.v[0] += 19 * q;
    uint64_t c;
    c = t.v[0] >> 51; t.v[0] &= M51; t.v[1] += c;
    c = t.v[1] >> 51; t.v[1] &= M51; t.v[2] += c;
    c = t.v[2] >> 51; t.v[2] &= M51; t.v[3] += c;
    c = t.v[3] >> 51; t.v[3] &= M51; t.v[4] += c;
    t.v[4] &= M51;
    uint64_t w[4];
    w[0] = t.v[0] | (t.v[1] << 51);
    w[1] = (t.v[1] >> 13) | (t.v[2] << 38);
    w[2] = (t.v[2] >> 26) | (t.v[3] << 25);
    w[3] = (t.v[3] >> 39) | (t.v[4] << 12);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) s[8 * i + j] = (uint8_t)(w[i] >> (8 * j));
}

static void fe_add(fe *r, const fe *a, const fe *b) { for (int i = 0; i < 5; i++) r->v[i] = a->v[i] + b->v[i]; }
/* a - b with a bias of 2p so limbs stay non-negative (inputs carried: limbs < 2^52) */
static void fe_sub(fe *r, const fe *a, const fe *b) {
    r->v[0] = a->v[0] + 0xFFFFFFFFFFFDAULL - b->v[0];
    for (int i = 1; i < 5; i++) r->v[i] = a->v[i] + 0xFFFFFFFFFFFFEULL - b->v[i];
    fe_carry(r);
}
static void fe_neg(fe *r, const fe *a) { fe z; fe_0(&z); fe_sub(r, &z, a); }

static void fe_mul(fe *r, const fe *a, const fe *b) {
    u128 t[5];
    uint64_t a0 = a->v[0], a1 = a->v[1], a2 = a->v[2], a3 = a->v[3], a4 = a->v[4];
    uint64_t b0 = b->v[0], b1 = b->v[1], b2 = b->v[2], b3 = b->v[3], b4 = b->v[4];
    uint64_t b1_19 = 19 * b1, b2_19 = 19 * b2, b3_19 = 19 * b3, b4_19 = 19 * b4;
    t[0] = (u128)a0 * b0 + (u128)a1 * b4_19 + (u128)a2 * b3_19 + (u128)a3 * b2_19 + (u128)a4 * b1_19;
    t[1] = (u128)a0 * b1 + (u128)a1 * b0 + (u128)a2 * b4_19 + (u128)a3 * b3_19 + (u128)a4 * b2_19;
    t[2] = (u128)a0 * b2 + (u128)a1 * b1 + (u128)a2 * b0 + (u128)a3 * b4_19 + (u128)a4 * b3_19;
    t[3] = (u128)a0 * b3 + (u128)a1 * b2 + (u128)a2 * b1 + (u128)a3 * b0 + (u128)a4 * b4_19;
    t[4] = (u128)a0 * b4 + (u128)a1 * b3 + (u128)a2 * b2 + (u128)a3 * b1 + (u128)a4 * b0;
    uint64_t c;
    uint64_t r0, r1, r2, r3, r4;
    r0 = (uint64_t)t[0] & M51; c = (uint64_t)(t[0] >> 51);
    t[1] += c; r1 = (uint64_t)t[1] & M51; c = (uint64_t)(t[1] >> 51);
    t[2] += c; r2 = (uint64_t)t[2] & M51; c = (uint64_t)(t[2] >> 51);
    t[3] += c; r3 = (uint64_t)t[3] & M51; c = (uint64_t)(t[3] >> 51);
    t[4] += c; r4 = (uint64_t)t[4] & M51; c = (uint64_t)(t[4] >> 51);
    r0 += c * 19; c = r0 >> 51; r0 &= M51; r1 += c;
    r->v[0] = r0; r->v[1] = r1; r->v[2] = r2; r->v[3] = r3; r->v[4] = r4;
}
static void fe_sq(fe *r, const fe *a) { fe_mul(r, a, a); }

static int fe_iszero(const fe *a) { uint8_t s[32]; fe_tobytes(s, a); uint8_t x = 0; for (int i = 0; i < 32; i++) x |= s[i]; return x == 0; }
static int fe_eq(const fe *a, const fe *b) { uint8_t s[32], t[32]; fe_tobytes(s, a); fe_tobytes(t, b); return memcmp(s, t, 32) == 0; }
static int fe_isodd(const fe *a) { uint8_t s[32]; fe_tobytes(s, a); return s[0] & 1; }

/* generic a^e, e given as 32 little-endian bytes (square-and-multiply, MSB first) */
static void fe_pow(fe *r, const fe *a, const uint8_t e[32]) {
    fe acc; fe_1(&acc);
    for (int i = 255; i >= 0; i--) {
        fe_sq(&acc, &acc);
        if ((e[i >> 3] >> (i & 7)) & 1) fe_mul(&acc, &acc, a);
    }
    *r = acc;
}
static void fe_invert(fe *r, const fe *a) {
    uint8_t e[32]; memset(e, 0xff, 32); e[0] = 0xeb; e[31] = 0x7f; /* p-2 */
    fe_pow(r, a, e);
}

static const uint8_t D_BYTES[32] = {0xa3, 0x78, 0x59, 0x13, 0xca, 0x4d, 0xeb, 0x75, 0xab, 0xd8, 0x41, 0x41, 0x4d, 0x0a, 0x70, 0x00,
                                    0x98, 0xe8, 0x79, 0x77, 0x79, 0x40, 0xc7, 0x8c, 0x73, 0xfe, 0x6f, 0x2b, 0xee, 0x6c, 0x03, 0x52};
static const uint8_t SQRTM1_BYTES[32] = {0xb0, 0xa0, 0x0e, 0x4a, 0x27, 0x1b, 0xee, 0xc4, 0x78, 0xe4, 0x2f, 0xad, 0x06, 0x18, 0x43, 0x2f,
                                         0xa7, 0xd7, 0xfb, 0x3d, 0x99, 0x00, 0x4d, 0x2b, 0x0b, 0xdf, 0xc1, 0x4f, 0x80, 0x24, 0x83, 0x2b};
static const uint8_t GX_BYTES[32] = {0x1a, 0xd5, 0x25, 0x8f, 0x60, 0x2d, 0x56, 0xc9, 0xb2, 0xa7, 0x25, 0x95, 0x60, 0xc7, 0x2c, 0x69,
                                     0x5c, 0xdc, 0xd6, 0xfd, 0x31, 0xe2, 0xa4, 0xc0, 0xfe, 0x53, 0x6e, 0xcd, 0xd3, 0x36, 0x69, 0x21};
static const uint8_t GY_BYTES[32] = {0x58, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66,
                                     0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66};
/* l = 2^252 + 27742317777372353535851937790883648493, little-endian */
static const uint8_t L_BYTES[32] = {0xed, 0xd3, 0xf5, 0x5c, 0x1a, 0x63, 0x12, 0x58, 0xd6, 0x9c, 0xf7, 0xa2, 0xde, 0xf9, 0xde, 0x14,
                                    0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x10};

/* starkyx `decompress`: (AffinePoint(x,y), root).  Returns 0 when (y^2-1)/(d y^2+1) is not a square
 * (the reference panics there). */
int orc_ed25519_decompress(const uint8_t in[32], uint8_t xy[64], uint8_t root_out[32]) {
    int sign = in[31] >> 7;
    fe y, yy, u, v, d, one, vinv, x2, beta, chk, negx2, sm1;
    fe_frombytes(&y, in); /* masks bit 255 */
    fe_frombytes(&d, D_BYTES);
    fe_frombytes(&sm1, SQRTM1_BYTES);
    fe_1(&one);
    fe_sq(&yy, &y);
    fe_sub(&u, &yy, &one);
    fe_mul(&v, &yy, &d);
    fe_add(&v, &v, &one); fe_carry(&v);
    fe_invert(&vinv, &v);
    fe_mul(&x2, &u, &vinv);
    /* sqrt: beta = x2^((p+3)/8); fix up with sqrt(-1); choose the even root */
    uint8_t e[32]; memset(e, 0xff, 32); e[0] = 0xfe; e[31] = 0x0f; /* (p+3)/8 = 2^252 - 2 */
    fe_pow(&beta, &x2, e);
    fe_sq(&chk, &beta);
    fe_neg(&negx2, &x2);
    if (fe_eq(&chk, &negx2) && !fe_iszero(&x2)) { fe_mul(&beta, &beta, &sm1); fe_sq(&chk, &beta); }
    int ok = fe_eq(&chk, &x2);
    if (!ok) { /* the reference panics here; convention shared with the CUDA kernel: identity, root 0 */
        memset(xy, 0, 64); xy[32] = 1; memset(root_out, 0, 32);
        return 0;
    }
    if (fe_isodd(&beta)) fe_neg(&beta, &beta);
    fe x = beta;
    if (sign) fe_neg(&x, &beta);
    fe_tobytes(root_out, &beta);
    fe_tobytes(xy, &x);
    fe_tobytes(xy + 32, &y);
    return ok;
}

/* extended twisted Edwards coordinates, a = -1 (complete addition law) */
typedef struct { fe X, Y, Z, T; } ge;

static void ge_from_affine(ge *p, const uint8_t xy[64]) {
    fe_frombytes(&p->X, xy); fe_frombytes(&p->Y, xy + 32); fe_1(&p->Z); fe_mul(&p->T, &p->X, &p->Y);
}
static void ge_identity(ge *p) { fe_0(&p->X); fe_1(&p->Y); fe_1(&p->Z); fe_0(&p->T); }

static void ge_add(ge *r, const ge *p, const ge *q) {
    fe a, b, c, d, e, f, g, h, t0, t1, d2;
    fe_frombytes(&d2, D_BYTES); fe_add(&d2, &d2, &d2); fe_carry(&d2);
    fe_sub(&t0, &p->Y, &p->X); fe_sub(&t1, &q->Y, &q->X); fe_mul(&a, &t0, &t1);
    fe_add(&t0, &p->Y, &p->X); fe_carry(&t0); fe_add(&t1, &q->Y, &q->X); fe_carry(&t1); fe_mul(&b, &t0, &t1);
    fe_mul(&c, &p->T, &q->T); fe_mul(&c, &c, &d2);
    fe_mul(&d, &p->Z, &q->Z); fe_add(&d, &d, &d); fe_carry(&d);
    fe_sub(&e, &b, &a); fe_sub(&f, &d, &c); fe_add(&g, &d, &c); fe_carry(&g); fe_add(&h, &b, &a); fe_carry(&h);
    fe_mul(&r->X, &e, &f); fe_mul(&r->Y, &g, &h); fe_mul(&r->T, &e, &h); fe_mul(&r->Z, &f, &g);
}
static void ge_dbl(ge *r, const ge *p) { ge_add(r, p, p); }

static void ge_to_affine(uint8_t xy[64], const ge *p) {
    fe zi, x, y;
    fe_invert(&zi, &p->Z); fe_mul(&x, &p->X, &zi); fe_mul(&y, &p->Y, &zi);
    fe_tobytes(xy, &x); fe_tobytes(xy + 32, &y);
}

/* EcOpResultHint::ScalarMul: point * scalar (scalar = 8 LE u32 limbs = 32 LE bytes, NOT reduced) */
void orc_ed25519_scalar_mul(const uint8_t scalar[32], const uint8_t xy[64], uint8_t out[64]) {
    ge acc, p;
    ge_identity(&acc);
    ge_from_affine(&p, xy);
    for (int i = 255; i >= 0; i--) {
        ge_dbl(&acc, &acc);
        if ((scalar[i >> 3] >> (i & 7)) & 1) ge_add(&acc, &acc, &p);
    }
    ge_to_affine(out, &acc);
}

/* EcOpResultHint::Add */
void orc_ed25519_add(const uint8_t a[64], const uint8_t b[64], uint8_t out[64]) {
    ge p, q, r;
    ge_from_affine(&p, a); ge_from_affine(&q, b); ge_add(&r, &p, &q); ge_to_affine(out, &r);
}

/* the literal affine law x3=(x1y2+x2y1)/(1+d x1x2y1y2), y3=(y1y2+x1x2)/(1-d x1x2y1y2), applied in an
 * LSB-first double-and-add, i.e. what a BigUint `AffinePoint * scalar` does.  Slow; used by the tests
 * to show the extended-coordinate result equals the affine one. */
static void affine_add(fe *x3, fe *y3, const fe *x1, const fe *y1, const fe *x2, const fe *y2) {
    fe d, x1y2, x2y1, y1y2, x1x2, t, one, den, inv, num;
    fe_frombytes(&d, D_BYTES); fe_1(&one);
    fe_mul(&x1y2, x1, y2); fe_mul(&x2y1, x2, y1); fe_mul(&y1y2, y1, y2); fe_mul(&x1x2, x1, x2);
    fe_mul(&t, &x1x2, &y1y2); fe_mul(&t, &t, &d);
    fe_add(&num, &x1y2, &x2y1); fe_carry(&num); fe_add(&den, &one, &t); fe_carry(&den); fe_invert(&inv, &den);
    fe rx; fe_mul(&rx, &num, &inv);
    fe_add(&num, &y1y2, &x1x2); fe_carry(&num); fe_sub(&den, &one, &t); fe_invert(&inv, &den);
    fe ry; fe_mul(&ry, &num, &inv);
    *x3 = rx; *y3 = ry;
}
void orc_ed25519_affine_double_and_add(const uint8_t scalar[32], const uint8_t xy[64], uint8_t out[64]) {
    fe rx, ry, tx, ty;
    fe_0(&rx); fe_1(&ry);
    fe_frombytes(&tx, xy); fe_frombytes(&ty, xy + 32);
    for (int i = 0; i < 256; i++) {
        if ((scalar[i >> 3] >> (i & 7)) & 1) affine_add(&rx, &ry, &rx, &ry, &tx, &ty);
        affine_add(&tx, &ty, &tx, &ty, &tx, &ty);
    }
    fe_tobytes(out, &rx); fe_tobytes(out + 32, &ry);
}

/* (div, rem) = LE512(digest).div_rem(l): bitwise shift-subtract over 64-bit limbs.
 * PX/frontend/uint/num/biguint/mod.rs:481-488 */
static void divrem_l(const uint8_t digest[64], uint8_t div[40], uint8_t rem[32]) {
    uint64_t l[5] = {0, 0, 0, 0, 0}, r[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) for (int j = 7; j >= 0; j--) l[i] = (l[i] << 8) | L_BYTES[8 * i + j];
    uint8_t q[64];
    memset(q, 0, sizeof q);
    for (int bit = 511; bit >= 0; bit--) {
        /* r = (r << 1) | bit */
        for (int i = 4; i > 0; i--) r[i] = (r[i] << 1) | (r[i - 1] >> 63);
        r[0] = (r[0] << 1) | ((digest[bit >> 3] >> (bit & 7)) & 1);
        /* if r >= l: r -= l */
        int ge_ = 1;
        for (int i = 4; i >= 0; i--) { if (r[i] != l[i]) { ge_ = r[i] > l[i]; break; } }
        if (ge_) {
            uint64_t borrow = 0;
            for (int i = 0; i < 5; i++) {
                uint64_t a = r[i], b = l[i];
                uint64_t dd = a - b - borrow;
                borrow = (a < b) || (a == b && borrow);
                r[i] = dd;
            }
            q[bit >> 3] |= (uint8_t)(1u << (bit & 7));
        }
    }
    memcpy(div, q, 40);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) rem[8 * i + j] = (uint8_t)(r[i] >> (8 * j));
}

static int lt_l(const uint8_t s[32]) {
    for (int i = 31; i >= 0; i--) { if (s[i] != L_BYTES[i]) return s[i] < L_BYTES[i]; }
    return 0;
}

/* eddsa.rs:161-203, one signature.  EC request order: ScalarMul(s,G), Decompress(pk), IsValid,
 * ScalarMul(h,A), Decompress(R), IsValid, Add(R, hA). */
void orc_ed25519_witness(const uint8_t pk[32], const uint8_t sig[64], const uint8_t *msg, uint32_t msg_len,
                         uint8_t *out) {
    uint8_t buf[64 + 1024];
    memset(out, 0, ORC_SIG_OUT_BYTES);
    if (msg_len > 1024) msg_len = 1024;
    memcpy(buf, sig, 32);
    memcpy(buf + 32, pk, 32);
    if (msg_len) memcpy(buf + 64, msg, msg_len);
    orc_sha512(buf, 64 + msg_len, out);           /* curta_sha512_variable(R‖A‖M, 64+len) */
    divrem_l(out, out + 96, out + 64);            /* h_scalar = rem (first 8 limbs), div */
    uint32_t flags = 0;
    if (lt_l(sig + 32)) flags |= 1;               /* s < l */
    uint8_t g[64];
    memcpy(g, GX_BYTES, 32); memcpy(g + 32, GY_BYTES, 32);
    orc_ed25519_scalar_mul(sig + 32, g, out + 136);                 /* p1 = s*G */
    if (orc_ed25519_decompress(pk, out + 200, out + 264)) flags |= 2; /* A, A_root */
    orc_ed25519_scalar_mul(out + 64, out + 200, out + 296);         /* h*A */
    if (orc_ed25519_decompress(sig, out + 360, out + 424)) flags |= 4; /* Rp, R_root */
    orc_ed25519_add(out + 360, out + 296, out + 456);               /* Rp + hA */
    if (memcmp(out + 136, out + 456, 64) == 0) flags |= 8;
    out[520] = (uint8_t)flags;
}

/* DUMMY constants, eddsa.rs:27-42 */
static const uint8_t DUMMY_PK[32] = {138, 136, 227, 221, 116, 9, 241, 149, 253, 82, 219, 45, 60, 186, 93, 114,
                                     202, 103, 9, 191, 29, 148, 18, 27, 243, 116, 136, 1, 180, 15, 111, 92};
static const uint8_t DUMMY_SIG[64] = {55, 20, 104, 158, 84, 120, 194, 17, 6, 237, 157, 164, 85, 88, 158, 137,
                                      187, 119, 187, 240, 159, 73, 80, 63, 133, 162, 74, 91, 48, 53, 6, 138,
                                      1, 41, 22, 121, 249, 46, 198, 145, 155, 102, 3, 210, 168, 135, 173, 55,
                                      252, 72, 45, 126, 169, 178, 191, 7, 153, 67, 112, 90, 150, 33, 140, 7};

/* curta_eddsa_verify_sigs_conditional, eddsa.rs:72-127: inactive lanes run on the dummy triple */
void orc_ed25519_batch(uint32_t n, const uint8_t *pks, const uint8_t *sigs, const uint8_t *msgs,
                       const uint32_t *msg_lens, const uint8_t *active, uint8_t *out, int threads) {
    static const uint8_t zero_msg[124] = {0};
    (void)threads;
#pragma omp parallel for schedule(dynamic) num_threads(threads > 0 ? threads : 1)
    for (uint32_t i = 0; i < n; i++) {
        if (!active || active[i])
            orc_ed25519_witness(pks + 32 * (size_t)i, sigs + 64 * (size_t)i, msgs + 124 * (size_t)i, msg_lens[i],
                                out + ORC_SIG_OUT_BYTES * (size_t)i);
        else
            orc_ed25519_witness(DUMMY_PK, DUMMY_SIG, zero_msg, 32, out + ORC_SIG_OUT_BYTES * (size_t)i);
    }
}

/* =====================================================================================================================
 * Ed25519 scalar-multiplication execution trace (EdDSA accelerator) -- the C restatement beside the
 * Python-integer one (oracle/ed_trace.py, which documents the construction and the layout; columns: include/bsx.h
 * BSX_ED25519_TRACE_COLS).  Reference: Ed25519Stark::prove, PX/frontend/ecc/curve25519/curta/stark.rs:182-219; the AIR
 * is starkyx's (un-vendored): layout our own, PARITY UNPINNED.  Deliberately NOT the kernels' way where there is a
 * choice: affine additions with a field inversion each (the kernels: extended coordinates + batched inversion), results
 * from the 51-bit-limb field arithmetic above (the kernels: from the integer division), quotients by an exact division
 * from the LOW end with p^-1 mod 2^64 (the kernels: an iteration from the high end).
 * ===================================================================================================================== */
#define EDT_COLS 1540
#define EDT_OFFSET (1 << 22)
static const uint64_t P_LIMBS64[4] = {0xffffffffffffffedULL, 0xffffffffffffffffULL, 0xffffffffffffffffULL, 0x7fffffffffffffffULL};

static void fe_limbs16(uint32_t l[16], const fe *a) {
    uint8_t s[32];
    fe_tobytes(s, a);
    for (int i = 0; i < 16; i++) l[i] = (uint32_t)s[2 * i] | ((uint32_t)s[2 * i + 1] << 8);
}
static void poly_mac(int64_t V[31], const uint32_t a[16], const uint32_t b[16]) {
    for (int i = 0; i < 16; i++)
        for (int j = 0; j < 16; j++) V[i + j] += (int64_t)((uint64_t)a[i] * b[j]);
}
/* M = V(2^16) - r >= 0, a multiple of p -> quotient limbs (16 x 16 bits); returns 0 if M is not an exact multiple */
static int exact_quotient(const int64_t V[31], const uint32_t r[16], uint32_t q16[16]) {
    /* V(2^16) - r as 9 x 64-bit limbs (two's complement while accumulating) */
    uint64_t M[9] = {0};
    __int128 c = 0;
    int64_t limbs[36] = {0};
    for (int k = 0; k < 31; k++) limbs[k] = V[k] - (r && k < 16 ? (int64_t)r[k] : 0);
    for (int k = 0; k < 36; k++) {           /* to 16-bit digits */
        c += limbs[k];
        limbs[k] = (int64_t)(c & 0xffff);
        c >>= 16;
    }
    if (c != 0) return 0;
    for (int k = 0; k < 36; k++) M[k / 4] |= (uint64_t)limbs[k] << (16 * (k % 4));
    /* p^-1 mod 2^64 by Newton's iteration (p odd) */
    uint64_t inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - P_LIMBS64[0] * inv;
    uint64_t q[5];
    for (int i = 0; i < 5; i++) {
        q[i] = M[i] * inv;
        /* M -= q[i] * p << (64 i) */
        unsigned __int128 borrow = 0, carry = 0;
        for (int j = 0; j < 4 && i + j < 9; j++) {
            const unsigned __int128 prod = (unsigned __int128)q[i] * P_LIMBS64[j] + carry;
            carry = prod >> 64;
            const unsigned __int128 sub = (unsigned __int128)(uint64_t)prod + borrow;
            const uint64_t before = M[i + j];
            M[i + j] = before - (uint64_t)sub;
            borrow = (sub >> 64) + (before < (uint64_t)sub ? 1 : 0);
        }
        for (int j = i + 4; j < 9; j++) {
            const unsigned __int128 sub = carry + borrow;
            carry = 0;
            const uint64_t before = M[j];
            M[j] = before - (uint64_t)sub;
            borrow = (sub >> 64) + (before < (uint64_t)sub ? 1 : 0);
            if (!borrow) break;
        }
    }
    for (int i = 0; i < 9; i++)
        if (M[i]) return 0;
    if (q[4] != 0) return 0;                 /* quotient below 2^256 */
    for (int i = 0; i < 16; i++) q16[i] = (uint32_t)(q[i / 4] >> (16 * (i % 4))) & 0xffff;
    return 1;
}
/* one operation's 92 columns at col (stride = rows per column).  V: left-hand-side polynomial; res: the result limbs
 * (subtracted from V unless the operation is a division, whose left-hand side already contains it) */
static int emit_op(int64_t V[31], const uint32_t res[16], int den, uint64_t *col, size_t stride) {
    uint32_t q[16];
    if (!exact_quotient(V, den ? NULL : res, q)) return 0;
    static const uint32_t P16[16] = {0xffed, 0xffff, 0xffff, 0xffff, 0xffff, 0xffff, 0xffff, 0xffff,
                                     0xffff, 0xffff, 0xffff, 0xffff, 0xffff, 0xffff, 0xffff, 0x7fff};
    int64_t van[31];
    for (int k = 0; k < 31; k++) van[k] = V[k] - (!den && k < 16 ? (int64_t)res[k] : 0);
    for (int i = 0; i < 16; i++)
        for (int j = 0; j < 16; j++) van[i + j] -= (int64_t)((uint64_t)q[i] * P16[j]);
    for (int k = 0; k < 16; k++) { col[(size_t)k * stride] = res[k]; col[(size_t)(16 + k) * stride] = q[k]; }
    int64_t prev = 0;
    for (int k = 0; k < 30; k++) {
        const int64_t num = prev - van[k];
        if (num & 0xffff) return 0;
        prev = num >> 16;                    /* arithmetic shift of an exact multiple */
        const int64_t sh = prev + EDT_OFFSET;
        if (sh < 0 || sh >= ((int64_t)1 << 32)) return 0;
        col[(size_t)(32 + k) * stride] = (uint64_t)sh & 0xffff;
        col[(size_t)(62 + k) * stride] = (uint64_t)sh >> 16;
    }
    return van[30] == prev;
}
/* the eight witnessed operations of (x1, y1) + (x2, y2); the sum leaves in (x3, y3) */
static int add_rows(const fe *x1, const fe *y1, const fe *x2, const fe *y2, fe *x3, fe *y3, uint64_t *col, size_t stride) {
    fe d, one, xn, yn, m1, m2, f, df, t, u, den, inv;
    uint32_t lx1[16], ly1[16], lx2[16], ly2[16], lr[16], la[16], lb[16];
    int64_t V[31];
    int ok = 1;
    fe_frombytes(&d, D_BYTES); fe_1(&one);
    fe_limbs16(lx1, x1); fe_limbs16(ly1, y1); fe_limbs16(lx2, x2); fe_limbs16(ly2, y2);
    const size_t op = 92 * stride;
    /* 0: xn = x1 y2 + x2 y1 */
    fe_mul(&t, x1, y2); fe_mul(&u, x2, y1); fe_add(&xn, &t, &u); fe_carry(&xn);
    memset(V, 0, sizeof V); poly_mac(V, lx1, ly2); poly_mac(V, lx2, ly1); fe_limbs16(lr, &xn); ok &= emit_op(V, lr, 0, col, stride);
    /* 1: yn = y1 y2 + x1 x2 */
    fe_mul(&t, y1, y2); fe_mul(&u, x1, x2); fe_add(&yn, &t, &u); fe_carry(&yn);
    memset(V, 0, sizeof V); poly_mac(V, ly1, ly2); poly_mac(V, lx1, lx2); fe_limbs16(lr, &yn); ok &= emit_op(V, lr, 0, col + op, stride);
    /* 2, 3: m1 = x1 y1, m2 = x2 y2 */
    fe_mul(&m1, x1, y1);
    memset(V, 0, sizeof V); poly_mac(V, lx1, ly1); fe_limbs16(lr, &m1); ok &= emit_op(V, lr, 0, col + 2 * op, stride);
    fe_mul(&m2, x2, y2);
    memset(V, 0, sizeof V); poly_mac(V, lx2, ly2); fe_limbs16(lr, &m2); ok &= emit_op(V, lr, 0, col + 3 * op, stride);
    /* 4: f = m1 m2 */
    fe_mul(&f, &m1, &m2);
    fe_limbs16(la, &m1); fe_limbs16(lb, &m2);
    memset(V, 0, sizeof V); poly_mac(V, la, lb); fe_limbs16(lr, &f); ok &= emit_op(V, lr, 0, col + 4 * op, stride);
    /* 5: df = d f */
    fe_mul(&df, &d, &f);
    fe_limbs16(la, &d); fe_limbs16(lb, &f);
    memset(V, 0, sizeof V); poly_mac(V, la, lb); fe_limbs16(lr, &df); ok &= emit_op(V, lr, 0, col + 5 * op, stride);
    /* 6: x3 = xn / (1 + df):  df x3 + x3 - xn = carry p */
    fe_add(&den, &one, &df); fe_carry(&den); fe_invert(&inv, &den); fe_mul(x3, &xn, &inv);
    fe_limbs16(la, &df); fe_limbs16(lr, x3); fe_limbs16(lb, &xn);
    memset(V, 0, sizeof V); poly_mac(V, la, lr);
    for (int k = 0; k < 16; k++) V[k] += (int64_t)lr[k] - (int64_t)lb[k];
    ok &= emit_op(V, lr, 1, col + 6 * op, stride);
    /* 7: y3 = yn / (1 - df):  df y3 + yn - y3 = carry p */
    fe_sub(&den, &one, &df); fe_invert(&inv, &den); fe_mul(y3, &yn, &inv);
    fe_limbs16(lr, y3); fe_limbs16(lb, &yn);
    memset(V, 0, sizeof V); poly_mac(V, la, lr);
    for (int k = 0; k < 16; k++) V[k] += (int64_t)lb[k] - (int64_t)lr[k];
    ok &= emit_op(V, lr, 1, col + 7 * op, stride);
    return ok;
}
static int mul_rows(const uint8_t *scalar, const fe *px, const fe *py, int real, uint64_t *col0, size_t stride, uint8_t *result) {
    fe tx = *px, ty = *py, ax, ay, sx, sy, dx, dy;
    int ok = 1;
    fe_0(&ax); fe_1(&ay);
    for (int j = 0; j < 256; j++) {
        uint64_t *col = col0 + j;
        const int bit = scalar ? (scalar[j >> 3] >> (j & 7)) & 1 : 0;
        uint32_t l[16];
        col[0] = (uint64_t)bit; col[stride] = (uint64_t)real; col[2 * stride] = j == 0; col[3 * stride] = j == 255;
        fe_limbs16(l, &tx); for (int i = 0; i < 16; i++) col[(size_t)(4 + i) * stride] = l[i];
        fe_limbs16(l, &ty); for (int i = 0; i < 16; i++) col[(size_t)(20 + i) * stride] = l[i];
        fe_limbs16(l, &ax); for (int i = 0; i < 16; i++) col[(size_t)(36 + i) * stride] = l[i];
        fe_limbs16(l, &ay); for (int i = 0; i < 16; i++) col[(size_t)(52 + i) * stride] = l[i];
        ok &= add_rows(&ax, &ay, &tx, &ty, &sx, &sy, col + 68 * stride, stride);
        ok &= add_rows(&tx, &ty, &tx, &ty, &dx, &dy, col + (68 + 8 * 92) * stride, stride);
        if (bit) { ax = sx; ay = sy; }
        tx = dx; ty = dy;
    }
    if (result) { fe_tobytes(result, &ax); fe_tobytes(result + 32, &ay); }
    return ok;
}
/* trace[EDT_COLS][2^log_rows]; results: n_muls x 64 bytes or NULL; returns 1, or 0 if an identity failed to hold (it cannot) */
int orc_ed25519_trace(const uint8_t *scalars, const uint8_t *points, uint32_t n_muls, uint32_t log_rows, uint8_t *results,
                      uint64_t *trace, int threads) {
    const size_t n_rows = (size_t)1 << log_rows;
    if ((size_t)n_muls * 256 > n_rows || log_rows < 8) return 0;
    const uint32_t slots = (uint32_t)(n_rows / 256);
    int ok = 1;
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic) num_threads(threads) reduction(& : ok)
    for (uint32_t m = 0; m < slots; m++) {
        fe px, py;
        if (m < n_muls) {
            fe_frombytes(&px, points + (size_t)m * 64); fe_frombytes(&py, points + (size_t)m * 64 + 32);
            ok &= mul_rows(scalars + (size_t)m * 32, &px, &py, 1, trace + (size_t)m * 256, n_rows, results ? results + (size_t)m * 64 : NULL);
        } else {
            fe_0(&px); fe_1(&py);
            ok &= mul_rows(NULL, &px, &py, 0, trace + (size_t)m * 256, n_rows, NULL);
        }
    }
    return ok;
}

"""Ed25519 scalar-multiplication execution trace, restated with Python integers.

TEST INFRASTRUCTURE ONLY -- never imported by the product package; small cases only (pure-Python loops).

What it restates: the trace half of `Ed25519Stark::prove` (PX/frontend/ecc/curve25519/curta/stark.rs:182-219: 256 rows per
scalar multiplication, `write_trace_instructions` per row) for the `ScalarMul` operations of the EdDSA schedule
(PX/frontend/ecc/curve25519/curta/stark.rs:93-124 collects them; two per signature, s*G and h*A).  The AIR itself is
starkyx's `scalar_mul_batch` (UN-VENDORED, starkyx@ad8eb4ba): nothing under /root/reference holds its column assignment,
so the LAYOUT IS OURS and PARITY IS UNPINNED.  What is restated is the published construction: an affine
double-and-add, one bit per row, every field operation of the two Edwards additions of a row witnessed the way starkyx's
`FpMulInstruction` / `FpInnerProductInstruction` / `FpMulConstInstruction` / `FpDenInstruction` do it -- operands and
result as 16 limbs of 16 bits, the quotient `carry` with  lhs - result = carry * p,  and the witness polynomial
w(x) = (lhs(x) - result(x) - carry(x) p(x)) / (x - 2^16)  stored as (w_k + OFFSET) split into a low and a high 16-bit half.

Layout (include/bsx.h BSX_ED25519_TRACE_COLS; column c, row r at trace[c, r]); row 256 m + j is step j of multiplication m:
  0 bit j of the scalar | 1 real row (0 on padding) | 2 j == 0 | 3 j == 255
  4..19 temp.x  20..35 temp.y  (temp = 2^j P) | 36..51 acc.x  52..67 acc.y  (acc = (k mod 2^j) P)
  68 + 92 o, o = 0..15: field operation o = result[16] carry[16] witness_low[30] witness_high[30]
     o = 0..7 : sum = acc + temp      o = 8..15 : dbl = temp + temp, each Edwards addition (x1, y1) + (x2, y2) being
     0 xn = x1 y2 + x2 y1   1 yn = y1 y2 + x1 x2   2 m1 = x1 y1   3 m2 = x2 y2   4 f = m1 m2   5 df = d f
     6 x3 = xn / (1 + df)   7 y3 = yn / (1 - df)
  next row: temp' = dbl, acc' = bit ? sum : acc.   Padding rows (beyond 256 * n) are the rows of 0 * (0, 1) with column 1 = 0.
"""
from __future__ import annotations

import numpy as np

P = 2**255 - 19
D = (-121665 * pow(121666, P - 2, P)) % P
NL, NW, B = 16, 30, 1 << 16
OFFSET = 1 << 22            # |w_k| < 2^22 for every operation here (two 16-term products of 16-bit limbs per coefficient)
COLS = 68 + 16 * 92
P_LIMBS = [(P >> (16 * i)) & 0xFFFF for i in range(NL)]


def limbs(v: int):
    assert 0 <= v < 1 << 256
    return [(v >> (16 * i)) & 0xFFFF for i in range(NL)]


def _pmul(a, b):
    out = [0] * (2 * NL - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            out[i + j] += x * y
    return out


def _witness(vanishing):
    """v(x) = (x - 2^16) w(x): exact synthetic division from the low end; returns (low, high) halves of w_k + OFFSET."""
    assert len(vanishing) == 2 * NL - 1
    w, prev = [], 0
    for k in range(NW):
        num = prev - vanishing[k]
        assert num % B == 0
        prev = num // B
        w.append(prev)
    assert vanishing[2 * NL - 2] == w[-1]
    sh = [x + OFFSET for x in w]
    assert all(0 <= x < 1 << 32 for x in sh)
    return [x & 0xFFFF for x in sh], [x >> 16 for x in sh]


def _emit(lhs_poly, lhs_int, result):
    """lhs_int = carry * p + result (mul-type) with lhs_poly(2^16) = lhs_int; the vanishing polynomial subtracts result."""
    carry, rem = divmod(lhs_int - result, P)
    assert rem == 0 and 0 <= carry < 1 << 256
    rl, cl = limbs(result), limbs(carry)
    cp = _pmul(cl, P_LIMBS)
    van = [lhs_poly[k] - (rl[k] if k < NL else 0) - cp[k] for k in range(2 * NL - 1)]
    lo, hi = _witness(van)
    return rl + cl + lo + hi


def fp_mul(a: int, b: int):
    r = a * b % P
    return r, _emit(_pmul(limbs(a), limbs(b)), a * b, r)


def fp_inner(a1, b1, a2, b2):
    r = (a1 * b1 + a2 * b2) % P
    poly = [x + y for x, y in zip(_pmul(limbs(a1), limbs(b1)), _pmul(limbs(a2), limbs(b2)))]
    return r, _emit(poly, a1 * b1 + a2 * b2, r)


def fp_den(a: int, b: int, plus: bool):
    """result = a / (1 + b) (plus) or a / (1 - b): the equation is  b res + res - a = carry p  resp.  b res + a - res = carry p,
    so the `result` subtracted inside _emit is already part of the left-hand side (passed as 0 there)."""
    den = (1 + b) % P if plus else (1 - b) % P
    res = a * pow(den, P - 2, P) % P
    al, rl = limbs(a), limbs(res)
    poly = _pmul(limbs(b), rl)
    for k in range(NL):
        poly[k] += (rl[k] - al[k]) if plus else (al[k] - rl[k])
    n = b * res + (res - a if plus else a - res)
    carry, rem = divmod(n, P)
    assert rem == 0 and 0 <= carry < 1 << 256
    cl = limbs(carry)
    cp = _pmul(cl, P_LIMBS)
    lo, hi = _witness([poly[k] - cp[k] for k in range(2 * NL - 1)])
    return res, rl + cl + lo + hi


def ed_add(x1, y1, x2, y2):
    """the eight witnessed field operations of one affine Edwards addition (a = -1); returns ((x3, y3), 8 * 92 values)"""
    cols = []
    xn, c = fp_inner(x1, y2, x2, y1); cols += c
    yn, c = fp_inner(y1, y2, x1, x2); cols += c
    m1, c = fp_mul(x1, y1); cols += c
    m2, c = fp_mul(x2, y2); cols += c
    f, c = fp_mul(m1, m2); cols += c
    df, c = fp_mul(D, f); cols += c
    x3, c = fp_den(xn, df, True); cols += c
    y3, c = fp_den(yn, df, False); cols += c
    return (x3, y3), cols


def scalar_mul_rows(k: int, pt):
    """256 rows of k * pt; returns (rows as lists of COLS ints, result point)"""
    temp, acc, rows = pt, (0, 1), []
    for j in range(256):
        bit = (k >> j) & 1
        s, cs = ed_add(acc[0], acc[1], temp[0], temp[1])
        d, cd = ed_add(temp[0], temp[1], temp[0], temp[1])
        rows.append([bit, 1, int(j == 0), int(j == 255)] + limbs(temp[0]) + limbs(temp[1]) + limbs(acc[0]) + limbs(acc[1]) + cs + cd)
        acc = s if bit else acc
        temp = d
    return rows, acc


def ed25519_trace(scalars, points, log_rows: int):
    """scalars: ints < 2^256; points: affine (x, y) ints in [0, p) -> (uint64 [COLS, 2^log_rows], results [(x, y)])"""
    n_rows = 1 << log_rows
    assert 256 * len(scalars) <= n_rows
    tr = np.zeros((COLS, n_rows), np.uint64)
    results = []
    for m, (k, pt) in enumerate(zip(scalars, points)):
        rows, res = scalar_mul_rows(k, pt)
        tr[:, 256 * m:256 * (m + 1)] = np.array(rows, np.uint64).T
        results.append(res)
    if 256 * len(scalars) < n_rows:
        rows, _ = scalar_mul_rows(0, (0, 1))
        pad = np.array(rows, np.uint64).T
        pad[1, :] = 0
        for m in range(len(scalars), n_rows // 256):
            tr[:, 256 * m:256 * (m + 1)] = pad
    return tr, results

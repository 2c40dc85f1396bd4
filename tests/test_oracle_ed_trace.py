"""The Ed25519 scalar-multiplication trace (include/bsx.h BSX_ED25519_TRACE_COLS), on the CPU:

1. oracle/ed_trace.py (Python integers) is pinned by RE-CHECKING the trace it emits with code that shares nothing with it
   but the column table: every field operation's identity as an integer equation and as a polynomial identity in the
   16-bit limbs, the Edwards-addition relations between operations, the row-to-row chaining, and k * P against the
   pinned signature oracle -- incl. s * G and h * A of real mocha-4 commit signatures (tests/golden/mocha4.json).
   The reference's own column assignment is starkyx's (un-vendored): parity with it is UNPINNED and says so everywhere.
2. the kernel-side code (blobstreamx_b200/csrc/ed_trace.cuh) compiled for the host reproduces that trace bit for bit.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from blobstreamx_b200 import inputs as bx_inputs
from oracle import ed_trace as T
from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
P, B = 2**255 - 19, 1 << 16
L = 2**252 + 27742317777372353535851937790883648493


def _val(tr, col, row, n=16):
    return sum(int(tr[col + i, row]) << (16 * i) for i in range(n))


def _poly(tr, col, row, n):
    return [int(tr[col + i, row]) for i in range(n)]


def _pmul(a, b):
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            out[i + j] += x * y
    return out


def _padd(*ps):
    out = [0] * max(len(p) for p in ps)
    for p in ps:
        for i, x in enumerate(p):
            out[i] += x
    return out


def _neg(p):
    return [-x for x in p]


P_POLY = [0xFFED] + [0xFFFF] * 14 + [0x7FFF]
D_POLY = [(T.D >> (16 * i)) & 0xFFFF for i in range(16)]


def check_op(tr, row, o, lhs_poly, lhs_int, den=False):
    """operation o of the row: lhs = carry * p + result as integers, and lhs(x) - result(x) - carry(x) p(x) = (x - 2^16) w(x)
    as polynomials, w = (low + 2^16 high) - OFFSET; for a division the result is already inside lhs."""
    c = 68 + 92 * o
    res, carry = _val(tr, c, row), _val(tr, c + 16, row)
    assert res < P
    assert lhs_int == carry * P + (0 if den else res)
    w = [int(tr[c + 32 + k, row]) + (int(tr[c + 62 + k, row]) << 16) - T.OFFSET for k in range(30)]
    van = _padd(lhs_poly, _neg(_pmul(_poly(tr, c + 16, row, 16), P_POLY)), [] if den else _neg(_poly(tr, c, row, 16)))
    van += [0] * (31 - len(van))
    assert van == _padd(_pmul(w, [-B, 1]))
    return res


def check_add(tr, row, o0, p1, p2):
    """the eight operations at o0.. of (x1, y1) + (x2, y2); returns the sum read off the trace"""
    (x1, y1), (x2, y2) = p1, p2
    l = lambda v: [(v >> (16 * i)) & 0xFFFF for i in range(16)]
    xn = check_op(tr, row, o0, _padd(_pmul(l(x1), l(y2)), _pmul(l(x2), l(y1))), x1 * y2 + x2 * y1)
    yn = check_op(tr, row, o0 + 1, _padd(_pmul(l(y1), l(y2)), _pmul(l(x1), l(x2))), y1 * y2 + x1 * x2)
    m1 = check_op(tr, row, o0 + 2, _pmul(l(x1), l(y1)), x1 * y1)
    m2 = check_op(tr, row, o0 + 3, _pmul(l(x2), l(y2)), x2 * y2)
    f = check_op(tr, row, o0 + 4, _pmul(l(m1), l(m2)), m1 * m2)
    df = check_op(tr, row, o0 + 5, _pmul(D_POLY, l(f)), T.D * f)
    x3 = _val(tr, 68 + 92 * (o0 + 6), row)
    y3 = _val(tr, 68 + 92 * (o0 + 7), row)
    check_op(tr, row, o0 + 6, _padd(_pmul(l(df), l(x3)), l(x3), _neg(l(xn))), df * x3 + x3 - xn, den=True)
    check_op(tr, row, o0 + 7, _padd(_pmul(l(df), l(y3)), l(yn), _neg(l(y3))), df * y3 + yn - y3, den=True)
    assert (x3 * (1 + df) - xn) % P == 0 and (y3 * (1 - df) - yn) % P == 0
    return x3, y3


def check_trace(tr, scalars, points, rows=None):
    """every row (or the given ones) re-derived from the columns alone + the chaining between consecutive rows"""
    n_rows = tr.shape[1]
    assert tr.shape[0] == T.COLS and int(tr.max()) < B
    for r in (range(n_rows) if rows is None else rows):
        m, j = divmod(r, 256)
        real = m < len(scalars)
        k, pt = (scalars[m], points[m]) if real else (0, (0, 1))
        assert [int(tr[c, r]) for c in range(4)] == [(k >> j) & 1, int(real), int(j == 0), int(j == 255)]
        temp = (_val(tr, 4, r), _val(tr, 20, r))
        acc = (_val(tr, 36, r), _val(tr, 52, r))
        if j == 0:
            assert temp == pt and acc == (0, 1)
        s = check_add(tr, r, 0, acc, temp)
        d = check_add(tr, r, 8, temp, temp)
        if j < 255:
            assert (_val(tr, 4, r + 1), _val(tr, 20, r + 1)) == d
            assert (_val(tr, 36, r + 1), _val(tr, 52, r + 1)) == (s if (k >> j) & 1 else acc)


def _cases(rng, n_random):
    pts = [po.G, po.ed_mul(int.from_bytes(rng.bytes(32), "little"), po.G), (0, 1), (0, P - 1)]     # (0, -1): order 2
    ks = [0, 1, 2**256 - 1, L, L - 1, 2**255, 1 << 200, 0xAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA, 2**252]
    scalars, points = [], []
    for i, k in enumerate(ks):
        scalars.append(k); points.append(pts[i % len(pts)])
    for _ in range(n_random):
        scalars.append(int.from_bytes(rng.bytes(32), "little"))
        points.append(po.ed_mul(int.from_bytes(rng.bytes(32), "little") % L, po.G))
    return scalars, points


def test_oracle_trace_is_a_valid_double_and_add():
    rng = np.random.default_rng(7)
    scalars, points = _cases(rng, 1)
    tr, results = T.ed25519_trace(scalars, points, 12)                  # 10 multiplications in 16 slots: 6 padding ones
    for k, pt, res in zip(scalars, points, results):
        assert res == po.ed_mul(k, pt)
    # all rows of three multiplications and of one padding multiplication, a stride over the rest
    rows = list(range(0, 256)) + list(range(256 * 2, 256 * 3)) + list(range(256 * 9, 256 * 11)) + list(range(0, 4096, 37))
    check_trace(tr, scalars, points, rows)
    last = 256 * 2 + 255                                                # k = 2^256 - 1: the result is the sum of the last row
    assert (_val(tr, 68 + 92 * 6, last), _val(tr, 68 + 92 * 7, last)) == results[2]


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "mocha4.json")) as f:
        return json.load(f)


def _fixture_muls(golden, height, n_sigs):
    """(s, G) and (h, A) of the first signed lanes of a fixture commit + the oracle's s*G / h*A"""
    hdr, commit, vals = golden["headers"][height], golden["commits"][height], golden["validators"][height]
    recs = bx_inputs.get_validator_data_from_block(vals, hdr, commit, 100)
    scalars, points, want = [], [], []
    for i, cs in enumerate(commit["signatures"]):
        if int(cs["block_id_flag"]) != 2 or len(scalars) >= 2 * n_sigs:
            continue
        r = recs[i]
        pk, sig = r[0:32].tobytes(), r[32:96].tobytes()
        msg = r[96:96 + int.from_bytes(r[220:224].tobytes(), "little")].tobytes()
        w = po.ed_witness(pk, sig, msg)
        assert w["verified"]
        scalars += [int.from_bytes(sig[32:], "little"), w["h"]]
        points += [po.G, w["A"]]
        want += [w["sG"], w["hA"]]
    return scalars, points, want


def test_fixture_signatures_through_the_trace(golden):
    scalars, points, want = _fixture_muls(golden, "157001", 2)
    tr, results = T.ed25519_trace(scalars, points, 10)
    assert results == want
    check_trace(tr, scalars, points, list(range(0, 1024, 5)) + [255, 511, 767, 1023])


@pytest.fixture(scope="module")
def hc():
    d = os.path.join(HERE, "host_check")
    subprocess.check_call(["make", "-C", d, "-s"], stderr=subprocess.DEVNULL)
    return C.CDLL(os.path.join(d, "libed_host_check.so"))


def pack(scalars, points):
    sc = np.frombuffer(b"".join(k.to_bytes(32, "little") for k in scalars), np.uint8).copy()
    pt = np.frombuffer(b"".join(x.to_bytes(32, "little") + y.to_bytes(32, "little") for x, y in points), np.uint8).copy()
    return sc, pt


def test_kernel_source_on_the_host_reproduces_the_oracle_trace(hc, golden):
    rng = np.random.default_rng(8)
    scalars, points = _cases(rng, 2)
    fs, fp, _ = _fixture_muls(golden, "10000", 1)
    scalars, points = scalars + fs, points + fp
    n, log_rows = len(scalars), 12
    assert 256 * n < (1 << log_rows)
    want, results = T.ed25519_trace(scalars, points, log_rows)
    sc, pt = pack(scalars, points)
    tr = np.zeros((T.COLS, 1 << log_rows), np.uint64)
    res = np.zeros((n, 64), np.uint8)
    hc.hc_ed25519_trace(sc.ctypes.data_as(C.c_void_p), pt.ctypes.data_as(C.c_void_p), C.c_uint32(n), C.c_uint32(log_rows),
                        res.ctypes.data_as(C.c_void_p), tr.ctypes.data_as(C.c_void_p))
    assert np.array_equal(tr, want)
    assert res.tobytes() == b"".join(po.ed_point_bytes(p) for p in results)


def _sqrt(a):
    a %= P
    if a == 0:
        return 0
    if pow(a, (P - 1) // 2, P) != 1:
        return None
    r = pow(a, (P + 3) // 8, P)
    if r * r % P != a:
        r = r * pow(2, (P - 1) // 4, P) % P
    return r


def _points_with_small_xy(t_max):
    """curve points with x * y = t for small t: -x^2 + t^2 / x^2 = 1 + d t^2 is a quadratic in x^2"""
    out = []
    for t in range(1, t_max):
        c = (1 + T.D * t * t) % P
        disc = _sqrt(c * c + 4 * t * t)
        if disc is None:
            continue
        for sgn in (1, -1):
            x = _sqrt((-c + sgn * disc) * pow(2, P - 2, P))
            if x:
                y = t * pow(x, P - 2, P) % P
                assert (-x * x + y * y - 1 - T.D * x * x * y * y) % P == 0
                out.append((t, (x, y)))
                break
    return out


def test_quotient_corner_remainders_below_19(hc):
    """The kernel-side quotient by p = 2^255 - 19 iterates q <- (N + 19 q) >> 255 and lands one below the quotient when the
    remainder is smaller than 19 (q* - q) -- always for the exact divisions, never by chance for a product.  Curve points with
    x * y = t, t = 5, 8, 10, 16, 17 (and 25, 28, 32, 37 on the other side of 19), put such remainders into the products
    m1 = x1 y1 / m2 = x2 y2 of the first row: kernel source on the host == Python integers on every row, and the independent
    checker accepts the rows."""
    rng = np.random.default_rng(19)
    pts = _points_with_small_xy(40)
    assert [t for t, _ in pts if t < 19] == [5, 8, 10, 16, 17] and len(pts) >= 8
    points = [p for _, p in pts]
    scalars = [int.from_bytes(rng.bytes(32), "little") | 1 for _ in points]
    n, log_rows = len(scalars), 12
    want, results = T.ed25519_trace(scalars, points, log_rows)
    for m, (t, _) in enumerate(pts):
        assert _val(want, 68 + 92 * 10, 256 * m) == t and _val(want, 68 + 92 * 3, 256 * m) == t   # m1 of the doubling, m2 of the sum
        assert results[m] == po.ed_mul(scalars[m], points[m])
    sc, pt = pack(scalars, points)
    tr = np.zeros((T.COLS, 1 << log_rows), np.uint64)
    hc.hc_ed25519_trace(sc.ctypes.data_as(C.c_void_p), pt.ctypes.data_as(C.c_void_p), C.c_uint32(n), C.c_uint32(log_rows), None,
                        tr.ctypes.data_as(C.c_void_p))
    assert np.array_equal(tr, want)
    check_trace(tr, scalars, points, [256 * m for m in range(n)] + [256 * m + 1 for m in range(n)])


def test_c_restatement_equals_the_python_one(golden):
    """oracle/ed25519.c orc_ed25519_trace (affine additions with an inversion each, results from 51-bit-limb field arithmetic,
    quotients by exact division from the low end) == oracle/ed_trace.py (Python divmod) on the edge cases, the quotient-corner
    points and fixture signatures; then ALL 98 real signatures of the 157001 commit through the C one: s*G and h*A as the pinned
    signature oracle computes them, rows accepted by the independent checker."""
    from oracle import cbind as orc
    rng = np.random.default_rng(23)
    scalars, points = _cases(rng, 2)
    corner = [p for _, p in _points_with_small_xy(19)]
    scalars, points = scalars + [int.from_bytes(rng.bytes(32), "little") | 1 for _ in corner], points + corner
    fs, fp, _ = _fixture_muls(golden, "10000", 1)
    scalars, points = scalars + fs, points + fp
    want, results = T.ed25519_trace(scalars, points, 13)
    sc, pt = pack(scalars, points)
    got, res = orc.ed25519_trace(sc, pt, 13, threads=orc.max_threads())
    assert np.array_equal(got, want) and res.tobytes() == b"".join(po.ed_point_bytes(p) for p in results)
    scalars, points, want_pts = _fixture_muls(golden, "157001", 98)
    assert len(scalars) == 196
    sc, pt = pack(scalars, points)
    tr, res = orc.ed25519_trace(sc, pt, 16, threads=orc.max_threads())
    assert res.tobytes() == b"".join(po.ed_point_bytes(p) for p in want_pts)
    check_trace(tr, scalars, points, list(range(0, 196 * 256, 997)) + [196 * 256 - 1, 196 * 256, 65535])


def test_field_operation_at_operand_extremes(hc):
    """One witnessed operation (ed_trace.cuh edt_op_rt, host build) on operands at the ends of [0, p): the limb products, the
    quotient and the witness offset are at their largest for p - 1 everywhere (an inner product of four such operands), at
    their smallest for 0 / 1; divisions with the matching result.  Every output equals the Python restatement's, whose own
    assertions bound the witness (0 <= w + 2^22 < 2^32) and whose columns the checker above accepts."""
    rng = np.random.default_rng(41)
    ext = [0, 1, 2, 18, 19, P - 1, P - 2, P - 19, 2**255 - 20, 1 << 254, (1 << 255) - (1 << 240), int("55" * 32, 16) % P, int("aa" * 32, 16) % P]
    vals = ext + [int.from_bytes(rng.bytes(32), "little") % P for _ in range(12)]
    u32 = C.c_uint32 * 16

    def run(kind, a1, b1, a2, b2, res, alias=False):
        r = u32(*T.limbs(res))
        out = (C.c_uint64 * 92)()
        pa1, pb1 = u32(*T.limbs(a1)), u32(*T.limbs(b1))
        pa2, pb2 = (pa1, pb1) if alias else (u32(*T.limbs(a2)), u32(*T.limbs(b2)))     # alias: the doubling's x y + x y, computed once
        hc.hc_edt_op(kind, pa1, pb1, pa2, pb2, r, out)
        return list(out), sum(int(x) << (16 * i) for i, x in enumerate(r))

    n = 0
    for i, a in enumerate(vals):
        for b in (vals[(i * 5 + 1) % len(vals)], vals[(i * 3 + 2) % len(vals)], a, P - 1, 0):
            want_r, want_cols = T.fp_mul(a, b)
            cols, r = run(0, a, b, 0, 0, 0)
            assert cols == want_cols and r == want_r
            c, d = vals[(i * 7 + 3) % len(vals)], vals[(i + 4) % len(vals)]
            for (x1, y1, x2, y2) in ((a, b, c, d), (a, b, a, b), (P - 1, P - 1, P - 1, P - 1)):
                want_r, want_cols = T.fp_inner(x1, y1, x2, y2)
                cols, r = run(1, x1, y1, x2, y2, 0)
                assert cols == want_cols and r == want_r
                if (x1, y1) == (x2, y2):
                    cols, r = run(1, x1, y1, x2, y2, 0, alias=True)
                    assert cols == want_cols and r == want_r
            for plus in (True, False):
                den = (1 + b) % P if plus else (1 - b) % P
                if den == 0:
                    continue
                want_r, want_cols = T.fp_den(a, b, plus)
                cols, _ = run(2 if plus else 3, b, 0, a, 0, want_r)
                assert cols == want_cols
                n += 1
    assert n > 200

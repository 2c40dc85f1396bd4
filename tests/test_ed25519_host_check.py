"""The kernel-side Ed25519 arithmetic (blobstreamx_b200/csrc/ed25519.cuh) compiled for the HOST and
checked against the oracle.  This validates the code the GPU runs (same source, same limb schedule)
without a GPU; the GPU parity test (tests/test_gpu_ed25519.py) repeats it through the C ABI."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_check")
P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493


@pytest.fixture(scope="module")
def hc():
    subprocess.check_call(["make", "-C", HERE, "-s"], stderr=subprocess.DEVNULL)
    return C.CDLL(os.path.join(HERE, "libed_host_check.so"))


def _buf(b):
    return (C.c_uint8 * len(b)).from_buffer_copy(bytes(b))


def _out(n):
    return (C.c_uint8 * n)()


def _fe_cases(rng, n):
    special = [0, 1, 2, 19, P - 1, P - 2, P, P + 1, 2**255 - 1, 2**255 - 20, (1 << 254), (1 << 255) - (1 << 200),
               int("aa" * 32, 16) >> 1, int("55" * 32, 16), (1 << 26) - 1, ((1 << 255) - 1) ^ ((1 << 51) - 1)]
    return special + [int.from_bytes(rng.bytes(32), "little") >> 1 for _ in range(n)]


def test_field_ops(hc):
    rng = np.random.default_rng(5)
    xs = _fe_cases(rng, 200)
    ys = list(reversed(_fe_cases(rng, 200)))
    for a, b in zip(xs, ys):
        ab, bb = a.to_bytes(32, "little"), b.to_bytes(32, "little")
        o = _out(32)
        hc.hc_fe_mul(_buf(ab), _buf(bb), o)
        assert int.from_bytes(bytes(o), "little") == (a * b) % P
        hc.hc_fe_sq(_buf(ab), o, 0)
        assert int.from_bytes(bytes(o), "little") == (a * a) % P
        hc.hc_fe_sq(_buf(ab), o, 1)
        assert int.from_bytes(bytes(o), "little") == (2 * a * a) % P
        hc.hc_fe_addsub_mul(_buf(ab), _buf(bb), o)
        assert int.from_bytes(bytes(o), "little") == ((a + b) * (a - b)) % P
    for a in xs[:40]:
        if a % P == 0:
            continue
        o = _out(32)
        hc.hc_fe_invert(_buf(a.to_bytes(32, "little")), o)
        assert int.from_bytes(bytes(o), "little") == pow(a, P - 2, P)


OFF = [0, 26, 51, 77, 102, 128, 153, 179, 204, 230]


def _limb_value(v):
    return sum(int(x) << o for x, o in zip(v, OFF))


def test_field_ops_at_limb_bounds(hc):
    """fe_mul takes f up to 3 units and g up to 2 units (1 unit = a carried value: |even| <= 1.01*2^25,
    |odd| <= 1.01*2^24), fe_sq up to 2 units, fe_tighten brings 3 units back under 2 -- exercised at the extremes
    (all-max, all-min, alternating, random signs), where a 32-bit overflow in the 19x terms would show."""
    rng = np.random.default_rng(11)
    unit = np.array([int(1.01 * 2**25) if i % 2 == 0 else int(1.01 * 2**24) for i in range(10)], np.int64)
    pats = [np.ones(10, np.int64), -np.ones(10, np.int64), np.array([1, -1] * 5), np.array([-1, 1] * 5),
            np.array([1, 1, -1, -1, 1, 1, -1, -1, 1, 1]), np.array([-1, -1, -1, -1, -1, 1, 1, 1, 1, 1])]
    pats += [rng.choice([-1, 1], 10) for _ in range(60)]
    i32 = C.c_int32 * 10
    for k, pf in enumerate(pats):
        for pg in (pats[(k + 1) % len(pats)], pats[(k * 7 + 3) % len(pats)], pf):
            frac_f = 1.0 if k < 12 else rng.uniform(0.5, 1.0, 10)
            frac_g = 1.0 if k < 12 else rng.uniform(0.5, 1.0, 10)
            f = (pf * unit * 3 * frac_f).astype(np.int64)
            g = (pg * unit * 2 * frac_g).astype(np.int64)
            o = _out(32)
            hc.hc_fe_mul_limbs(i32(*f.tolist()), i32(*g.tolist()), o)
            assert int.from_bytes(bytes(o), "little") == (_limb_value(f) * _limb_value(g)) % P
            for tw in (0, 1):
                hc.hc_fe_sq_limbs(i32(*g.tolist()), o, tw)
                assert int.from_bytes(bytes(o), "little") == ((1 + tw) * _limb_value(g) ** 2) % P
            t = i32()
            hc.hc_fe_tighten_limbs(i32(*f.tolist()), t)
            t = np.array(list(t), np.int64)
            assert _limb_value(t) % P == _limb_value(f) % P
            assert all(-38 <= t[i] <= (2**26 + 19 if i % 2 == 0 else 2**25 + 1) for i in range(10))
            # a tightened value is a legal g operand
            hc.hc_fe_mul_limbs(i32(*f.tolist()), i32(*t.tolist()), o)
            assert int.from_bytes(bytes(o), "little") == (_limb_value(f) * _limb_value(t)) % P


def test_divrem_l(hc):
    rng = np.random.default_rng(6)
    cases = [0, 1, L - 1, L, L + 1, 2 * L - 1, 2 * L, 2**512 - 1, 2**511, L * L, L * L - 1, (2**260 - 1) * L, (2**260 - 1) * L + L - 1,
             2**252, 2**253 - 1, 2**504]
    cases += [int.from_bytes(rng.bytes(64), "little") for _ in range(2000)]
    cases += [int.from_bytes(rng.bytes(64), "little") >> int(s) for s in rng.integers(0, 300, 300)]
    cases += [k * L + d for k in (1, 2**100, 2**259 + 12345) for d in (-1, 0, 1)]
    for x in cases:
        x %= 2**512
        rem, div = _out(32), _out(40)
        hc.hc_divrem_l(_buf(x.to_bytes(64, "little")), rem, div)
        q, r = divmod(x, L)
        assert int.from_bytes(bytes(rem), "little") == r and int.from_bytes(bytes(div), "little") == q, hex(x)


def test_decompress_and_scalarmult(hc):
    from oracle import cbind as orc, pyoracle as po
    rng = np.random.default_rng(7)
    pts = []
    cands = [bytes(32), (1).to_bytes(32, "little"), (P - 1).to_bytes(32, "little"), (2**255 - 1).to_bytes(32, "little"),
             po.GY.to_bytes(32, "little"), (po.GY | (1 << 255)).to_bytes(32, "little"), (1 | (1 << 255)).to_bytes(32, "little"),
             (P + 1).to_bytes(32, "little")] + [rng.bytes(32) for _ in range(60)]
    n_ok = n_bad = 0
    for c in cands:
        xy, root = _out(64), _out(32)
        ok = hc.hc_decompress(_buf(c), xy, root)
        wxy, wroot, wok = orc.ed25519_decompress(c)
        assert bool(ok) == wok and bytes(xy) == wxy and bytes(root) == wroot, c.hex()
        (px, py), proot, pok = po.ed_decompress(c)
        assert pok == wok and px.to_bytes(32, "little") + py.to_bytes(32, "little") == wxy and proot.to_bytes(32, "little") == wroot
        if ok:
            n_ok += 1
            pts.append(bytes(xy))
        else:
            n_bad += 1
    assert n_ok > 20 and n_bad > 10
    scalars = [0, 1, 2, 15, 16, 17, L - 1, L, L + 1, 2**256 - 1, 2**255, 8, 2**252] + [int.from_bytes(rng.bytes(32), "little") for _ in range(12)]
    g = po.ed_point_bytes(po.G)
    for k in scalars:
        kb = k.to_bytes(32, "little")
        o = _out(64)
        hc.hc_scalarmult(_buf(kb), _buf(g), o, 1)
        assert bytes(o) == orc.ed25519_scalar_mul(kb, g), k
    for i, k in enumerate(scalars):
        kb = k.to_bytes(32, "little")
        pt = pts[i % len(pts)]
        o = _out(64)
        hc.hc_scalarmult(_buf(kb), _buf(pt), o, 0)
        assert bytes(o) == orc.ed25519_scalar_mul(kb, pt), (k, pt.hex())


def test_witness_records(hc):
    """Full per-signature record vs the C oracle and the Python-int oracle: valid signatures, the
    DUMMY triple, corrupted signatures (must not verify), s >= l, undecodable points."""
    from nacl.signing import SigningKey
    from oracle import cbind as orc, pyoracle as po
    rng = np.random.default_rng(8)
    cases = []
    for i in range(24):
        sk = SigningKey(hashlib.sha256(b"hc%d" % i).digest())
        msg = rng.bytes(int(rng.integers(0, 124)))
        sig = sk.sign(msg).signature
        cases.append((bytes(sk.verify_key), sig, msg))
    cases.append((po.DUMMY_PUBLIC_KEY, po.DUMMY_SIGNATURE, bytes(32)))
    pk, sig, msg = cases[0]
    cases.append((pk, sig[:5] + bytes([sig[5] ^ 1]) + sig[6:], msg))                 # R corrupted
    cases.append((pk, sig[:40] + bytes([sig[40] ^ 4]) + sig[41:], msg))              # s corrupted
    cases.append((pk, sig, msg + b"x"))                                              # message changed
    cases.append((pk, sig[:32] + (2**256 - 1).to_bytes(32, "little"), msg))          # s = 2^256-1
    cases.append((pk, sig[:32] + L.to_bytes(32, "little"), msg))                     # s = l
    cases.append(((2).to_bytes(32, "little"), sig, msg))                             # pk not on curve
    cases.append((pk, (2).to_bytes(32, "little") + sig[32:], msg))                   # R not on curve
    for j in range(6):
        cases.append((rng.bytes(32), rng.bytes(64), rng.bytes(50)))
    seen = set()
    for pk, sig, msg in cases:
        digest = hashlib.sha512(sig[:32] + pk + msg).digest()
        out = _out(576)
        hc.hc_ed25519_witness(_buf(pk), _buf(sig), _buf(digest), out)
        want = orc.ed25519_witness(pk, sig, msg)
        assert bytes(out) == want, (pk.hex(), sig.hex())
        if (want[520] & 6) == 6:
            assert want == po.ed_witness_bytes(pk, sig, msg)
        seen.add(want[520])
    assert 0xF in seen and 0x7 in seen and any(not (f & 1) for f in seen) and any(not (f & 2) for f in seen)


def test_no_limb_overflow_under_ubsan():
    """Runs the witness core in a child process against a -fsanitize=signed-integer-overflow build: a limb bound
    violated anywhere in the point formulas (19x terms, pair sums, 64-bit accumulators) aborts the child."""
    import sys
    import textwrap
    subprocess.check_call(["make", "-C", HERE, "-s", "libed_host_check_ubsan.so"], stderr=subprocess.DEVNULL)
    code = textwrap.dedent(f"""
        import ctypes as C, hashlib, sys
        import numpy as np
        sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
        from oracle import cbind as orc
        from nacl.signing import SigningKey
        hc = C.CDLL({os.path.join(HERE, "libed_host_check_ubsan.so")!r})
        buf = lambda b: (C.c_uint8 * len(b)).from_buffer_copy(bytes(b))
        rng = np.random.default_rng(3)
        for i in range(120):
            sk = SigningKey(hashlib.sha256(b"u%d" % i).digest())
            msg = rng.bytes(int(rng.integers(0, 124)))
            pk, sig = bytes(sk.verify_key), sk.sign(msg).signature
            if i % 5 == 4:
                pk, sig, msg = rng.bytes(32), rng.bytes(64), rng.bytes(40)
            out = (C.c_uint8 * 576)()
            hc.hc_ed25519_witness(buf(pk), buf(sig), buf(hashlib.sha512(sig[:32] + pk + msg).digest()), out)
            assert bytes(out) == orc.ed25519_witness(pk, sig, msg)
            if i % 3 == 0:      # the FP64-pipe build: its 64-bit column arithmetic under the same sanitizer
                out2 = (C.c_uint8 * 576)()
                hc.hc_fp64_ed25519_witness(buf(pk), buf(sig), buf(hashlib.sha512(sig[:32] + pk + msg).digest()), out2)
                assert bytes(out2) == bytes(out)
                out3 = (C.c_uint8 * 576)()   # and the per-key table path
                hc.hc_fp64_ed25519_witness_keyed(buf(pk), buf(sig), buf(hashlib.sha512(sig[:32] + pk + msg).digest()), out3)
                assert bytes(out3) == bytes(out)
        print("clean")
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "clean" in r.stdout, r.stderr[-2000:]


# ---------------------------------------------------------------------------------------------------------------
# the FP64-pipe field (blobstreamx_b200/csrc/fe51d.cuh): 5 x 51-bit balanced limbs in doubles, products split exactly by
# two round-toward-zero FMAs.  The host build runs fma() under FE_TOWARDZERO in place of __fma_rz.
# ---------------------------------------------------------------------------------------------------------------
def _val51(v):
    return sum(int(x) << (51 * i) for i, x in enumerate(v))


def test_fp64_field_at_limb_bounds(hc):
    """A product needs |f_i g_j| < 2^103: carried values (|limb| <= 2^50 + 2^14 = 1 unit) in the combinations the point
    formulas use -- 1x1, 2x2, 3x2, 2x3 and 1x7 units -- at the extremes (all-max, all-min, alternating) and at random;
    results are integer-valued, carried, and equal the product mod p."""
    rng = np.random.default_rng(12)
    D5 = C.c_double * 5
    U = 2**50 + 2**14

    def limbs(units, mode):
        m = units * U
        if mode == 0: return [m] * 5
        if mode == 1: return [-m] * 5
        if mode == 2: return [m if i % 2 else -m for i in range(5)]
        if mode == 3: return [-m if i % 2 else m for i in range(5)]
        return [int(rng.integers(-m, m + 1)) for _ in range(5)]
    worst = 0
    for it in range(6000):
        uf, ug = [(1, 1), (2, 2), (3, 2), (2, 3), (1, 7)][it % 5]
        f, g = limbs(uf, it % 7 if it < 200 else 9), limbs(ug, (it // 7) % 7 if it < 200 else 9)
        o = D5()
        for tw in (0, 1):
            hc.hc_fed_mul(D5(*map(float, f)), D5(*map(float, g)), o, tw)
            assert _val51(o) % P == ((1 + tw) * _val51(f) * _val51(g)) % P, (it, tw)
            assert all(float(x).is_integer() for x in o)
            worst = max(worst, max(abs(x) for x in o))
        if uf <= 2:
            for tw in (0, 1):
                hc.hc_fed_sq(D5(*map(float, f)), o, tw)
                assert _val51(o) % P == ((1 + tw) * _val51(f) ** 2) % P, (it, tw)
                worst = max(worst, max(abs(x) for x in o))
    assert worst <= U


def test_fp64_field_roundtrip_and_invert(hc):
    rng = np.random.default_rng(13)
    for a in _fe_cases(rng, 300):
        o = _out(32)
        hc.hc_fed_roundtrip(_buf(a.to_bytes(32, "little")), o)
        assert int.from_bytes(bytes(o), "little") == (a % 2**255) % P
    for a in _fe_cases(rng, 40):
        if (a % 2**255) % P == 0:
            continue
        o = _out(32)
        hc.hc_fed_invert(_buf(a.to_bytes(32, "little")), o)
        assert int.from_bytes(bytes(o), "little") == pow(a % 2**255, P - 2, P)


def test_fp64_decompress_and_scalarmult(hc):
    from oracle import cbind as orc, pyoracle as po
    rng = np.random.default_rng(14)
    pts = []
    cands = [bytes(32), (1).to_bytes(32, "little"), (P - 1).to_bytes(32, "little"), (2**255 - 1).to_bytes(32, "little"),
             po.GY.to_bytes(32, "little"), (po.GY | (1 << 255)).to_bytes(32, "little"), (1 | (1 << 255)).to_bytes(32, "little"),
             (P + 1).to_bytes(32, "little")] + [rng.bytes(32) for _ in range(60)]
    for c in cands:
        xy, root = _out(64), _out(32)
        ok = hc.hc_fp64_decompress(_buf(c), xy, root)
        wxy, wroot, wok = orc.ed25519_decompress(c)
        assert bool(ok) == wok and bytes(xy) == wxy and bytes(root) == wroot, c.hex()
        if ok:
            pts.append(bytes(xy))
    scalars = [0, 1, 2, 7, 8, 9, 15, 16, 17, L - 1, L, L + 1, 2**256 - 1, 2**255, 2**252, int("88" * 32, 16), int("77" * 32, 16)] + \
        [int.from_bytes(rng.bytes(32), "little") for _ in range(12)]
    g = po.ed_point_bytes(po.G)
    for i, k in enumerate(scalars):
        kb = k.to_bytes(32, "little")
        o = _out(64)
        hc.hc_fp64_scalarmult(_buf(kb), _buf(g), o, 1)
        assert bytes(o) == orc.ed25519_scalar_mul(kb, g), k
        pt = pts[i % len(pts)]
        hc.hc_fp64_scalarmult(_buf(kb), _buf(pt), o, 0)
        assert bytes(o) == orc.ed25519_scalar_mul(kb, pt), (k, pt.hex())


def test_fp64_witness_records(hc):
    """The full per-signature record of the FP64 build equals the oracle's: valid signatures, DUMMY, corrupted, garbage."""
    from nacl.signing import SigningKey
    from oracle import cbind as orc, pyoracle as po
    rng = np.random.default_rng(15)
    cases = [(po.DUMMY_PUBLIC_KEY, po.DUMMY_SIGNATURE, bytes(32))]
    for i in range(40):
        sk = SigningKey(hashlib.sha256(b"fp%d" % i).digest())
        msg = rng.bytes(int(rng.integers(0, 124)))
        pk, sig = bytes(sk.verify_key), sk.sign(msg).signature
        if i % 4 == 1:
            sig = sig[:7] + bytes([sig[7] ^ 2]) + sig[8:]
        if i % 4 == 2:
            sig = sig[:32] + (int.from_bytes(sig[32:], "little") + L).to_bytes(32, "little")
        if i % 8 == 7:
            pk, sig, msg = rng.bytes(32), rng.bytes(64), rng.bytes(33)
        cases.append((pk, sig, msg))
    # edges of the R shortcut (R' = sG - hA compared with the encoded R instead of a decompression): A = the identity makes
    # hA the identity, s = 0 makes sG the identity, so R' = (0, 1) -- encoded canonically, with the sign bit set on x = 0,
    # and non-canonically as y = 1 + p; the same with an undecodable public key (A falls back to the identity)
    ident, zero_s = (1).to_bytes(32, "little"), bytes(32)
    P25519 = 2**255 - 19
    for pk in (ident, (2).to_bytes(32, "little")):
        cases.append((pk, ident + zero_s, b"edge"))
        cases.append((pk, (1 | 1 << 255).to_bytes(32, "little") + zero_s, b"edge"))
        cases.append((pk, (1 + P25519).to_bytes(32, "little") + zero_s, b"edge"))
        cases.append((pk, (1 + P25519 | 1 << 255).to_bytes(32, "little") + zero_s, b"edge"))
    sk = SigningKey(hashlib.sha256(b"fp-sign").digest())
    sig = sk.sign(b"sign bit").signature
    cases.append((bytes(sk.verify_key), sig[:31] + bytes([sig[31] ^ 0x80]) + sig[32:], b"sign bit"))   # -R: decodes, does not verify
    seen = set()
    for pk, sig, msg in cases:
        out = _out(576)
        hc.hc_fp64_ed25519_witness(_buf(pk), _buf(sig), _buf(hashlib.sha512(sig[:32] + pk + msg).digest()), out)
        want = orc.ed25519_witness(pk, sig, msg)
        assert bytes(out) == want, (pk.hex(), sig.hex())
        out_k = _out(576)      # the per-key table path (h*A from the tabulated windows of the key) gives the same record
        hc.hc_fp64_ed25519_witness_keyed(_buf(pk), _buf(sig), _buf(hashlib.sha512(sig[:32] + pk + msg).digest()), out_k)
        assert bytes(out_k) == want, ("keyed", pk.hex(), sig.hex())
        seen.add(want[520])
    assert 0xF in seen and len(seen) >= 3

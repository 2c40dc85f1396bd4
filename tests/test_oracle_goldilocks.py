"""Oracle checks for hot path 3 (Goldilocks u32 gates + Poseidon): the C oracle vs an independent Python-int
restatement, the reference's property tests (valid witness => all constraints zero, corrupted => non-zero;
PX/frontend/uint/num/u32/gates/arithmetic_u32.rs:541-613 and siblings) and the reference's single Poseidon KAT
(PX/frontend/hash/poseidon/poseidon256.rs:172-178).  CPU only."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import cbind as orc

P = 2**64 - 2**32 + 1
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("gen_poseidon_constants", os.path.join(ROOT, "scripts", "gen_poseidon_constants.py"))
gpc = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gpc)


def prod_range(limb, base):
    p = 1
    for x in range(base):
        p = p * (limb - x) % P
    return p


def py_gate_eval(gate, p0, p1, w):
    """w: list of python ints (one row). Direct transcription of eval_unfiltered."""
    c = []
    if gate == orc.GATE_U32_ARITHMETIC:
        n = p0
        for i in range(n):
            m0, m1, ad, lo, hi, inv = w[6 * i:6 * i + 6]
            c.append((inv * (0xFFFFFFFF - hi) - 1) * lo % P)
            c.append((hi * 2**32 + lo - (m0 * m1 + ad)) % P)
            cl = ch = 0
            for j in reversed(range(32)):
                l = w[6 * n + 32 * i + j]
                c.append(prod_range(l, 4))
                if j < 16: cl = (4 * cl + l) % P
                else: ch = (4 * ch + l) % P
            c += [(cl - lo) % P, (ch - hi) % P]
    elif gate == orc.GATE_U32_ADD_MANY:
        na, n = p0, p1
        for i in range(n):
            b = (na + 3) * i
            comp = sum(w[b:b + na + 1])
            res, car = w[b + na + 1], w[b + na + 2]
            c.append((car * 2**32 + res - comp) % P)
            cr = cc = 0
            for j in reversed(range(19)):
                l = w[(na + 3) * n + 19 * i + j]
                c.append(prod_range(l, 4))
                if j < 16: cr = (4 * cr + l) % P
                else: cc = (4 * cc + l) % P
            c += [(cr - res) % P, (cc - car) % P]
    elif gate == orc.GATE_U32_SUBTRACTION:
        n = p0
        for i in range(n):
            x, y, bi, res, bo = w[5 * i:5 * i + 5]
            c.append((res - (x - y - bi + 2**32 * bo)) % P)
            comb = 0
            for j in reversed(range(16)):
                l = w[5 * n + 16 * i + j]
                c.append(prod_range(l, 4))
                comb = (4 * comb + l) % P
            c += [(comb - res) % P, bo * (1 - bo) % P]
    elif gate == orc.GATE_U32_COMPARISON:
        nb, nc = p0, p1
        cb = -(-nb // nc)
        f, s = w[4:4 + nc], w[4 + nc:4 + 2 * nc]
        c.append((sum(v * (1 << cb) ** i for i, v in enumerate(f)) - w[0]) % P)
        c.append((sum(v * (1 << cb) ** i for i, v in enumerate(s)) - w[1]) % P)
        msd = 0
        for i in range(nc):
            c += [prod_range(f[i], 1 << cb), prod_range(s[i], 1 << cb)]
            d, dum, eq, inter = (s[i] - f[i]) % P, w[4 + 2 * nc + i], w[4 + 3 * nc + i], w[4 + 4 * nc + i]
            c += [(d * dum - (1 - eq)) % P, eq * d % P, (inter - eq * msd) % P]
            msd = (inter + (1 - eq) * d) % P
        c.append((w[3] - msd) % P)
        bits = w[4 + 5 * nc:4 + 5 * nc + cb + 1]
        c += [b * (1 - b) % P for b in bits]
        c.append(((1 << cb) + w[3] - sum(b << i for i, b in enumerate(bits))) % P)
        c.append((w[2] - bits[cb]) % P)
    else:
        nl = p0
        for i in range(nl):
            aux = w[nl + 16 * i:nl + 16 * i + 16]
            c.append((sum(a * 4**j for j, a in enumerate(aux)) - w[i]) % P)
            c += [prod_range(a, 4) for a in aux]
    return c


GATES = [(orc.GATE_U32_ARITHMETIC, 3, 0), (orc.GATE_U32_ADD_MANY, 2, 5), (orc.GATE_U32_ADD_MANY, 4, 3), (orc.GATE_U32_ADD_MANY, 16, 3),
         (orc.GATE_U32_SUBTRACTION, 6, 0), (orc.GATE_U32_COMPARISON, 32, 16), (orc.GATE_U32_COMPARISON, 62, 31),
         (orc.GATE_U32_RANGE_CHECK, 7, 0), (orc.GATE_U32_RANGE_CHECK, 8, 0)]


def random_wires(rng, nw, rows):
    w = rng.integers(0, P, (nw, rows), dtype=np.uint64)
    w[:, 0] = [0, 1, P - 1, 2**32, 2**32 - 1, 3][0:1] * nw  # a zero row
    if rows > 1:
        w[:, 1] = P - 1
    return w


def valid_inputs(rng, gate, p0, p1, rows):
    """Random INPUT wires of valid u32 operations (the generators fill the rest)."""
    nw = orc.gate_num_wires(gate, p0, p1)
    w = np.zeros((nw, rows), np.uint64)
    u32 = lambda: rng.integers(0, 2**32, rows, dtype=np.uint64)
    if gate == orc.GATE_U32_ARITHMETIC:
        for i in range(p0):
            w[6 * i], w[6 * i + 1], w[6 * i + 2] = u32(), u32(), u32()
        w[0, 0], w[1, 0], w[2, 0] = 2**32 - 1, 2**32 - 1, 2**32 - 1      # high limb = u32::MAX - 1 ... edge
        if rows > 1:
            w[0, 1], w[1, 1], w[2, 1] = 0, 0, 0
    elif gate == orc.GATE_U32_ADD_MANY:
        for i in range(p1):
            for j in range(p0 + 1):
                w[(p0 + 3) * i + j] = u32()
    elif gate == orc.GATE_U32_SUBTRACTION:
        for i in range(p0):
            w[5 * i], w[5 * i + 1], w[5 * i + 2] = u32(), u32(), rng.integers(0, 2, rows, dtype=np.uint64)
    elif gate == orc.GATE_U32_COMPARISON:
        w[0], w[1] = rng.integers(0, 2**p0, rows, dtype=np.uint64), rng.integers(0, 2**p0, rows, dtype=np.uint64)
        w[1, 0] = w[0, 0]
    else:
        for i in range(p0):
            w[i] = u32()
    return w


@pytest.mark.parametrize("gate,p0,p1", GATES)
def test_gate_eval_vs_python(gate, p0, p1):
    rng = np.random.default_rng(100 + gate * 10 + p0)
    nw, ncn = orc.gate_num_wires(gate, p0, p1), orc.gate_num_constraints(gate, p0, p1)
    rows = 6
    w = random_wires(rng, nw, rows)
    got = orc.gate_eval(gate, p0, p1, w, threads=2)
    assert got.shape == (ncn, rows)
    for r in range(rows):
        want = py_gate_eval(gate, p0, p1, [int(x) % P for x in w[:, r]])
        assert len(want) == ncn
        assert [int(x) for x in got[:, r]] == want, (gate, r)


def test_gate_sizes_match_reference():
    # standard_recursion_config: 135 wires / 80 routed: arithmetic 3 ops, 108 constraints;
    # add_many(2) 5 ops 110; subtraction 6 ops 114
    assert orc.gate_num_wires(orc.GATE_U32_ARITHMETIC, 3) == 114 and orc.gate_num_constraints(orc.GATE_U32_ARITHMETIC, 3) == 108
    assert orc.gate_num_constraints(orc.GATE_U32_ADD_MANY, 2, 5) == 110 and orc.gate_num_wires(orc.GATE_U32_ADD_MANY, 2, 5) == 120
    assert orc.gate_num_constraints(orc.GATE_U32_SUBTRACTION, 6) == 114 and orc.gate_num_wires(orc.GATE_U32_SUBTRACTION, 6) == 126
    assert orc.gate_num_constraints(orc.GATE_U32_COMPARISON, 32, 16) == 6 + 80 + 2


@pytest.mark.parametrize("gate,p0,p1", GATES)
def test_gate_constraint_and_canonicity(gate, p0, p1):
    """test_gate_constraint: generator-filled rows satisfy every constraint; flipping any single wire breaks one."""
    rng = np.random.default_rng(200 + gate * 10 + p0)
    rows = 64
    w = orc.gate_witness(gate, p0, p1, valid_inputs(rng, gate, p0, p1, rows), threads=2)
    c = orc.gate_eval(gate, p0, p1, w, threads=2)
    assert not c.any()
    nw = w.shape[0]
    for trial in range(12):
        bad = w.copy()
        col, r = int(rng.integers(0, nw)), int(rng.integers(0, rows))
        bad[col, r] = (int(bad[col, r]) + 1 + int(rng.integers(0, 3))) % P
        cb = orc.gate_eval(gate, p0, p1, bad, threads=2)
        assert cb[:, r].any(), (gate, col)
        assert not np.delete(cb, r, axis=1).any()


def test_poseidon_constants_and_kat():
    rc = orc.poseidon_round_constants()
    assert [int(x) for x in rc] == gpc.round_constants()
    assert int(rc[0]) == 0xB585F766F2144405 and int(rc[11]) == 0xC54302F225DB2C76
    # the reference's only Poseidon KAT (byte variant with bit-reversed packing, poseidon256.rs:96-123,172-178)
    leaf = bytes.fromhex("d68d62c262c2ec08961c1104188cde86f51695878759666ad61490c8ec66745c")
    rev8 = lambda b: int("{:08b}".format(b)[::-1], 2)
    els = [sum(rev8(b) << (8 * j) for j, b in enumerate(leaf[i:i + 4])) for i in range(0, 32, 4)]
    h = orc.poseidon_hash_no_pad(np.array(els, np.uint64))
    out = b"".join(bytes(rev8((int(e) >> (8 * j)) & 0xFF) for j in range(8)) for e in h)
    assert out.hex() == "faa1095f1959da5713d6ad8b21b54936f167dc8e3f205b129b8eb8740aa10c0b"


def test_poseidon_vs_python():
    rng = np.random.default_rng(9)
    rc = gpc.round_constants()
    for n in (0, 1, 4, 7, 8, 9, 16, 17, 64, 128):
        x = rng.integers(0, P, n, dtype=np.uint64)
        assert [int(v) for v in orc.poseidon_hash_no_pad(x)] == gpc.hash_no_pad([int(v) for v in x], rc), n
    s = rng.integers(0, 2**64, 12, dtype=np.uint64)  # non-canonical inputs are reduced
    assert [int(v) for v in orc.poseidon_permute(s)] == gpc.permute([int(v) % P for v in s], rc)


def test_mapreduce_poseidon_tree():
    """mapreduce_merkle_tree_root (PX/utils/poseidon/mod.rs:9-66): leaves = hash of B U64 inputs (2 elements
    each, low limb first), then pairwise hashing to the root -- batch API vs a straight loop."""
    rng = np.random.default_rng(10)
    B, J = 32, 32
    blocks = rng.integers(0, 2**63, B * J, dtype=np.uint64)
    els = np.stack([blocks & np.uint64(0xFFFFFFFF), blocks >> np.uint64(32)], axis=1).reshape(-1)
    offs = np.arange(J + 1, dtype=np.uint32) * (2 * B)
    level = orc.poseidon_batch(els, offs, threads=2)
    rc = gpc.round_constants()
    assert [int(v) for v in level[3]] == gpc.hash_no_pad([int(v) for v in els[3 * 2 * B:4 * 2 * B]], rc)
    while len(level) > 1:
        flat = level.reshape(-1)
        level = orc.poseidon_batch(flat, np.arange(len(level) // 2 + 1, dtype=np.uint32) * 8)
    assert level.shape == (1, 4)

"""GPU parity for hot path 3: u32 gate constraint evaluation / witness generators and Poseidon vs the oracle."""
import numpy as np
import pytest

from tests.test_oracle_goldilocks import GATES, P, random_wires, valid_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from blobstreamx_b200 import lib
    c = lib.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def orc():
    from oracle import cbind
    return cbind


@pytest.mark.parametrize("gate,p0,p1", GATES)
@pytest.mark.parametrize("rows", [1, 1000])
def test_gate_eval_random_wires(ctx, orc, gate, p0, p1, rows):
    """test_eval_fns analogue: arbitrary (also non-canonical) wires, all constraints equal the oracle's."""
    rng = np.random.default_rng(300 + gate * 10 + p0 + rows)
    nw = orc.gate_num_wires(gate, p0, p1)
    assert ctx.gate_num_wires(gate, p0, p1) == nw and ctx.gate_num_constraints(gate, p0, p1) == orc.gate_num_constraints(gate, p0, p1)
    w = random_wires(rng, nw, rows)
    if rows > 4:
        w[:, 2] = np.uint64(2**64 - 1)       # non-canonical
        w[:, 3] = rng.integers(0, 4, nw, dtype=np.uint64)  # in-range limbs
    assert (ctx.gl_gate_eval(gate, p0, p1, w) == orc.gate_eval(gate, p0, p1, w, threads=8)).all()


@pytest.mark.parametrize("gate,p0,p1", GATES)
def test_gate_eval_edge_values(ctx, orc, gate, p0, p1):
    """Wires drawn from the corners of the field and of the 64-bit range (0, 1, 2, 3, 4, p-1, p, p+1, 2^32-1, 2^32,
    2^32+1, 2^63, 2^64-2^32, 2^64-1, ...) mixed with random values: every carry / borrow branch of the weak reductions."""
    rng = np.random.default_rng(900 + gate * 10 + p0)
    nw = orc.gate_num_wires(gate, p0, p1)
    corners = np.array([0, 1, 2, 3, 4, 5, P - 1, P - 2, P - 3, P, P + 1, P + 2, 2**32 - 1, 2**32, 2**32 + 1, 2**33 - 1, 2**63, 2**63 - 1,
                        2**64 - 2**32, 2**64 - 2**32 - 1, 2**64 - 2**32 + 2, 2**64 - 1, 2**64 - 2, 2**64 - 4, 0xFFFFFFFE00000001,
                        0xFFFFFFFF, 0xFFFFFFFF00000000, 0x00000001FFFFFFFF, 0x8000000080000000], dtype=np.uint64)
    rows = 4096
    w = corners[rng.integers(0, len(corners), (nw, rows))]
    mix = rng.random((nw, rows)) < 0.25
    w[mix] = rng.integers(0, 2**64, int(mix.sum()), dtype=np.uint64)
    assert (ctx.gl_gate_eval(gate, p0, p1, w) == orc.gate_eval(gate, p0, p1, w, threads=8)).all()


def test_poseidon_edge_values(ctx, orc):
    rng = np.random.default_rng(77)
    corners = np.array([0, 1, P - 1, P, P + 1, 2**32 - 1, 2**32, 2**64 - 2**32, 2**64 - 1, 2**63, 0xFFFFFFFF00000000], dtype=np.uint64)
    n, L_ = 3000, 12
    x = corners[rng.integers(0, len(corners), n * L_)]
    offs = (np.arange(n + 1) * L_).astype(np.uint32)
    assert (ctx.gl_poseidon_batch(x, offs) == orc.poseidon_batch(x, offs, threads=8)).all()


@pytest.mark.parametrize("gate,p0,p1", GATES)
def test_gate_witness_and_zero_constraints(ctx, orc, gate, p0, p1):
    """test_gate_constraint analogue on the GPU: generator kernel output == oracle generator, and it satisfies
    every constraint; a corrupted wire breaks at least one (test_canonicity)."""
    rng = np.random.default_rng(400 + gate * 10 + p0)
    rows = 4096 + 37
    inp = valid_inputs(rng, gate, p0, p1, rows)
    w = ctx.gl_gate_witness(gate, p0, p1, inp)
    assert (w == orc.gate_witness(gate, p0, p1, inp, threads=8)).all()
    c = ctx.gl_gate_eval(gate, p0, p1, w)
    assert not c.any()
    bad = w.copy()
    bad[w.shape[0] - 1, 5] = (int(bad[w.shape[0] - 1, 5]) + 1) % P
    cb = ctx.gl_gate_eval(gate, p0, p1, bad)
    assert cb[:, 5].any() and not np.delete(cb, 5, axis=1).any()


def test_gate_eval_large_property(ctx):
    """BASELINE-size run (2^18 rows of U32Arithmetic): valid witnesses give all-zero constraints."""
    from oracle import cbind as orc
    rng = np.random.default_rng(5)
    rows = 1 << 18
    w = ctx.gl_gate_witness(orc.GATE_U32_ARITHMETIC, 3, 0, valid_inputs(rng, orc.GATE_U32_ARITHMETIC, 3, 0, rows))
    assert not ctx.gl_gate_eval(orc.GATE_U32_ARITHMETIC, 3, 0, w).any()


def test_poseidon_batch(ctx, orc):
    rng = np.random.default_rng(12)
    lens = [0, 1, 4, 7, 8, 9, 15, 16, 17, 64, 128, 8, 8, 8] + [int(x) for x in rng.integers(0, 40, 300)]
    offs = np.zeros(len(lens) + 1, np.uint32)
    offs[1:] = np.cumsum(lens)
    x = rng.integers(0, 2**64, int(offs[-1]), dtype=np.uint64)   # includes non-canonical values
    assert (ctx.gl_poseidon_batch(x, offs) == orc.poseidon_batch(x, offs, threads=8)).all()
    # reference KAT through the GPU (poseidon256.rs:172-178)
    leaf = bytes.fromhex("d68d62c262c2ec08961c1104188cde86f51695878759666ad61490c8ec66745c")
    rev8 = lambda b: int("{:08b}".format(b)[::-1], 2)
    els = np.array([sum(rev8(b) << (8 * j) for j, b in enumerate(leaf[i:i + 4])) for i in range(0, 32, 4)], np.uint64)
    h = ctx.gl_poseidon_batch(els, np.array([0, 8], np.uint32))[0]
    out = b"".join(bytes(rev8((int(e) >> (8 * j)) & 0xFF) for j in range(8)) for e in h)
    assert out.hex() == "faa1095f1959da5713d6ad8b21b54936f167dc8e3f205b129b8eb8740aa10c0b"


def test_mapreduce_poseidon_tree(ctx, orc):
    """mapreduce_merkle_tree_root over 32 jobs x 32 U64 inputs (PX/utils/poseidon/mod.rs:9-66)."""
    rng = np.random.default_rng(13)
    B, J = 32, 32
    blocks = rng.integers(0, 2**63, B * J, dtype=np.uint64)
    els = np.stack([blocks & np.uint64(0xFFFFFFFF), blocks >> np.uint64(32)], axis=1).reshape(-1)
    offs = np.arange(J + 1, dtype=np.uint32) * (2 * B)
    g, o = ctx.gl_poseidon_batch(els, offs), orc.poseidon_batch(els, offs)
    assert (g == o).all()
    while len(g) > 1:
        offs = np.arange(len(g) // 2 + 1, dtype=np.uint32) * 8
        g, o = ctx.gl_poseidon_batch(g.reshape(-1), offs), orc.poseidon_batch(o.reshape(-1), offs)
        assert (g == o).all()

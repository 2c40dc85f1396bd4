"""The CPU restatement of the prover inner loops (oracle/plonk.c; plonky2 is un-vendored, so PARITY UNPINNED against its
vectors) pinned by what the algorithms must satisfy: the transform equals the naive DFT with the documented root of unity,
inverse(forward) is the identity, the coset extension equals direct polynomial evaluation at shift * w_N^k, the Merkle
tree is built from the KAT-pinned Poseidon primitive, and a FRI fold evaluates as sum_j beta^j f_j(y)."""
import numpy as np
import pytest

from oracle import cbind as orc

P = orc.GL_P


def test_field_constants_are_self_consistent():
    g, w32 = orc.gl_coset_shift(), orc.gl_root_of_unity(32)
    assert pow(g, (P - 1) >> 32, P) == w32 and pow(w32, 1 << 31, P) == P - 1
    assert all(pow(g, (P - 1) // q, P) != 1 for q in (2, 3, 5, 17, 257, 65537))       # g generates the whole group
    for k in (1, 3, 10, 20):
        w = orc.gl_root_of_unity(k)
        assert pow(w, 1 << k, P) == 1 and pow(w, 1 << (k - 1), P) == P - 1
        assert pow(orc.gl_root_of_unity(k + 1), 2, P) == w


@pytest.mark.parametrize("log_n", [1, 2, 4, 6])
def test_ntt_equals_naive_dft(log_n):
    rng = np.random.default_rng(log_n)
    n = 1 << log_n
    x = [int(v) % P for v in rng.integers(0, 2**64, n, dtype=np.uint64)]
    w = orc.gl_root_of_unity(log_n)
    want = [sum(x[j] * pow(w, j * k, P) for j in range(n)) % P for k in range(n)]
    got = orc.gl_ntt(np.array(x, np.uint64))
    assert got.tolist() == want
    assert orc.gl_ntt(got, inverse=True).tolist() == x


def test_ntt_round_trip_and_linearity_large():
    rng = np.random.default_rng(3)
    a = rng.integers(0, P, (3, 1 << 12), dtype=np.uint64)
    fa = orc.gl_ntt(a)
    assert (orc.gl_ntt(fa, inverse=True) == a).all()
    s = ((a[0].astype(object) + a[1].astype(object)) % P).astype(np.uint64)
    assert (orc.gl_ntt(s) == ((fa[0].astype(object) + fa[1].astype(object)) % P).astype(np.uint64)).all()


@pytest.mark.parametrize("log_n,rate_bits", [(3, 3), (5, 1), (6, 3)])
def test_lde_is_polynomial_evaluation_on_the_coset(log_n, rate_bits):
    rng = np.random.default_rng(7)
    n, N = 1 << log_n, 1 << (log_n + rate_bits)
    c = [int(v) % P for v in rng.integers(0, 2**64, n, dtype=np.uint64)]
    got = orc.gl_lde(np.array(c, np.uint64), rate_bits)
    g, wN = orc.gl_coset_shift(), orc.gl_root_of_unity(log_n + rate_bits)
    for k in list(range(8)) + [N - 1, N // 2 + 3]:
        x = g * pow(wN, k, P) % P
        assert int(got[k]) == sum(cj * pow(x, j, P) for j, cj in enumerate(c)) % P
    # the extension of a low-degree polynomial restricted to every 2^rate_bits-th point is a size-n coset transform
    assert int(got[0]) == sum(cj * pow(g, j, P) for j, cj in enumerate(c)) % P


def test_merkle_tree_from_the_pinned_primitive():
    rng = np.random.default_rng(9)
    leaves = rng.integers(0, P, (16, 11), dtype=np.uint64)
    d = orc.gl_merkle(leaves, cap_height=1)
    assert d.shape == (16 + 8 + 4 + 2, 4)
    for i in (0, 7, 15):
        assert (d[i] == orc.poseidon_hash_no_pad(leaves[i])).all()
    # two_to_one(l, r) = hash_n_to_hash_no_pad(l ‖ r): one permutation of (l, r, 0, 0, 0, 0) either way
    assert (d[16] == orc.poseidon_hash_no_pad(np.concatenate([d[0], d[1]]))).all()
    assert (d[16 + 8 + 4 + 1] == orc.poseidon_hash_no_pad(np.concatenate([d[16 + 8 + 2], d[16 + 8 + 3]]))).all()
    short = rng.integers(0, P, (4, 3), dtype=np.uint64)            # hash_or_noop: <= 4 elements are padded, not hashed
    ds = orc.gl_merkle(short, cap_height=2)
    assert (ds[:, :3] == short).all() and not ds[:, 3].any()


def _ext_mul(a, b):
    return ((a[0] * b[0] + 7 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


@pytest.mark.parametrize("arity_bits", [1, 4])
def test_fri_fold_identity(arity_bits):
    """f(x) = sum_j x^j f_j(x^arity)  =>  the folded polynomial is g(y) = sum_j beta^j f_j(y): checked at a point."""
    rng = np.random.default_rng(11)
    arity, n = 1 << arity_bits, 64
    f = rng.integers(0, P, (n, 2), dtype=np.uint64)
    beta = (int(rng.integers(0, P, dtype=np.uint64)), int(rng.integers(0, P, dtype=np.uint64)))
    g = orc.gl_fri_fold(f, arity_bits, beta)
    y = (123456789, 987654321)
    gy, yp = (0, 0), (1, 0)
    for i in range(n // arity):
        t = _ext_mul((int(g[i][0]), int(g[i][1])), yp)
        gy = ((gy[0] + t[0]) % P, (gy[1] + t[1]) % P)
        yp = _ext_mul(yp, y)
    want, bp = (0, 0), (1, 0)
    for j in range(arity):
        fj, yp = (0, 0), (1, 0)
        for i in range(n // arity):
            t = _ext_mul((int(f[i * arity + j][0]), int(f[i * arity + j][1])), yp)
            fj = ((fj[0] + t[0]) % P, (fj[1] + t[1]) % P)
            yp = _ext_mul(yp, y)
        t = _ext_mul(fj, bp)
        want = ((want[0] + t[0]) % P, (want[1] + t[1]) % P)
        bp = _ext_mul(bp, beta)
    assert gy == want


def test_quotient_combine_vanishes_on_a_valid_trace():
    """Constraints of a valid U32Arithmetic trace are zero on the subgroup, so the alpha-combination is zero there."""
    rng = np.random.default_rng(13)
    rows = 64
    w = np.zeros((orc.gate_num_wires(orc.GATE_U32_ARITHMETIC, 3, 0), rows), np.uint64)
    for i in range(3):
        w[6 * i:6 * i + 3] = rng.integers(0, 2**32, (3, rows), dtype=np.uint64)
    w = orc.gate_witness(orc.GATE_U32_ARITHMETIC, 3, 0, w)
    c = orc.gate_eval(orc.GATE_U32_ARITHMETIC, 3, 0, w)
    q = orc.gl_quotient_combine(c, [12345, 67890], np.ones(rows, np.uint64))
    assert not q.any()
    w[1, 5] ^= 1
    q = orc.gl_quotient_combine(orc.gate_eval(orc.GATE_U32_ARITHMETIC, 3, 0, w), [12345, 67890], np.ones(rows, np.uint64))
    assert q[:, 5].all() and not np.delete(q, 5, axis=1).any()

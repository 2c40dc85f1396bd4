"""GPU parity for attestation proofs (bsx_attestation_proofs): for every height of trees with 1 .. N leaves the side
nodes equal the aunts of the oracle's generic Tendermint proof generator (TX/input/tendermint_utils.rs:276-336) over the
data-root tuples, rebuild the data commitment, and reproduce the fixture commitment of mocha-4 10000 -> 10004."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from blobstreamx_b200 import lib
    c = lib.Context(0)
    yield c
    c.close()


def _tuple(height: int, data_hash: bytes) -> bytes:
    return bytes(24) + int(height).to_bytes(8, "big") + data_hash          # BX/circuits/builder.rs:82-103


def _root_from_aunts(index, total, leaf_hash, aunts):
    """compute_hash_from_aunts, TX/input/tendermint_utils.rs:225-273"""
    if total == 1:
        assert not aunts
        return leaf_hash
    k = 1
    while k * 2 < total:
        k *= 2
    inner = lambda l, r: hashlib.sha256(b"\x01" + l + r).digest()
    if index < k:
        return inner(_root_from_aunts(index, k, leaf_hash, aunts[:-1]), aunts[-1])
    return inner(aunts[-1], _root_from_aunts(index - k, total - k, leaf_hash, aunts[:-1]))


@pytest.mark.parametrize("N", [1, 2, 8, 13, 32])
def test_attestation_every_leaf_count(ctx, N):
    from oracle import cbind as orc
    rng = np.random.default_rng(40 + N)
    t = N + 1                                        # tree k has k leaves (k = 0 .. N)
    dh = rng.integers(0, 256, (t, N, 32), dtype=np.uint8)
    start = (5_000_000 + 100 * np.arange(t)).astype(np.uint64)
    end = start + np.arange(t, dtype=np.uint64)
    q_tree = np.repeat(np.arange(t, dtype=np.uint32), N + 2)
    q_height = (start[q_tree].astype(np.int64) + np.tile(np.arange(-1, N + 1), t)).astype(np.uint64)   # one below, all, one above
    g = ctx.attestation_proofs(dh, N, start, end, q_tree, q_height)
    _, roots, _ = ctx.data_commitment_batch(dh, N, start, end)
    assert (g["roots"] == roots).all()
    for q, (tr, h) in enumerate(zip(q_tree, q_height)):
        n, i = int(end[tr] - start[tr]), int(h) - int(start[tr])
        assert g["num_leaves"][q] == n
        if not 0 <= i < n:
            assert g["key"][q] == 0xFFFFFFFF and g["depth"][q] == 0 and not g["side_nodes"][q].any()
            continue
        tuples = [_tuple(int(start[tr]) + k, dh[tr, k].tobytes()) for k in range(n)]
        aunts, root = orc.tm_aunts_from_slices(tuples, i)
        d = int(g["depth"][q])
        assert g["key"][q] == i and d == len(aunts)
        assert (g["side_nodes"][q, :d] == aunts).all() and not g["side_nodes"][q, d:].any()
        assert root == roots[tr].tobytes()
        leaf = hashlib.sha256(b"\x00" + tuples[i]).digest()
        assert _root_from_aunts(i, n, leaf, [a.tobytes() for a in aunts]) == root


def test_attestation_fixture_and_full_size(ctx, golden):
    """mocha-4 10000 -> 10004 (data commitment 5F1B8536...), and a 2048-leaf tree queried at every height."""
    from oracle import cbind as orc
    hs = [bytes.fromhex(golden["headers"][str(h)]["data_hash"]) for h in range(10000, 10004)]
    dh = np.zeros((1, 4, 32), np.uint8)
    for k, b in enumerate(hs):
        dh[0, k] = np.frombuffer(b, np.uint8)
    g = ctx.attestation_proofs(dh, 4, [10000], [10004], np.zeros(4, np.uint32), np.arange(10000, 10004, dtype=np.uint64))
    assert g["roots"][0].tobytes().hex().upper().startswith("5F1B8536")
    for i in range(4):
        leaf = hashlib.sha256(b"\x00" + _tuple(10000 + i, hs[i])).digest()
        assert _root_from_aunts(i, 4, leaf, [a.tobytes() for a in g["side_nodes"][i, :g["depth"][i]]]) == g["roots"][0].tobytes()
    rng = np.random.default_rng(41)
    N, n = 2048, 1999
    dh = rng.integers(0, 256, (1, N, 32), dtype=np.uint8)
    g = ctx.attestation_proofs(dh, N, [7_000_000], [7_000_000 + n], np.zeros(n, np.uint32), np.arange(7_000_000, 7_000_000 + n, dtype=np.uint64))
    tuples = [_tuple(7_000_000 + k, dh[0, k].tobytes()) for k in range(n)]
    for i in list(range(0, n, 97)) + [n - 1, n - 2, 1023, 1024, 1535, 1536]:
        aunts, root = orc.tm_aunts_from_slices(tuples, i)
        d = int(g["depth"][i])
        assert d == len(aunts) and (g["side_nodes"][i, :d] == aunts).all() and root == g["roots"][0].tobytes()

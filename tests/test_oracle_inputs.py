"""Input shaping on the CPU: the oracle's generic Tendermint proof generator and range-input restatement
against the reference's fixtures (header hash chain of mocha-4 10000-10004, data commitments) and against the host
shaper (blobstreamx_b200/inputs.py, the mirror of BX/circuits/input.rs)."""
import numpy as np
import pytest

from blobstreamx_b200 import inputs as I
from blobstreamx_b200 import synthetic as S
from oracle import cbind as orc

FIELDS = ("dh_leaf", "dh_aunts", "lb_leaf", "lb_aunts", "start_headers", "end_headers", "start_header", "end_header")


def _same(o, m):
    for k in FIELDS:
        assert (o[k].reshape(-1) == getattr(m, k).reshape(-1)).all(), k


def test_aunts_match_fixture_header_hashes(golden):
    """Every leaf's aunts rebuild the header hash the NEXT header commits to (last_block_id.hash), with the path of
    compute_hash_from_aunts (TX/input/tendermint_utils.rs:225-273)."""
    for h in ("10000", "10001", "10002", "10003"):
        leaves = I.header_leaves(golden["headers"][h])
        want = bytes.fromhex(golden["headers"][str(int(h) + 1)]["last_block_id"]["hash"])
        assert orc.tm_root_from_slices(leaves) == want
        for idx in range(14):
            aunts, root = orc.tm_aunts_from_slices(leaves, idx)
            assert root == want and len(aunts) == (4 if idx < 12 else 3)
            if idx < 12:
                dig, r2 = orc.tm_merkle_proof(leaves[idx], aunts.tobytes(), 4, idx)
                assert r2 == want


def test_range_inputs_fixture(golden):
    """10000 -> 10004 in the shapes of the reference's small test circuits (BX/circuits/header_range.rs:193-214)."""
    trees = {int(k): I.HeaderTree.build(I.header_leaves(v)) for k, v in golden["headers"].items()}
    for (a, b, J, B) in ((10000, 10004, 2, 4), (10000, 10004, 4, 2), (10002, 10004, 2, 2), (10000, 10001, 2, 4)):
        m = I.get_header_range_map_inputs(trees, a, b, J, B)          # the fixture chain goes on to 10004: latest = 10004
        o = orc.header_range_inputs(J, B, I.pack_range_headers(trees, a, J, B), a, b, min(10004, a + J * B))
        assert o["bad"] == 0
        _same(o, m)
        w = orc.prove_data_commitment(J, B, *[o[k] for k in FIELDS[:6]], a, o["start_header"], b, o["end_header"])
        assert w["fail"] == 0
    # the shaped inputs of 10000 -> 10004 prove the fixture's data commitment
    assert w is not None
    o = orc.header_range_inputs(2, 4, I.pack_range_headers(trees, 10000, 2, 4), 10000, 10004)
    w = orc.prove_data_commitment(2, 4, *[o[k] for k in FIELDS[:6]], 10000, o["start_header"], 10004, o["end_header"])
    assert w["data_commitment"].hex().upper().startswith("5F1B8536")


@pytest.mark.parametrize("J,B,nb", [(4, 8, None), (4, 8, 19), (2, 4, 1), (8, 4, 32), (4, 4, 5)])
def test_range_inputs_synthetic(J, B, nb):
    m, _, chain = S.header_range_inputs(J, B, nb, with_skip=False)
    o = orc.header_range_inputs(J, B, I.pack_range_headers(chain.trees, m.start_block, J, B), m.start_block, m.end_block)
    assert o["bad"] == 0
    _same(o, m)


def test_range_inputs_wrong_leaf_size():
    m, _, chain = S.header_range_inputs(2, 4, None, with_skip=False)
    rec = I.pack_range_headers(chain.trees, m.start_block, 2, 4)
    leaves = list(chain.trees[m.start_block + 2].leaves)
    leaves[6] = leaves[6][:-1]                       # a 33-byte data_hash field
    rec[2] = I.pack_header_record(leaves)
    assert orc.header_range_inputs(2, 4, rec, m.start_block, m.end_block)["bad"] == 1


@pytest.mark.parametrize("J,B,nb,extra", [(4, 8, 19, 13), (4, 8, 19, 40), (2, 4, 1, 7), (8, 4, 12, 3), (4, 4, 5, 0), (4, 4, 16, 9)])
def test_range_inputs_chain_longer_than_the_range(J, B, nb, extra):
    """The hint of a map job is clamped to the last fetchable block (latest_block - 2, BX/circuits/input.rs:160-163), NOT to
    the range's end: when the chain goes on past the target, the job holding the end and the jobs after it carry real
    proofs and headers.  Oracle == host shaper; the circuit's outputs (subchain records, commitment) are what the
    chain-tip case gives, only the witness of the disabled slots differs."""
    m, _, chain = S.header_range_inputs(J, B, nb, with_skip=False, extra_blocks=extra)
    latest = min(m.start_block + nb + extra, m.start_block + J * B)
    trees = {k: v for k, v in chain.trees.items() if k <= latest}
    rec = I.pack_range_headers(trees, m.start_block, J, B)
    o = orc.header_range_inputs(J, B, rec, m.start_block, m.end_block, latest)
    assert o["bad"] == 0
    _same(o, m)
    tip = orc.header_range_inputs(J, B, rec, m.start_block, m.end_block)          # as if the chain ended at the target
    w = orc.prove_data_commitment(J, B, *[o[k] for k in FIELDS[:6]], m.start_block, o["start_header"], m.end_block, o["end_header"])
    wt = orc.prove_data_commitment(J, B, *[tip[k] for k in FIELDS[:6]], m.start_block, tip["start_header"], m.end_block, tip["end_header"])
    assert w["fail"] == 0 == wt["fail"] and w["data_commitment"] == wt["data_commitment"]
    assert (w["reduce_nodes"] == wt["reduce_nodes"]).all()
    if extra and nb < J * B:
        assert not (o["lb_leaf"] == tip["lb_leaf"]).all()      # real proofs where the chain-tip case has zeros
        assert int(o["dh_leaf"].reshape(J * B, 34).any(axis=1).sum()) == latest - m.start_block   # data_hash proofs of [start, latest)

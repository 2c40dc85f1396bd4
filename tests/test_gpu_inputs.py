"""GPU parity for input shaping (bsx_header_trees, bsx_header_range_inputs): header hashes and map-circuit proofs built
on the device from encoded header records equal the oracle's, reproduce the fixture hash chain, and feed
prove_data_commitment to the fixture's data commitment."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
FIELDS = ("dh_leaf", "dh_aunts", "lb_leaf", "lb_aunts", "start_headers", "end_headers", "start_header", "end_header")


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


def test_header_trees_fixture_chain(ctx, orc, golden):
    from blobstreamx_b200 import inputs as I
    hs = ["10000", "10001", "10002", "10003", "10004"]
    leaves = [I.header_leaves(golden["headers"][h]) for h in hs]
    rec = np.stack([I.pack_header_record(l) for l in leaves])
    roots, lv = ctx.header_trees(rec, levels=True)
    assert roots[0].tobytes().hex().upper() == "A0123D5E4B8B8888A61F931EE2252D83568B97C223E0ECA9795B29B8BD8CBA2D"
    for i in range(4):
        assert roots[i].tobytes() == bytes.fromhex(golden["headers"][hs[i + 1]]["last_block_id"]["hash"])
    for i, l in enumerate(leaves):
        assert roots[i].tobytes() == orc.tm_root_from_slices(l)
        for idx in range(12):                        # the aunts the device would emit for any provable leaf
            aunts, _ = orc.tm_aunts_from_slices(l, idx)
            got = [lv[i, idx ^ 1], lv[i, 14 + ((idx >> 1) ^ 1)], lv[i, 21 + ((idx >> 2) ^ 1)], lv[i, 25 + ((idx >> 3) ^ 1)]]
            assert (np.stack(got) == aunts).all()


def test_range_inputs_fixture_to_data_commitment(ctx, orc, golden):
    from blobstreamx_b200 import inputs as I
    trees = {int(k): I.HeaderTree.build(I.header_leaves(v)) for k, v in golden["headers"].items()}
    J, B = 2, 4
    rec = I.pack_range_headers(trees, 10000, J, B)
    g = ctx.header_range_inputs(rec[None], [10000], [10004], J, B)
    o = orc.header_range_inputs(J, B, rec, 10000, 10004)
    assert g["fail"][0] == 0
    for k in FIELDS:
        assert (g[k][0].reshape(-1) == o[k].reshape(-1)).all(), k
    w = ctx.prove_data_commitment(1, J, B, *[g[k][0] for k in FIELDS[:6]], np.array([10000], np.uint64), g["start_header"][0],
                                  np.array([10004], np.uint64), g["end_header"][0])
    assert w["fail"][0] == 0 and w["data_commitments"][0].tobytes().hex().upper().startswith("5F1B8536")


def test_range_inputs_synthetic_batch(ctx, orc):
    """Several ranges in one call: full, partially filled (dummy proofs and dummy jobs), a single block, one with a
    wrong-size field (flagged, the others untouched)."""
    from blobstreamx_b200 import inputs as I
    from blobstreamx_b200 import synthetic as S
    J, B = 4, 8
    fills = (None, 19, 1, 8, 25, 32, 3)
    sets = [S.header_range_inputs(J, B, nb, start=7_000_000 + 100 * r, seed=S.SEED + 3 * r, with_skip=False) for r, nb in enumerate(fills)]
    recs = np.stack([I.pack_range_headers(c.trees, m.start_block, J, B) for m, _, c in sets])
    leaves = list(sets[4][2].trees[sets[4][0].start_block + 5].leaves)
    leaves[4] = leaves[4] + b"\x00"                  # a 73-byte last_block_id field in range 4
    recs[4, 5] = I.pack_header_record(leaves)
    sb = np.array([m.start_block for m, _, _ in sets], np.uint64)
    eb = np.array([m.end_block for m, _, _ in sets], np.uint64)
    g = ctx.header_range_inputs(recs, sb, eb, J, B)
    for r, (m, _, _) in enumerate(sets):
        o = orc.header_range_inputs(J, B, recs[r], m.start_block, m.end_block)
        assert (g["fail"][r] != 0) == (o["bad"] != 0) == (r == 4)
        for k in FIELDS:
            if r == 4 and k in ("lb_leaf", "lb_aunts", "dh_aunts", "start_headers", "end_headers"):
                continue                              # the flagged range's content is unspecified around the bad header
            assert (g[k][r].reshape(-1) == o[k].reshape(-1)).all(), (r, k)


def test_range_inputs_full_size_property(ctx):
    """header_range_1024 shape (32 x 32): the device-shaped inputs of a synthetic chain prove without assertion failures
    and give the same commitment as the host-shaped ones."""
    from blobstreamx_b200 import inputs as I
    from blobstreamx_b200 import synthetic as S
    J, B = 32, 32
    m, _, chain = S.header_range_inputs(J, B, 700, with_skip=False)
    rec = I.pack_range_headers(chain.trees, m.start_block, J, B)
    g = ctx.header_range_inputs(rec[None], [m.start_block], [m.end_block], J, B)
    assert g["fail"][0] == 0
    for k in FIELDS:
        assert (g[k][0].reshape(-1) == getattr(m, k).reshape(-1)).all(), k
    a = ctx.prove_data_commitment(1, J, B, *[g[k][0] for k in FIELDS[:6]], np.array([m.start_block], np.uint64), g["start_header"][0],
                                  np.array([m.end_block], np.uint64), g["end_header"][0])
    assert a["fail"][0] == 0


@pytest.mark.parametrize("extra", [0, 5, 64])
def test_range_inputs_latest_block_beyond_the_range_end(ctx, orc, extra):
    """latest_blocks: jobs are clamped to the last fetchable block, not to the range's end (BX/circuits/input.rs:160-163) --
    device == oracle == host shaper for a chain that goes on past the target, several fills in one call."""
    from blobstreamx_b200 import inputs as I
    from blobstreamx_b200 import synthetic as S
    J, B = 4, 8
    fills = (19, 1, 8, 25, 32, 3)
    sets = [S.header_range_inputs(J, B, nb, start=8_000_000 + 100 * r, seed=S.SEED + 5 * r, with_skip=False, extra_blocks=extra)
            for r, nb in enumerate(fills)]
    recs = np.stack([I.pack_range_headers(c.trees, m.start_block, J, B) for m, _, c in sets])
    sb = np.array([m.start_block for m, _, _ in sets], np.uint64)
    eb = np.array([m.end_block for m, _, _ in sets], np.uint64)
    lt = np.array([m.start_block + nb + extra for (m, _, _), nb in zip(sets, fills)], np.uint64)
    g = ctx.header_range_inputs(recs, sb, eb, J, B, latest_blocks=lt)
    assert not g["fail"].any()
    for r, (m, _, _) in enumerate(sets):
        o = orc.header_range_inputs(J, B, recs[r], m.start_block, m.end_block, int(lt[r]))
        for k in FIELDS:
            assert (g[k][r].reshape(-1) == o[k].reshape(-1)).all(), (r, k)
            assert (g[k][r].reshape(-1) == getattr(m, k).reshape(-1)).all(), (r, k)
    a = ctx.prove_data_commitment(len(fills), J, B, *[g[k] for k in FIELDS[:6]], sb, g["start_header"], eb, g["end_header"])
    assert not a["fail"].any()

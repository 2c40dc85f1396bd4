"""GPU parity for verify_header / verify_skip / next_header (through the C ABI) vs the oracle: every
SHA-256 digest in Curta request order, every Ed25519 record, the assertion masks and outputs."""
import copy

import numpy as np
import pytest

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


def _same(got, want, i=0):
    assert got["fail"][i] == want["fail"], (hex(got["fail"][i]), hex(want["fail"]))
    assert (got["sha256_digests"][i] == want["sha256_digests"]).all()
    assert (got["ed"][i] == want["ed"]).all()


def test_next_header_fixture(ctx, orc, golden):
    """BASELINE config 1: next_header 10000 -> 10001 on the reference's own fixture (2 validators padded to 100)."""
    from blobstreamx_b200 import inputs as I
    k = I.get_step_inputs(golden["headers"]["10000"], golden["headers"]["10001"], golden["commits"]["10001"],
                          golden["validators"]["10001"])
    got = ctx.next_header([k])
    want = orc.next_header(k, threads=8)
    _same(got, want)
    assert got["fail"][0] == 0 and got["sha256_digests"].shape[1] == 282
    assert got["data_commitments"][0].tobytes() == want["data_commitment"]
    assert want["data_commitment"].hex().upper() == golden["data_commitments"]["10000-10001"]
    # tampered witnesses give the oracle's assertion mask
    for field, off in (("last_block_id_proof", 5), ("prev_next_validators_proof", 40), ("data_hash_proof", 3), ("prev_header", 0)):
        kk = copy.deepcopy(k)
        kk[field][off] ^= 1
        g, w = ctx.next_header([kk]), orc.next_header(kk, threads=8)
        _same(g, w)
        assert w["fail"] != 0


def test_skip_fixtures(ctx, orc, golden):
    from blobstreamx_b200 import inputs as I
    ks = []
    for trusted, target in (("10000", "10500"), ("3000", "3100")):
        ks.append(I.get_skip_inputs(golden["headers"][trusted], golden["validators"][trusted], golden["headers"][target],
                                    golden["commits"][target], golden["validators"][target]))
    got = ctx.verify_skip(ks)          # two instances in one call
    for i, k in enumerate(ks):
        want = orc.verify_skip(k, threads=8)
        _same(got, want, i)
        assert want["fail"] == 0 and got["sha256_digests"].shape[1] == 490
    k = copy.deepcopy(ks[1])
    k["trusted_header"][0] ^= 1
    g, w = ctx.verify_skip([k]), orc.verify_skip(k, threads=8)
    _same(g, w)
    assert w["fail"] & 128


def test_skip_synthetic_100_validators(ctx, orc):
    """100 real signatures, nil + absent votes, and every class of tampering the circuit asserts on."""
    from blobstreamx_b200 import inputs as I, synthetic as S
    vs = S.ValidatorSet.make()
    _, skip, chain = S.header_range_inputs(2, 4, valset=vs)
    got, want = ctx.verify_skip([skip]), orc.verify_skip(skip, threads=8)
    _same(got, want)
    assert want["fail"] == 0 and (got["ed"][0][:, 520] == 0xF).all()
    commit = S.make_commit(chain, chain.start + 8, absent=(3,), nil=(90,))
    k = I.get_skip_inputs(chain.headers[0], vs.validators, chain.headers[-1], commit, vs.validators)
    _same(ctx.verify_skip([k]), orc.verify_skip(k, threads=8))
    cases = []
    for what in range(9):
        kk = copy.deepcopy(k)
        t = kk["target"]
        if what == 0: t["validators"][0, 40] ^= 1              # signature
        if what == 1: t["validators"][5, 96 + 20] ^= 1         # signed message (block hash)
        if what == 2: t["validators"][7, 224] ^= 1             # voting power -> validators hash
        if what == 3: t["validators_hash_proof"][60] ^= 1      # aunt
        if what == 4: t["chain_id_enc"][4] ^= 1
        if what == 5: t["height"] += 1
        if what == 6: kk["trusted_pubkeys"][2, 1] ^= 1
        if what == 7: kk["trusted_block"] = t["height"] - 1    # skip distance
        if what == 8:
            t["validators"][:, 236] = 0                        # nobody signed: thresholds
        cases.append(kk)
    got = ctx.verify_skip(cases)
    for i, kk in enumerate(cases):
        want = orc.verify_skip(kk, threads=8)
        _same(got, want, i)
        assert want["fail"] != 0, i


@pytest.mark.parametrize("n_max", [2, 4, 32])
def test_verify_header_small_validator_sets(ctx, orc, golden, n_max):
    """VALIDATOR_SET_SIZE_MAX other than 100 (the reference's test circuits use 2, 4, 8, 32)."""
    from blobstreamx_b200 import inputs as I
    k = I.get_step_inputs(golden["headers"]["10000"], golden["headers"]["10001"], golden["commits"]["10001"],
                          golden["validators"]["10001"], n_max=n_max)
    got, want = ctx.next_header([k], N=n_max), orc.next_header(k, threads=4)
    _same(got, want)
    assert want["fail"] == 0
    g2, w2 = ctx.verify_header([k["next"]], N=n_max), orc.verify_header(k["next"], threads=4)
    _same(g2, w2)


def test_header_range_combined(ctx, orc):
    """bsx_header_range: skip + map/reduce of several ranges in one call (the two halves run on two streams);
    every output equals the oracle's, and the skip target header is the range's end header."""
    import bench
    from blobstreamx_b200 import synthetic as S
    vs = S.ValidatorSet.make()
    J, B = 4, 8
    sets = [S.header_range_inputs(J, B, nb, start=3_000_000 + 1000 * r, seed=S.SEED + r, valset=vs) for r, nb in enumerate((None, 19, 2))]
    ms, skips = [x[0] for x in sets], [x[1] for x in sets]
    m = bench.tile_ranges(ms, len(ms))
    m["n_jobs"], m["batch"] = J, B
    got = ctx.header_range(skips, m)
    for r, (mm, k) in enumerate(zip(ms, skips)):
        ws = orc.verify_skip(k, threads=8)
        assert ws["fail"] == 0 and got["skip"]["fail"][r] == 0
        assert (got["skip"]["sha256_digests"][r] == ws["sha256_digests"]).all() and (got["skip"]["ed"][r] == ws["ed"]).all()
        wm = orc.prove_data_commitment(J, B, mm.dh_leaf, mm.dh_aunts, mm.lb_leaf, mm.lb_aunts, mm.start_headers, mm.end_headers,
                                       mm.start_block, mm.start_header, mm.end_block, mm.end_header)
        assert wm["fail"] == 0 and got["fail"][r] == 0
        assert (got["map_digests"][r] == wm["map_digests"]).all() and (got["reduce_nodes"][r] == wm["reduce_nodes"]).all()
        assert got["data_commitments"][r].tobytes() == wm["data_commitment"]
        assert k["target"]["header"].tobytes() == mm.end_header.tobytes()   # circuit wiring: skip output = range end header


def test_header_range_chunked_pipeline(ctx, orc):
    """bsx_header_range with more ranges than one pipeline chunk (chunks of 4 rotate over three copy/compute streams,
    the last chunk ragged): every range's map/reduce witness and skip digests still land in the caller's arrays."""
    import bench
    from blobstreamx_b200 import synthetic as S
    vs = S.ValidatorSet.make()
    J, B = 2, 4
    fills = (None, 5, 2, None, 7, 3, None, 3, 8, 6, None, 4, 2)          # 13 ranges -> chunks 4,4,4,1
    sets = [S.header_range_inputs(J, B, nb, start=4_000_000 + 100 * r, seed=S.SEED + 7 * r, valset=vs) for r, nb in enumerate(fills)]
    ms, skips = [x[0] for x in sets], [x[1] for x in sets]
    m = bench.tile_ranges(ms, len(ms))
    m["n_jobs"], m["batch"] = J, B
    got = ctx.header_range(skips, m)
    assert not got["fail"].any() and not got["skip"]["fail"].any()
    for r, (mm, k) in enumerate(zip(ms, skips)):
        wm = orc.prove_data_commitment(J, B, mm.dh_leaf, mm.dh_aunts, mm.lb_leaf, mm.lb_aunts, mm.start_headers, mm.end_headers,
                                       mm.start_block, mm.start_header, mm.end_block, mm.end_header)
        assert (got["map_digests"][r] == wm["map_digests"]).all() and (got["map_subchains"][r] == wm["map_subchains"]).all(), r
        assert (got["reduce_digests"][r] == wm["reduce_digests"]).all() and (got["reduce_nodes"][r] == wm["reduce_nodes"]).all(), r
        assert got["data_commitments"][r].tobytes() == wm["data_commitment"], r
    for r in (0, 5, 12):
        ws = orc.verify_skip(skips[r], threads=8)
        assert (got["skip"]["sha256_digests"][r] == ws["sha256_digests"]).all() and (got["skip"]["ed"][r] == ws["ed"]).all()


def test_sign_bit_assertions(ctx, orc):
    """marshal_int64_varint asserts bit 63 of its argument is zero (TX/builder/shared.rs:77-80: voting powers, the height)
    and verify_non_negative_round asserts the round's sign bit (TX/builder/validator.rs:73-78): a power, height or round
    >= 2^63 must be flagged (circuit witness generation would abort), the same in the kernel and in the oracle."""
    from blobstreamx_b200 import synthetic as S
    vs = S.ValidatorSet.make()
    _, skip, _ = S.header_range_inputs(2, 4, valset=vs)
    cases, bits = [], []
    for what in range(4):
        kk = copy.deepcopy(skip)
        t = kk["target"]
        if what == 0: t["validators"][11, 231] |= 0x80; bits.append(2)            # target voting power, bit 63
        if what == 1: kk["trusted_powers"][4] |= np.uint64(1 << 63); bits.append(256)
        if what == 2: t["height"] = int(t["height"]) | (1 << 63); bits.append(64)
        if what == 3: t["round"] = 1 << 63; bits.append(16)
        cases.append(kk)
    got = ctx.verify_skip(cases)
    for i, kk in enumerate(cases):
        want = orc.verify_skip(kk, threads=8)
        _same(got, want, i)
        assert want["fail"] & bits[i], (i, want["fail"])

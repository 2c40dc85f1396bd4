"""GPU parity for the device-side encoders (bsx_encode_headers, bsx_validator_records, bsx_present_on_trusted): bit-exact
against oracle/tendermint.c on the reference's fixtures and on synthetic batches with every proto3 default, vote kind,
over-long message and below-threshold trusted set; and chained into the kernels that consume them (header hashes of the
fixture chain, verify_skip accepting the fixture)."""
import numpy as np
import pytest

from tests._encode_cases import random_commits, random_header_fields

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


def _oracle_header_records(orc, fields):
    out = np.zeros((len(fields), 512), np.uint8)
    for i, r in enumerate(fields):
        lens, body = orc.encode_header_fields(r)
        out[i, :14] = lens
        out[i, 16:16 + len(body)] = np.frombuffer(body, np.uint8)
    return out


def test_encode_headers_fixtures_to_header_hash(ctx, orc, golden):
    from blobstreamx_b200 import inputs as I
    hs = sorted(golden["headers"], key=int)
    fields = I.pack_header_fields([golden["headers"][h] for h in hs])
    rec = ctx.encode_headers(fields)
    assert np.array_equal(rec, _oracle_header_records(orc, fields))
    assert np.array_equal(rec, np.stack([I.pack_header_record(I.header_leaves(golden["headers"][h])) for h in hs]))
    roots = ctx.header_trees(rec)
    for h, root in zip(hs, roots):
        if h in golden["commits"]:
            assert root.tobytes().hex().upper() == golden["commits"][h]["block_id"]["hash"]


@pytest.mark.parametrize("n", [1, 40, 5000])
def test_encode_headers_random(ctx, orc, n):
    fields = random_header_fields(min(n, 200), seed=n)
    fields = np.tile(fields, (n + len(fields) - 1) // len(fields))[:n]
    rec = ctx.encode_headers(fields)
    want = _oracle_header_records(orc, fields[:200])
    assert np.array_equal(rec[: len(want)], want)
    if n > 200:
        assert np.array_equal(rec[200:400], want[: len(rec[200:400])])


@pytest.mark.parametrize("height", ["10001", "10500", "157001", "3100"])
def test_validator_records_fixtures(ctx, orc, golden, height):
    from blobstreamx_b200 import inputs as I
    hdr, commit, vals = golden["headers"][height], golden["commits"][height], golden["validators"][height]
    cm, sg = I.pack_commit(hdr, commit, vals, 100)
    got = ctx.validator_records(cm, sg, 100, records=True, hash_fields=True)
    want = orc.validator_records(cm, sg, 100)
    assert got["fail"][0] == 0 and want["bad"] == 0
    for k in ("validators", "pubkeys", "powers", "byte_lengths"):
        assert np.array_equal(got[k][0], want[k]), k
    assert np.array_equal(got["validators"][0], I.get_validator_data_from_block(vals, hdr, commit, 100))


@pytest.mark.parametrize("n,N", [(1, 4), (12, 20), (300, 100)])
def test_validator_records_random(ctx, orc, n, N):
    cm, tg, tr, n_tg, n_tr = random_commits(n, N, seed=n + N)
    got = ctx.validator_records(cm, tg, N, records=True, hash_fields=True)
    only_hf = ctx.validator_records(cm, tg, N, records=False, hash_fields=True)
    val = got["validators"].copy()
    fail = ctx.present_on_trusted(tg, n_tg, tr, n_tr, val)
    n_short = 0
    for c in range(n):
        want = orc.validator_records(cm[c], tg[c], N)
        assert got["fail"][c] == (512 if want["bad"] else 0), c
        for k in ("validators", "pubkeys", "powers", "byte_lengths"):
            assert np.array_equal(got[k][c], want[k]), (c, k)
        bad = orc.present_on_trusted(tg[c], int(n_tg[c]), tr[c], int(n_tr[c]), want["validators"])
        assert np.array_equal(val[c], want["validators"]), c
        assert fail[c] == (1024 if bad else 0), c
        n_short += bad
    for k in ("pubkeys", "powers", "byte_lengths"):
        assert np.array_equal(only_hf[k], got[k])
    if n >= 12:
        assert 0 < n_short < n and (got["fail"] != 0).sum() >= 2


def test_skip_fixture_from_device_shaped_validators(ctx, orc, golden):
    """verify_skip accepts the 10000 -> 10500 and 3000 -> 3100 fixtures when the validator records, the trusted hash fields
    and present_on_trusted_header all come from the device-side shapers."""
    from blobstreamx_b200 import inputs as I
    for trusted, target in (("10000", "10500"), ("3000", "3100")):
        hdr, commit, vals = golden["headers"][target], golden["commits"][target], golden["validators"][target]
        tvals = golden["validators"][trusted]
        k = I.get_skip_inputs(golden["headers"][trusted], tvals, hdr, commit, vals)
        cm, sg = I.pack_commit(hdr, commit, vals, 100)
        val = ctx.validator_records(cm, sg, 100)["validators"]
        tr = np.zeros(100, I.COMMIT_SIG_DTYPE)
        for i, v in enumerate(tvals):
            tr[i]["address"] = np.frombuffer(bytes.fromhex(v["address"]), np.uint8)
        tcm = np.zeros((), I.COMMIT_DTYPE)
        tcm["n_signatures"] = len(tvals)
        for i, v in enumerate(tvals):
            tr[i]["pubkey"] = np.frombuffer(I._b64(v["pub_key"]), np.uint8)
            tr[i]["voting_power"] = int(v["voting_power"])
        hf = ctx.validator_records(tcm, tr, 100, records=False, hash_fields=True)
        assert ctx.present_on_trusted(sg, [len(vals)], tr, [len(tvals)], val)[0] == 0
        assert np.array_equal(val[0], k["target"]["validators"])
        assert np.array_equal(hf["pubkeys"][0], k["trusted_pubkeys"]) and np.array_equal(hf["powers"][0], k["trusted_powers"])
        assert np.array_equal(hf["byte_lengths"][0], k["trusted_byte_lengths"])
        k2 = dict(k, target=dict(k["target"], validators=val[0]), trusted_pubkeys=hf["pubkeys"][0], trusted_powers=hf["powers"][0],
                  trusted_byte_lengths=hf["byte_lengths"][0])
        r = ctx.verify_skip([k2])
        assert r["fail"][0] == 0 and r["sha256_digests"].shape[1] == 490
        assert np.array_equal(r["sha256_digests"][0], orc.verify_skip(k)["sha256_digests"])

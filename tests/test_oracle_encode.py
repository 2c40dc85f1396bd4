"""Off-chain input shaping restated in oracle/tendermint.c (protobuf header fields, CanonicalVote sign-bytes, validator
records, trusted-set walk) against the host shaper blobstreamx_b200/inputs.py, which the golden tests pin to the
reference's fixtures: header hashes (test_oracle_golden.test_header_hash_chain) and signatures that verify over the
sign-bytes with libsodium (test_commit_signatures)."""
import base64

import numpy as np
import pytest

from blobstreamx_b200 import inputs as I
from oracle import cbind as orc
from oracle import pyoracle as po

from tests._encode_cases import random_commits, random_header_fields

HEIGHTS = ["10000", "10001", "10500", "10501", "157001", "3000", "3001", "3100"]


def test_header_fields_fixtures(golden):
    for h, hdr in golden["headers"].items():
        rec = I.pack_header_fields([hdr])[0]
        lens, body = orc.encode_header_fields(rec)
        leaves = I.header_leaves(hdr)
        assert [len(x) for x in leaves] == lens.tolist() and body == b"".join(leaves), h
        # and through the tree: the header hash of the fixtures
        if h in golden["commits"]:
            off = [int(x) for x in np.concatenate([[0], np.cumsum(lens.astype(np.int64))])]
            items = [body[off[k]:off[k + 1]] for k in range(14)]
            assert orc.tm_root_from_slices(items).hex().upper() == golden["commits"][h]["block_id"]["hash"]


def test_header_fields_proto3_defaults():
    f = random_header_fields(40)
    for r in f:
        lens, body = orc.encode_header_fields(r)
        # python restatement of the same rules
        leaves = [I._vi(0x08, int(r["version_block"])) + I._vi(0x10, int(r["version_app"])),
                  I._ld(0x0A, r["chain_id"][: r["chain_id_len"]].tobytes()), I._vi(0x08, int(r["height"])),
                  I._timestamp((int(r["time_seconds"]), int(r["time_nanos"])))]
        if r["has_last_block_id"]:
            parts = I._vi(0x08, int(r["parts_total"])) + I._ld(0x12, r["parts_hash"].tobytes())
            leaves.append(I._ld(0x0A, r["last_block_hash"].tobytes()) + I._ld(0x12, parts))
        else:
            leaves.append(b"")
        leaves += [I._ld(0x0A, r["hashes"][k][: r["hash_len"][k]].tobytes()) for k in range(9)]
        assert lens.tolist() == [len(x) for x in leaves] and body == b"".join(leaves)


@pytest.mark.parametrize("height", HEIGHTS)
def test_validator_records_fixtures(golden, height):
    hdr, commit, vals = golden["headers"][height], golden["commits"][height], golden["validators"][height]
    cm, sg = I.pack_commit(hdr, commit, vals, 100)
    got = orc.validator_records(cm, sg, 100)
    want = I.get_validator_data_from_block(vals, hdr, commit, 100)
    assert got["bad"] == 0 and np.array_equal(got["validators"], want)
    pks, powers, blens = I.validator_hash_fields(vals, 100)
    assert np.array_equal(got["pubkeys"], pks) and np.array_equal(got["powers"], powers) and np.array_equal(got["byte_lengths"], blens)
    # sign-bytes on their own, against the python oracle (signatures verify over these: test_commit_signatures)
    for cs in commit["signatures"]:
        if int(cs["block_id_flag"]) != 2:
            continue
        secs, nanos = I.parse_time(cs["timestamp"])
        bid = commit["block_id"]
        msg = orc.vote_sign_bytes(hdr["chain_id"].encode(), int(commit["height"]), int(commit["round"]), bytes.fromhex(bid["hash"]),
                                  int(bid["parts"]["total"]), bytes.fromhex(bid["parts"]["hash"]), secs, nanos)
        assert msg == po.canonical_vote_sign_bytes(hdr["chain_id"], int(commit["height"]), int(commit["round"]), bid, cs["timestamp"])


def test_validator_records_edges():
    cm, tg, _, _, _ = random_commits(12, 20)
    for c in range(12):
        got = orc.validator_records(cm[c], tg[c], 20)
        assert got["bad"] == (1 if c in (2, 4) else 0), c
        k = min(int(cm[c]["n_signatures"]), 20)
        v = got["validators"]
        for i in range(20):
            signed = i < k and tg[c][i]["block_id_flag"] == 2
            assert v[i, 236] == signed and v[i, 237] == 0
            if not signed:
                assert v[i, 32:96].tobytes() == I.DUMMY_SIGNATURE and not v[i, 96:220].any() and v[i, 220] == 32
            if i >= k:
                assert v[i, :32].tobytes() == I.DUMMY_PUBLIC_KEY and got["byte_lengths"][i] == 46 and got["powers"][i] == 0
            elif c != 4:
                want = I.vote_sign_bytes(cm[c]["chain_id"][: cm[c]["chain_id_len"]].tobytes().decode(), int(cm[c]["height"]),
                                         int(cm[c]["round"]),
                                         dict(hash=cm[c]["block_hash"].tobytes().hex(), parts=dict(total=int(cm[c]["parts_total"]),
                                              hash=cm[c]["parts_hash"].tobytes().hex())) if cm[c]["has_block_id"] else None,
                                         (int(tg[c][i]["ts_seconds"]), int(tg[c][i]["ts_nanos"])))
                if signed:
                    ln = int(v[i, 220:224].view(np.uint32)[0])
                    assert v[i, 96:96 + ln].tobytes() == want and not v[i, 96 + ln:220].any()
                assert got["byte_lengths"][i] == len(I.validator_bytes(bytes(32), int(tg[c][i]["voting_power"])))


@pytest.mark.parametrize("pair", [("10000", "10500"), ("3000", "3100"), ("10000", "10001"), ("3000", "3001")])
def test_present_on_trusted_fixtures(golden, pair):
    trusted, target = pair
    hdr, commit, vals = golden["headers"][target], golden["commits"][target], golden["validators"][target]
    tvals = golden["validators"][trusted]
    want = I.get_validator_data_from_block(vals, hdr, commit, 100)
    I.update_present_on_trusted_header(want, commit, vals, tvals)
    _, tg = I.pack_commit(hdr, commit, vals, 100)
    tr = np.zeros(100, I.COMMIT_SIG_DTYPE)
    for i, v in enumerate(tvals):
        tr[i]["address"] = np.frombuffer(bytes.fromhex(v["address"]), np.uint8)
    got = orc.validator_records(*I.pack_commit(hdr, commit, vals, 100), 100)["validators"]
    assert orc.present_on_trusted(tg, len(vals), tr, len(tvals), got) == 0
    assert np.array_equal(got, want) and got[:, 237].sum() >= 1


def test_present_on_trusted_random():
    cm, tg, tr, n_tg, n_tr = random_commits(24, 40, seed=5)
    n_bad = 0
    for c in range(24):
        val = np.zeros((40, 240), np.uint8)
        bad = orc.present_on_trusted(tg[c], int(n_tg[c]), tr[c], int(n_tr[c]), val)
        # python walk (TX/input/conversion.rs:186-240)
        k = int(n_tg[c])
        total, shared, want = int(tg[c]["voting_power"][:k].astype(object).sum()), 0, np.zeros(40, np.uint8)
        addr = [tg[c][i]["address"].tobytes() for i in range(k)]
        for s in range(int(n_tr[c])):
            if not (float(total) * (1.0 / 3.0) > float(shared)):
                break
            a = tr[c][s]["address"].tobytes()
            if a in addr:
                i = addr.index(a)
                for j in range(k):
                    if tg[c][j]["block_id_flag"] in (2, 3) and tg[c][j]["sig_address"].tobytes() == a:
                        shared += int(tg[c][i]["voting_power"])
                        want[i] = 1
        assert np.array_equal(val[:, 237], want) and bad == int(float(total) * (1.0 / 3.0) > float(shared)), c
        n_bad += bad
    assert 0 < n_bad < 24

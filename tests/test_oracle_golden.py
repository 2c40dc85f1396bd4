"""Pin the CPU oracle (oracle/*.c and oracle/pyoracle.py) against the reference's own fixtures and
KATs.  CPU only.  If these fail, no GPU parity claim means anything."""
import base64
import hashlib

import numpy as np
import pytest

from blobstreamx_b200 import inputs as bx_inputs
from oracle import cbind as orc
from oracle import pyoracle as po

H = bytes.fromhex


def test_sha256_kats():
    # PX/frontend/hash/sha/sha256/curta.rs:223-224
    assert orc.sha256(b"\x00").hex() == "6e340b9cffb37a989ca544e6bb780a2c78901d3fb33738768511a30617afa01d"
    # variable-length KAT :331-341 (first 39 bytes of a 64-byte buffer)
    buf = H("00de6ad0941095ada2a7996e6a888581928203b8b69e07ee254d289f5b9c9caea193c2ab01902d" + "00" * 25)
    assert len(buf) == 64
    assert orc.sha256(buf[:39]).hex() == "84f633a570a987326947aafd434ae37f151e98d5e6d429137a4cc378d4a7988e"
    padded, last_chunk = orc.sha256_pad_variable(buf, 39)
    assert len(padded) == 128 and last_chunk == 0
    assert padded[:39] == buf[:39] and padded[39] == 0x80 and padded[40:56] == bytes(16)
    assert padded[56:64] == (39 * 8).to_bytes(8, "big") and padded[64:] == bytes(64)
    # full-buffer length: 0x80 lands in chunk 1, length at the end of chunk 1
    padded, last_chunk = orc.sha256_pad_variable(buf, 64)
    assert last_chunk == 1 and padded[64] == 0x80 and padded[120:128] == (512).to_bytes(8, "big")


def test_sha512_kats():
    # PX/frontend/hash/sha/sha512/curta.rs:236,245,254
    assert orc.sha512(b"").hex() == ("cf83e1357eefb8bdf1542850d66d8007d620e4050b5715dc83f4a921d36ce9ce"
                                     "47d0d13c5d85f2b0ff8318d2877eec2f63b931bd47417a81a538327af927da3e")
    assert orc.sha512(b"plonky2").hex() == ("7c6159dd615db8c15bc76e23d36106e77464759979a0fcd1366e531f552cfa08"
                                            "52dbf5c832f00bb279cbc945b44a132bff3ed0028259813b6a07b57326e88c87")
    msg = H("35c323757c20640a294345c89c0bfcebe3d554fdb0c7b7a0bdb72222c531b1ecf7ec1c43f4de9d49556de87b86b26a98"
            "942cb078486fdb44de38b80864c3973153756363696e6374204c616273")
    assert orc.sha512(msg).hex() == ("4388243c4452274402673de881b2f942ff5730fd2c7d8ddb94c3e3d789fb3754"
                                     "380cba8faa40554d9506a0730a681e88ab348a04bc5c41d18926f140b59aed39")


def test_sha_random_vs_hashlib():
    rng = np.random.default_rng(1)
    for n in list(range(0, 260)) + [511, 512, 513, 1000]:
        m = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert orc.sha256(m) == hashlib.sha256(m).digest(), n
        assert orc.sha512(m) == hashlib.sha512(m).digest(), n


def test_header_hash_chain(golden):
    hs = golden["headers"]
    # block 10000's hash is recorded in its own commit; 10001..10003 in the next header's last_block_id
    assert po.header_hash(hs["10000"]).hex().upper() == golden["commits"]["10000"]["block_id"]["hash"]
    assert po.header_hash(hs["10000"]).hex().upper() == "A0123D5E4B8B8888A61F931EE2252D83568B97C223E0ECA9795B29B8BD8CBA2D"
    for h in range(10000, 10004):
        want = hs[str(h + 1)]["last_block_id"]["hash"]
        got_py = po.header_hash(hs[str(h)])
        assert got_py.hex().upper() == want
        assert orc.tm_root_from_slices(po.header_fields(hs[str(h)])) == got_py
        assert bx_inputs.header_hash(hs[str(h)]) == got_py
    for h, c in golden["commits"].items():
        assert po.header_hash(hs[h]).hex().upper() == c["block_id"]["hash"], h


def test_header_leaf_sizes(golden):
    # [4, 9, 3, 11-12, 72, 34 x8, 22]
    for h in ("10000", "10004", "157001"):
        sizes = [len(x) for x in po.header_fields(golden["headers"][h])]
        assert sizes[:2] == [4, 9] and sizes[2] in (3, 4) and sizes[4:] == [72] + [34] * 8 + [22] and sizes[3] in (11, 12)
        assert po.header_fields(golden["headers"][h]) == bx_inputs.header_leaves(golden["headers"][h])


def test_data_commitment_fixtures(golden):
    hs = golden["headers"]
    for rng_, want in golden["data_commitments"].items():
        a, b = (int(x) for x in rng_.split("-"))
        dhs = [H(hs[str(i)]["data_hash"]) for i in range(a, b)]
        assert po.data_commitment(dhs, a).hex().upper() == want
        # fixed-shape in-circuit evaluation gives the same root for every MAX_LEAVES >= n
        for B in (4, 8, 32):
            if B < len(dhs):
                continue
            padded = np.zeros((B, 32), np.uint8)
            for i, d in enumerate(dhs):
                padded[i] = np.frombuffer(d, np.uint8)
            dig, root, fail = orc.get_data_commitment(padded, a, b)
            assert root.hex().upper() == want and fail == 0
            assert dig[0].tobytes() == po.leaf_hash(po.encode_data_root_tuple(dhs[0], a))


def test_tuple_and_tree_kats(golden):
    k = golden["kats"]
    assert po.encode_data_root_tuple(b"\xff" * 32, 256).hex() == k["tuple_height256_ff"]
    leaves = [bytes(48)] * 32
    ld = np.stack([np.frombuffer(po.leaf_hash(x), np.uint8) for x in leaves])
    inner, root = orc.tm_merkle_tree(ld, 32)
    assert root.hex() == k["tm_tree_32x48zero_root"]
    inner_py, root_py = po.merkle_tree_schedule([x.tobytes() for x in ld], 32)
    assert root_py == root and [x.tobytes() for x in inner] == inner_py
    assert po.tm_root(leaves) == root
    aunts = b"".join(H(a) for a in k["tm_proof_depth4_aunts"])
    dig, root = orc.tm_merkle_proof(bytes(48), aunts, 4, 0)
    assert root.hex() == k["tm_proof_depth4_root"]
    dpy, rpy = po.merkle_proof_schedule(bytes(48), [aunts[32 * i:32 * i + 32] for i in range(4)], [False] * 4)
    assert rpy == root and [d.tobytes() for d in dig] == dpy


def test_tree_disabled_leaves_match_variable_shape():
    rng = np.random.default_rng(5)
    for N in (1, 2, 3, 5, 8, 13, 32, 64, 100):
        leaves = [rng.integers(0, 256, 40, dtype=np.uint8).tobytes() for _ in range(N)]
        ld = np.stack([np.frombuffer(po.leaf_hash(x), np.uint8) for x in leaves])
        for nb in sorted({1, 2, N // 2, N - 1, N} - {0}):
            _, root = orc.tm_merkle_tree(ld, nb)
            assert root == po.tm_root(leaves[:nb]), (N, nb)
            assert po.merkle_tree_schedule([x.tobytes() for x in ld], nb)[1] == root
        # nb_enabled beyond the padded size leaves everything enabled (tendermint.rs:184-194)
        P = 1 << (N - 1).bit_length() if N > 1 else 1
        _, r_all = orc.tm_merkle_tree(ld, P + 7)
        assert r_all == po.merkle_tree_schedule([x.tobytes() for x in ld], P + 7)[1]


def test_varint_and_validator_marshal(golden):
    for v, want in golden["kats"]["varint"]:
        b, n = orc.marshal_int64_varint(v)
        assert list(b[:n]) == want and set(b[n:]) <= {0}
        assert po.marshal_int64_varint9(v) == (b, n)
    m = golden["kats"]["validator_marshal"]
    b, n = orc.marshal_validator(H(m["pubkey"]), m["power"])
    assert b[:n].hex() == m["bytes"] and b[:n] == po.marshal_validator(H(m["pubkey"]), m["power"])


@pytest.mark.parametrize("height", ["10000", "157001", "3000", "10500"])
def test_validators_hash_fixtures(golden, height):
    vals = golden["validators"][height]
    want = golden["headers"][height]["validators_hash"]
    pairs = [(base64.b64decode(v["pub_key"]), int(v["voting_power"])) for v in vals]
    assert po.validators_hash(pairs).hex().upper() == want
    pks, powers, blens = bx_inputs.validator_hash_fields(vals, 100)
    dig, root = orc.hash_validator_set(pks, powers, blens, len(vals))
    assert root.hex().upper() == want
    assert dig.shape[0] == 100 + 127
    assert dig[0].tobytes() == po.leaf_hash(po.marshal_validator(*pairs[0]))


def test_hash_in_message_kat():
    # TX/builder/verify.rs:598-602
    hh = H("8909e1b73b7d987e95a7541d96ed484c17a4b0411e98ee4b7c890ad21302ff8c")
    msg = H("6b080211de3202000000000022480a208909e1b73b7d987e95a7541d96ed484c17a4b0411e98ee4b7c890ad21302ff8c"
            "12240801122061263df4855e55fcab7aab0a53ee32cf4f29a1101b56de4a9d249d44e4cf96282a0b089dce84a60610ebb7a8"
            "1932076d6f6368612d33")
    assert msg[16:48] == hh and msg[1:3] == b"\x08\x02"
    assert int.from_bytes(msg[4:12], "little") == 0x232DE


def test_dummy_signature_is_rfc8032():
    from nacl.signing import SigningKey

    sk = SigningKey(bytes([1] * 32))
    assert bytes(sk.verify_key) == po.DUMMY_PUBLIC_KEY == bx_inputs.DUMMY_PUBLIC_KEY
    assert sk.sign(bytes(32)).signature == po.DUMMY_SIGNATURE == bx_inputs.DUMMY_SIGNATURE
    w = po.ed_witness(po.DUMMY_PUBLIC_KEY, po.DUMMY_SIGNATURE, bytes(32))
    assert w["verified"] and w["s_lt_l"] and w["a_ok"] and w["r_ok"]
    assert orc.ed25519_witness(po.DUMMY_PUBLIC_KEY, po.DUMMY_SIGNATURE, bytes(32)) == \
        po.ed_witness_bytes(po.DUMMY_PUBLIC_KEY, po.DUMMY_SIGNATURE, bytes(32))


@pytest.mark.parametrize("height", ["10000", "10001", "157001"])
def test_commit_signatures(golden, height):
    """sign-bytes restatement + Ed25519 equation vs libsodium on the reference's commits
    (157001: 98 commit sigs, 1 nil, 1 absent)."""
    from nacl.exceptions import BadSignatureError
    from nacl.signing import VerifyKey

    hdr, commit, vals = golden["headers"][height], golden["commits"][height], golden["validators"][height]
    recs = bx_inputs.get_validator_data_from_block(vals, hdr, commit, 100)
    n_signed = 0
    for i, cs in enumerate(commit["signatures"]):
        r = recs[i]
        if int(cs["block_id_flag"]) != 2:
            assert r[236] == 0 and r[32:96].tobytes() == po.DUMMY_SIGNATURE
            continue
        n_signed += 1
        pk, sig = r[0:32].tobytes(), r[32:96].tobytes()
        mlen = int.from_bytes(r[220:224].tobytes(), "little")
        msg = r[96:96 + mlen].tobytes()
        assert msg == po.canonical_vote_sign_bytes(hdr["chain_id"], int(commit["height"]), int(commit["round"]),
                                                   commit["block_id"], cs["timestamp"])
        assert 100 <= mlen <= 124 and msg[16:48].hex().upper() == commit["block_id"]["hash"]
        VerifyKey(pk).verify(msg, sig)  # raises on failure
        if i < 6:  # python-int witness is slow: spot check, then C oracle == python oracle
            assert orc.ed25519_witness(pk, sig, msg) == po.ed_witness_bytes(pk, sig, msg)
    assert n_signed == (98 if height == "157001" else 2)
    lens = recs[:, 220:224].copy().view(np.uint32).reshape(-1)
    out = orc.ed25519_batch(recs[:, 0:32], recs[:, 32:96], recs[:, 96:220], lens, recs[:, 236], threads=orc.max_threads())
    assert (out[:, 520] == 0xF).all()  # every lane (real or dummy) satisfies sG == R + hA
    # negative: flipped message bit must fail (eddsa.rs:344-386 must-panic test)
    bad = recs[0, 96:220].copy()
    bad[20] ^= 1
    w = orc.ed25519_witness(recs[0, 0:32].tobytes(), recs[0, 32:96].tobytes(), bad[: int(lens[0])].tobytes())
    assert w[520] & 8 == 0
    with pytest.raises(BadSignatureError):
        VerifyKey(recs[0, 0:32].tobytes()).verify(bad[: int(lens[0])].tobytes(), recs[0, 32:96].tobytes())


def test_ed25519_primitives_vs_python():
    rng = np.random.default_rng(7)
    from nacl.signing import SigningKey

    for t in range(4):
        sk = SigningKey(rng.integers(0, 256, 32, dtype=np.uint8).tobytes())
        pk = bytes(sk.verify_key)
        xy, root, ok = orc.ed25519_decompress(pk)
        (x, y), r, ok_py = po.ed_decompress(pk)
        assert ok and ok_py and xy == po.ed_point_bytes((x, y)) and root == r.to_bytes(32, "little")
        assert r % 2 == 0 and (r * r - (y * y - 1) * pow(po.D25519 * y * y + 1, po.P25519 - 2, po.P25519)) % po.P25519 == 0
        assert x == (r if pk[31] >> 7 == 0 else po.P25519 - r)
        k = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()  # full 256-bit scalar, not reduced
        want = po.ed_point_bytes(po.ed_mul(int.from_bytes(k, "little"), (x, y)))
        assert orc.ed25519_scalar_mul(k, xy) == want
        if t == 0:
            assert orc.ed25519_scalar_mul(k, xy, affine=True) == want  # literal affine double-and-add agrees
        q = po.ed_mul(12345 + t, po.G)
        assert orc.ed25519_add(xy, po.ed_point_bytes(q)) == po.ed_point_bytes(po.ed_add((x, y), q))
    # a y with no valid x: decompress reports failure (reference panics)
    bad = next(b for b in (int(i).to_bytes(32, "little") for i in range(2, 50)) if not po.ed_decompress(b)[2])
    assert orc.ed25519_decompress(bad)[2] is False


def _trees(golden, heights):
    return {h: bx_inputs.HeaderTree.build(bx_inputs.header_leaves(golden["headers"][str(h)])) for h in heights}


@pytest.mark.parametrize("B", [4, 8])
def test_prove_subchain_fixture(golden, B):
    """BX/circuits/builder.rs:524-564 test_prove_header_chain: 10000 -> 10004 in one batch."""
    trees = _trees(golden, range(10000, 10005))
    d = bx_inputs.get_data_commitment_inputs(trees, 10000, 10004, B)
    dig, sub = orc.prove_subchain(B, d.dh_leaf, d.dh_aunts, d.lb_leaf, d.lb_aunts, d.start_header, d.end_header,
                                  10000, 10000 + B, 10004, trees[10004].root)
    assert sub[0] == 1 and int.from_bytes(sub[4:8].tobytes(), "little") == 0
    assert sub[88:120].tobytes().hex().upper() == golden["data_commitments"]["10000-10004"]
    assert sub[56:88].tobytes() == trees[10004].root and int.from_bytes(sub[16:24].tobytes(), "little") == 10004
    # schedule: proof i = 9 digests data_hash (leaf, then (left,right) x4), then 9 for last_block_id
    dpy, rpy = po.merkle_proof_schedule(d.dh_leaf[0].tobytes(), [a.tobytes() for a in d.dh_aunts[0]], po.path_bits(6))
    assert [x.tobytes() for x in dig[0:9]] == dpy and rpy == trees[10000].root
    dpy, rpy = po.merkle_proof_schedule(d.lb_leaf[0].tobytes(), [a.tobytes() for a in d.lb_aunts[0]], po.path_bits(4))
    assert [x.tobytes() for x in dig[9:18]] == dpy and rpy == trees[10001].root
    # a broken link is reported
    bad = d.lb_leaf.copy()
    bad[1, 5] ^= 1
    _, sub2 = orc.prove_subchain(B, d.dh_leaf, d.dh_aunts, bad, d.lb_aunts, d.start_header, d.end_header,
                                 10000, 10000 + B, 10004, trees[10004].root)
    assert int.from_bytes(sub2[4:8].tobytes(), "little") != 0


@pytest.mark.parametrize("J,B", [(2, 4), (4, 2), (8, 4)])
def test_prove_data_commitment_fixture(golden, J, B):
    """map + reduce over 10000 -> 10004 (shape of BX/circuits/header_range.rs:193-214 small test)."""
    trees = _trees(golden, range(10000, 10005))
    m = bx_inputs.get_header_range_map_inputs(trees, 10000, 10004, J, B)
    r = orc.prove_data_commitment(J, B, m.dh_leaf, m.dh_aunts, m.lb_leaf, m.lb_aunts, m.start_headers, m.end_headers,
                                  10000, m.start_header, 10004, m.end_header, threads=2)
    assert r["fail"] == 0
    assert r["data_commitment"].hex().upper() == golden["data_commitments"]["10000-10004"]
    # sub-range 10002 -> 10004
    m = bx_inputs.get_header_range_map_inputs(trees, 10002, 10004, J, B)
    r = orc.prove_data_commitment(J, B, m.dh_leaf, m.dh_aunts, m.lb_leaf, m.lb_aunts, m.start_headers, m.end_headers,
                                  10002, m.start_header, 10004, m.end_header)
    assert r["fail"] == 0 and r["data_commitment"].hex().upper() == golden["data_commitments"]["10002-10004"]


def test_next_header_fixture(golden):
    """next_header config 1: 10000 -> 10001 (TX/step.rs tests + BX/circuits/next_header.rs:130-179)."""
    k = bx_inputs.get_step_inputs(golden["headers"]["10000"], golden["headers"]["10001"], golden["commits"]["10001"],
                                  golden["validators"]["10001"])
    r = orc.next_header(k)
    assert r["fail"] == 0
    assert r["data_commitment"].hex().upper() == golden["data_commitments"]["10000-10001"]
    assert r["sha256_digests"].shape[0] == 282
    # the validators hash computed in-circuit is digest #226 (root of the 127-node tree) only via select;
    # check the raw leaf digests instead
    v0 = golden["validators"]["10001"][0]
    assert r["sha256_digests"][0].tobytes() == po.leaf_hash(po.marshal_validator(base64.b64decode(v0["pub_key"]), int(v0["voting_power"])))


def test_skip_fixture(golden):
    """verify_skip 10000 -> 10500 and 157000-ish 100-validator case is covered by 3000 -> 3100."""
    for trusted, target in (("10000", "10500"), ("3000", "3100")):
        k = bx_inputs.get_skip_inputs(golden["headers"][trusted], golden["validators"][trusted], golden["headers"][target],
                                      golden["commits"][target], golden["validators"][target])
        r = orc.verify_skip(k)
        assert r["fail"] == 0, (trusted, target, r["fail"])
        assert r["sha256_digests"].shape[0] == 490
    # tamper: wrong trusted header hash
    k["trusted_header"] = k["trusted_header"].copy()
    k["trusted_header"][0] ^= 1
    assert orc.verify_skip(k)["fail"] & 128

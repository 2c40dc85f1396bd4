"""GPU parity: the CUDA path (through the C ABI) vs the CPU oracle, bit-exact.  Run on the B200 box."""
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


@pytest.fixture(scope="module")
def orc():
    from oracle import cbind
    return cbind


def _ragged(rng, lens):
    offs = np.zeros(len(lens) + 1, np.uint32)
    offs[1:] = np.cumsum(lens)
    return rng.integers(0, 256, int(offs[-1]), dtype=np.uint8), offs


def test_sha256_batch_kats_and_ragged(ctx, orc):
    rng = np.random.default_rng(0)
    lens = list(range(0, 200)) + [255, 256, 257, 511, 512, 513, 1000, 4096, 0, 0, 1]
    msgs, offs = _ragged(rng, lens)
    got = ctx.sha256_batch(msgs, offs)
    want = orc.sha256_batch(msgs, offs)
    assert (got == want).all()
    for i in (0, 1, 55, 56, 64, 119, 120):
        assert got[i].tobytes() == hashlib.sha256(msgs[offs[i]:offs[i + 1]].tobytes()).digest()
    # reference KAT: sha256(0x00)  (PX/frontend/hash/sha/sha256/curta.rs:223-224)
    k = ctx.sha256_batch(np.zeros(1, np.uint8), np.array([0, 1], np.uint32))
    assert k[0].tobytes().hex() == "6e340b9cffb37a989ca544e6bb780a2c78901d3fb33738768511a30617afa01d"
    # empty batch
    assert ctx.sha256_batch(np.zeros(0, np.uint8), np.array([0], np.uint32)).shape == (0, 32)


def test_sha512_batch(ctx, orc):
    rng = np.random.default_rng(1)
    lens = list(range(0, 300, 3)) + [111, 112, 113, 127, 128, 129, 239, 240, 241, 1000]
    msgs, offs = _ragged(rng, lens)
    got = ctx.sha512_batch(msgs, offs)
    assert (got == orc.sha512_batch(msgs, offs)).all()
    k = ctx.sha512_batch(np.frombuffer(b"plonky2", np.uint8), np.array([0, 7], np.uint32))
    assert k[0].tobytes().hex().startswith("7c6159dd615db8c15bc76e23d36106e7")


@pytest.mark.parametrize("leaf_len,depth,hashed", [(34, 4, False), (72, 4, False), (48, 4, False), (32, 4, True),
                                                    (1, 1, False), (200, 9, False), (40, 0, False)])
def test_merkle_proofs(ctx, orc, leaf_len, depth, hashed):
    rng = np.random.default_rng(leaf_len * 100 + depth)
    n = 777
    ll = 32 if hashed else leaf_len
    leaves = rng.integers(0, 256, (n, ll), dtype=np.uint8)
    aunts = rng.integers(0, 256, (n, depth, 32), dtype=np.uint8)
    bits = rng.integers(0, 1 << max(depth, 1), n, dtype=np.uint32)
    dig, roots = ctx.tm_merkle_proofs(leaves, leaf_len, aunts, depth, bits, hashed)
    for i in list(range(0, n, 97)) + [n - 1]:
        d, r = orc.tm_merkle_proof(leaves[i].tobytes(), aunts[i].tobytes(), depth, int(bits[i]), hashed)
        assert (dig[i] == d).all() and roots[i].tobytes() == r


def test_merkle_proof_kat(ctx, golden):
    k = golden["kats"]
    aunts = np.frombuffer(b"".join(bytes.fromhex(a) for a in k["tm_proof_depth4_aunts"]), np.uint8)
    _, roots = ctx.tm_merkle_proofs(np.zeros(48, np.uint8), 48, aunts, 4, np.zeros(1, np.uint32))
    assert roots[0].tobytes().hex() == k["tm_proof_depth4_root"]


@pytest.mark.parametrize("N", [1, 2, 3, 5, 32, 33, 64, 100, 128, 129, 1000, 2048, 4096])
def test_merkle_tree(ctx, orc, N):
    rng = np.random.default_rng(N)
    t = 5
    leaves = rng.integers(0, 256, (t, N, 32), dtype=np.uint8)
    P = 1
    while P < N:
        P *= 2
    nb = np.array([N, 1, max(1, N // 2), P + 3, 0], np.uint64)
    inner, roots = ctx.tm_merkle_tree(leaves, N, nb)
    for j in range(t):
        wi, wr = orc.tm_merkle_tree(leaves[j], int(nb[j]))
        assert (inner[j] == wi).all(), (N, j)
        assert roots[j].tobytes() == wr


def test_merkle_tree_kat(ctx, golden):
    ld = np.tile(np.frombuffer(hashlib.sha256(b"\x00" + bytes(48)).digest(), np.uint8), (32, 1))
    _, roots = ctx.tm_merkle_tree(ld, 32, np.array([32], np.uint64))
    assert roots[0].tobytes().hex() == golden["kats"]["tm_tree_32x48zero_root"]


@pytest.mark.parametrize("N", [1, 4, 32, 64, 100, 2048])
def test_data_commitment_batch(ctx, orc, N):
    rng = np.random.default_rng(N + 7)
    t = 4
    dh = rng.integers(0, 256, (t, N, 32), dtype=np.uint8)
    start = np.array([1_000_000, 5, 2**40, 77], np.uint64)
    end = start + np.array([N, max(1, N // 3), 0, N], np.uint64)
    dig, roots, fail = ctx.data_commitment_batch(dh, N, start, end)
    for j in range(t):
        wd, wr, wf = orc.get_data_commitment(dh[j], int(start[j]), int(end[j]))
        assert (dig[j] == wd).all() and roots[j].tobytes() == wr and fail[j] == wf


def test_data_commitment_fixture(ctx, golden):
    hs = golden["headers"]
    for rng_, want in golden["data_commitments"].items():
        a, b = (int(x) for x in rng_.split("-"))
        dh = np.zeros((1, 32, 32), np.uint8)
        for i in range(a, b):
            dh[0, i - a] = np.frombuffer(bytes.fromhex(hs[str(i)]["data_hash"]), np.uint8)
        _, roots, fail = ctx.data_commitment_batch(dh, 32, np.array([a], np.uint64), np.array([b], np.uint64))
        assert roots[0].tobytes().hex().upper() == want and fail[0] == 0


def _map_args(m):
    return (m.dh_leaf, m.dh_aunts, m.lb_leaf, m.lb_aunts, m.start_headers, m.end_headers)


@pytest.mark.parametrize("J,B,nblk", [(2, 4, 4), (4, 2, 4), (8, 4, 4), (1, 1, 1), (4, 8, 19), (32, 32, None), (32, 32, 700),
                                      (32, 64, None), (2, 128, 200), (2, 256, None), (4, 16, 64)])
def test_prove_data_commitment(ctx, orc, golden, J, B, nblk):
    """map + reduce vs oracle: fixture 10000->10004 for the small shapes, synthetic chains otherwise
    (header_range_1024 = 32x32, header_range_2048 = 32x64, partially filled 700)."""
    from blobstreamx_b200 import inputs as I, synthetic as S
    if nblk == 4:
        trees = {h: I.HeaderTree.build(I.header_leaves(golden["headers"][str(h)])) for h in range(10000, 10005)}
        m = I.get_header_range_map_inputs(trees, 10000, 10004, J, B)
    else:
        m, _, _ = S.header_range_inputs(J, B, nblk, with_skip=False)
    got = ctx.prove_data_commitment(1, J, B, *_map_args(m), np.array([m.start_block], np.uint64), m.start_header,
                                    np.array([m.end_block], np.uint64), m.end_header)
    want = orc.prove_data_commitment(J, B, *_map_args(m), m.start_block, m.start_header, m.end_block, m.end_header, threads=4)
    assert want["fail"] == 0 and got["fail"][0] == 0
    assert (got["map_digests"][0] == want["map_digests"]).all()
    assert (got["map_subchains"][0] == want["map_subchains"]).all()
    assert (got["reduce_digests"][0] == want["reduce_digests"]).all()
    assert (got["reduce_nodes"][0] == want["reduce_nodes"]).all()
    assert got["data_commitments"][0].tobytes() == want["data_commitment"]
    if nblk == 4:
        assert want["data_commitment"].hex().upper() == golden["data_commitments"]["10000-10004"]


def test_prove_data_commitment_failures(ctx, orc):
    """Broken witnesses must produce the same assertion masks as the oracle."""
    from blobstreamx_b200 import synthetic as S
    J, B = 4, 8
    m, _, _ = S.header_range_inputs(J, B, 27, with_skip=False)
    rng = np.random.default_rng(3)
    for trial in range(6):
        mm = [x.copy() for x in _map_args(m)]
        eh = m.end_header.copy()
        sb, eb = m.start_block, m.end_block
        if trial == 0: mm[2][5, 7] ^= 1            # last_block_id leaf: prev-header link
        if trial == 1: mm[1][9, 40] ^= 1           # data_hash aunt
        if trial == 2: eh[3] ^= 1                  # wrong global end header
        if trial == 3: mm[5][0, 0] ^= 1            # wrong batch end header
        if trial == 4: mm[4][2, 1] ^= 1            # wrong batch start header -> reduce link
        if trial == 5: eb += 1000                  # range too long
        got = ctx.prove_data_commitment(1, J, B, *mm, np.array([sb], np.uint64), m.start_header, np.array([eb], np.uint64), eh)
        want = orc.prove_data_commitment(J, B, *mm, sb, m.start_header, eb, eh)
        assert got["fail"][0] == want["fail"] and want["fail"] != 0, trial
        assert (got["map_subchains"][0] == want["map_subchains"]).all(), trial
        assert (got["map_digests"][0] == want["map_digests"]).all(), trial


def test_prove_subchain_batch_many_ranges(ctx, orc):
    """Several independent ranges in one call (the bench shape) + explicit per-job scalars."""
    from blobstreamx_b200 import synthetic as S
    J, B, R = 4, 16, 3
    ms = [S.header_range_inputs(J, B, nb, start=2_000_000 + 1000 * r, seed=S.SEED + r, with_skip=False)[0]
          for r, nb in enumerate((64, 40, 7))]
    cat = lambda f: np.concatenate([getattr(m, f) for m in ms])
    got = ctx.prove_data_commitment(R, J, B, cat("dh_leaf"), cat("dh_aunts"), cat("lb_leaf"), cat("lb_aunts"),
                                    cat("start_headers"), cat("end_headers"),
                                    np.array([m.start_block for m in ms], np.uint64), np.stack([m.start_header for m in ms]),
                                    np.array([m.end_block for m in ms], np.uint64), np.stack([m.end_header for m in ms]))
    for r, m in enumerate(ms):
        want = orc.prove_data_commitment(J, B, *_map_args(m), m.start_block, m.start_header, m.end_block, m.end_header)
        assert got["fail"][r] == 0 == want["fail"]
        assert (got["map_digests"][r] == want["map_digests"]).all()
        assert (got["reduce_nodes"][r] == want["reduce_nodes"]).all()
        assert got["data_commitments"][r].tobytes() == want["data_commitment"]
    # explicit-scalar entry point on range 1
    m = ms[1]
    bs = np.array([m.start_block + j * B for j in range(J)], np.uint64)
    dig, sub = ctx.prove_subchain_batch(B, *_map_args(m), bs, bs + np.uint64(B), np.full(J, m.end_block, np.uint64),
                                        np.tile(m.end_header, (J, 1)))
    assert (dig == got["map_digests"][1]).all() and (sub == got["map_subchains"][1]).all()

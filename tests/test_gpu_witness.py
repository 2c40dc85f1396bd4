"""GPU parity for the witness data formats: HashInputData (row a18) and the byte <-> bit-element encodings."""
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


def _requests(rng, sha512):
    """A request list shaped like the circuits': fixed 35/65/73-byte messages, variable requests over 64-byte
    (sha256) or 188-byte (sha512) buffers, plus edge lengths around the padding boundaries."""
    bufs, kinds, lens = [], [], []
    fixed = [35, 65, 73, 0, 1, 55, 56, 63, 64, 119, 120, 128, 200] if not sha512 else [0, 1, 111, 112, 127, 128, 239, 240, 300]
    for L in fixed * 3:
        bufs.append(rng.bytes(L)); kinds.append(0); lens.append(L)
    buf_len = 188 if sha512 else 64
    for L in list(range(0, buf_len + 1, 7)) + [buf_len, 39, 47, 55, 56] + ([111, 112, 172, 173] if sha512 else []):
        if L <= buf_len:
            bufs.append(rng.bytes(buf_len)); kinds.append(1); lens.append(L)
    if not sha512:  # variable request over a buffer that is not a multiple of 64 (rounded up by curta.rs:154-161)
        bufs.append(rng.bytes(100)); kinds.append(1); lens.append(77)
    offs = np.zeros(len(bufs) + 1, np.uint32)
    offs[1:] = np.cumsum([len(b) for b in bufs])
    return np.frombuffer(b"".join(bufs), np.uint8), offs, np.array(lens, np.uint32), np.array(kinds, np.uint8), bufs


@pytest.mark.parametrize("sha512", [False, True])
def test_hash_input_data(ctx, sha512):
    from oracle import cbind as orc
    rng = np.random.default_rng(21 + sha512)
    flat, offs, lens, kinds, bufs = _requests(rng, sha512)
    got = ctx.hash_input_data(flat, offs, lens, kinds, sha512)
    want = orc.hash_input_data(flat, offs, lens, kinds, sha512)
    for k in ("padded_chunks", "end_bits", "digest_bits", "digest_indices"):
        assert got[k].shape == want[k].shape and (got[k] == want[k]).all(), k
    # the layout is self-consistent: compressing the chunks up to digest_indices[r] gives the request's digest
    h = hashlib.sha512 if sha512 else hashlib.sha256
    chunk = 128 if sha512 else 64
    starts = np.concatenate([[0], np.flatnonzero(got["end_bits"]) + 1])
    for r in (0, 5, len(bufs) - 3, len(bufs) - 1):
        msg = bufs[r][: int(lens[r])]
        c0, c1 = int(starts[r]), int(got["digest_indices"][r])
        raw = got["padded_chunks"][c0:c1 + 1].astype(">u8" if sha512 else ">u4").tobytes()
        # padded prefix == standard padding of the message
        ml = len(msg)
        std = msg + b"\x80" + bytes((-(ml + 1 + chunk // 8)) % chunk) + (ml * 8).to_bytes(chunk // 8, "big")
        assert raw == std, r
        assert h(msg).digest()  # (digest itself is produced by bsx_sha256_batch; covered in test_gpu_parity)


def test_pack_unpack_bytes(ctx):
    rng = np.random.default_rng(3)
    data = rng.integers(0, 256, 100_003, dtype=np.uint8)
    el = ctx.witness_pack_bytes(data)
    want = ((data[:, None] >> np.arange(7, -1, -1)[None, :]) & 1).astype(np.uint64)   # MSB first (vars/byte.rs:49-57)
    assert (el == want).all()
    back, bad = ctx.witness_unpack_bytes(el)
    assert (back == data).all() and not bad
    el[17, 3] = 2
    _, bad = ctx.witness_unpack_bytes(el)
    assert bad

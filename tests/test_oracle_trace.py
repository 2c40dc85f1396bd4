"""The SHA-256 execution trace layout (include/bsx.h BSX_SHA256_TRACE_COLS; oracle/trace.c) pinned by recomputation:
the columns of every row must satisfy the SHA-256 round relations, rows must chain, and the digest read off the trace
(working state after round 63 plus the chaining value) must be hashlib's for every request -- fixed-length, variable-length
and multi-chunk messages built by the HashInputData restatement."""
import hashlib

import numpy as np

from oracle import cbind as orc


def _word(tr, col, row):
    return sum(int(tr[col + k, row]) << (8 * k) for k in range(4))


def _requests(rng, n=24):
    bufs, offs, lens, kinds, msgs = [], [0], [], [], []
    for i in range(n):
        if i % 3 == 0:                                   # variable request over a 64- or 128-byte buffer
            cap = 64 * (1 + i % 2)
            ln = int(rng.integers(0, cap + 1))
            b = rng.bytes(cap)
            bufs.append(b); lens.append(ln); kinds.append(1); msgs.append(b[:ln])
        else:                                            # fixed request (incl. multi-chunk and the 65-byte inner node shape)
            ln = int(rng.choice([1, 35, 55, 56, 64, 65, 73, 119, 120, 200]))
            b = rng.bytes(ln)
            bufs.append(b); lens.append(ln); kinds.append(0); msgs.append(b)
        offs.append(offs[-1] + len(bufs[-1]))
    return b"".join(bufs), offs, lens, kinds, msgs


def test_trace_rows_satisfy_sha256_and_give_the_digests():
    rng = np.random.default_rng(31)
    bufs, offs, lens, kinds, msgs = _requests(rng)
    hid = orc.hash_input_data(np.frombuffer(bufs, np.uint8), offs, lens, kinds)
    n = len(hid["padded_chunks"])
    log_rows = int(np.ceil(np.log2(64 * n)))
    tr = orc.sha256_trace(hid["padded_chunks"], hid["end_bits"], hid["digest_bits"], log_rows)
    assert tr.shape == (orc.SHA256_TRACE_COLS, 1 << log_rows) and not tr[:, 64 * n:].any()
    assert (tr[:100].max() <= 255) and (tr[101:105].max() <= 255)           # byte limbs
    M = 0xFFFFFFFF
    ror = lambda x, k: ((x >> k) | (x << (32 - k))) & M
    iv = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]
    chain, digests = list(iv), {}
    for b in range(n):
        for t in range(64):
            r = 64 * b + t
            st = [_word(tr, 4 + 4 * k, r) for k in range(8)]
            if t == 0:
                assert st == chain and tr[165, r] == 1
            a, bb, c, d, e, f, g, h = st
            w, kt = _word(tr, 0, r), int(tr[175, r])
            assert _word(tr, 48, r) == ror(e, 6) ^ ror(e, 11) ^ ror(e, 25) and _word(tr, 60, r) == ((e & f) ^ (~e & g & M))
            assert _word(tr, 76, r) == ror(a, 2) ^ ror(a, 13) ^ ror(a, 22) and _word(tr, 92, r) == ((a & bb) ^ (a & c) ^ (bb & c))
            t1 = h + _word(tr, 48, r) + _word(tr, 60, r) + kt + w
            assert _word(tr, 96, r) + (int(tr[100, r]) << 32) == t1
            t2 = _word(tr, 76, r) + _word(tr, 92, r)
            assert _word(tr, 101, r) + (int(tr[105, r]) << 32) == t2
            assert _word(tr, 106, r) + (int(tr[110, r]) << 32) == (t1 & M) + (t2 & M)
            assert _word(tr, 111, r) + (int(tr[115, r]) << 32) == d + (t1 & M)
            nxt = [_word(tr, 106, r), a, bb, c, _word(tr, 111, r), e, f, g]
            if t < 63:
                assert [_word(tr, 4 + 4 * k, r + 1) for k in range(8)] == nxt
            if t < 48:                                   # schedule: w_{t+16} appears as w of row t+16
                assert _word(tr, 160, r) == _word(tr, 0, r + 16) and _word(tr, 116, r) == _word(tr, 0, r + 1)
                s0 = ror(_word(tr, 116, r), 7) ^ ror(_word(tr, 116, r), 18) ^ (_word(tr, 116, r) >> 3)
                s1 = ror(_word(tr, 136, r), 17) ^ ror(_word(tr, 136, r), 19) ^ (_word(tr, 136, r) >> 10)
                assert _word(tr, 132, r) == s0 and _word(tr, 152, r) == s1
                assert _word(tr, 160, r) + (int(tr[164, r]) << 32) == s1 + _word(tr, 156, r) + s0 + w
            assert sum(int(tr[169 + k, r]) << k for k in range(6)) == t
        out = [(x + y) & M for x, y in zip(chain, nxt)]
        if tr[168, 64 * b + 63]:
            digests[b] = b"".join(x.to_bytes(4, "big") for x in out)
        chain = list(iv) if tr[167, 64 * b + 63] else out
    want = [hashlib.sha256(m).digest() for m in msgs]
    assert [digests[int(i)] for i in hid["digest_indices"]] == want


def test_sha512_trace_gives_the_digests():
    """SHA-512 trace (the EdDSA accelerator; 80 rows per chunk, 8 byte limbs per word): rows chain, the modular sums carry
    as recorded, and the digests read off the trace equal hashlib's for R ‖ A ‖ M messages of 64 .. 188 bytes."""
    rng = np.random.default_rng(32)
    msgs = [rng.bytes(64 + int(rng.integers(0, 125))) for _ in range(9)] + [b"", rng.bytes(111), rng.bytes(112), rng.bytes(239), rng.bytes(240)]
    bufs, offs, lens, kinds = [], [0], [], []
    for i, m in enumerate(msgs):
        if i % 2 == 0 and len(m) <= 188:                       # variable request over the 188-byte EdDSA buffer
            b = m + rng.bytes(188 - len(m))
            bufs.append(b); lens.append(len(m)); kinds.append(1)
        else:
            bufs.append(m); lens.append(len(m)); kinds.append(0)
        offs.append(offs[-1] + len(bufs[-1]))
    hid = orc.hash_input_data(np.frombuffer(b"".join(bufs), np.uint8), offs, lens, kinds, sha512=True)
    n = len(hid["padded_chunks"])
    log_rows = int(np.ceil(np.log2(80 * n)))
    tr = orc.sha512_trace(hid["padded_chunks"], hid["end_bits"], hid["digest_bits"], log_rows)
    assert tr.shape == (orc.SHA512_TRACE_COLS, 1 << log_rows) and not tr[:, 80 * n:].any()
    M = (1 << 64) - 1
    word = lambda col, r: sum(int(tr[col + k, r]) << (8 * k) for k in range(8))
    ror = lambda x, k: ((x >> k) | (x << (64 - k))) & M
    iv = [0x6a09e667f3bcc908, 0xbb67ae8584caa73b, 0x3c6ef372fe94f82b, 0xa54ff53a5f1d36f1, 0x510e527fade682d1, 0x9b05688c2b3e6c1f,
          0x1f83d9abfb41bd6b, 0x5be0cd19137e2179]
    chain, digests = list(iv), {}
    for b in range(n):
        for t in range(80):
            r = 80 * b + t
            st = [word(8 + 8 * k, r) for k in range(8)]
            if t == 0:
                assert st == chain
            a, bb, c, d, e, f, g, h = st
            w, kt = word(0, r), int(tr[336, r]) | (int(tr[337, r]) << 32)
            assert word(96, r) == ror(e, 14) ^ ror(e, 18) ^ ror(e, 41) and word(120, r) == ((e & f) ^ (~e & g & M))
            assert word(152, r) == ror(a, 28) ^ ror(a, 34) ^ ror(a, 39) and word(184, r) == ((a & bb) ^ (a & c) ^ (bb & c))
            t1 = h + word(96, r) + word(120, r) + kt + w
            assert word(192, r) + (int(tr[200, r]) << 64) == t1
            t2 = word(152, r) + word(184, r)
            assert word(201, r) + (int(tr[209, r]) << 64) == t2
            assert word(210, r) + (int(tr[218, r]) << 64) == (t1 & M) + (t2 & M)
            assert word(219, r) + (int(tr[227, r]) << 64) == d + (t1 & M)
            nxt = [word(210, r), a, bb, c, word(219, r), e, f, g]
            if t < 79:
                assert [word(8 + 8 * k, r + 1) for k in range(8)] == nxt
            if t < 64:
                assert word(316, r) == word(0, r + 16)
                s0 = ror(word(228, r), 1) ^ ror(word(228, r), 8) ^ (word(228, r) >> 7)
                s1 = ror(word(268, r), 19) ^ ror(word(268, r), 61) ^ (word(268, r) >> 6)
                assert word(260, r) == s0 and word(300, r) == s1
                assert word(316, r) + (int(tr[324, r]) << 64) == s1 + word(308, r) + s0 + w
        out = [(x + y) & M for x, y in zip(chain, nxt)]
        if tr[328, 80 * b + 79]:
            digests[b] = b"".join(x.to_bytes(8, "big") for x in out)
        chain = list(iv) if tr[327, 80 * b + 79] else out
    assert [digests[int(i)] for i in hid["digest_indices"]] == [hashlib.sha512(m).digest() for m in msgs]

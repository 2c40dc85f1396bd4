#!/usr/bin/env python
"""One-off stress run (GPU box): many random Ed25519 inputs -- valid signatures, corrupted ones, random garbage, edge
encodings (y = 0, 1, p-1, p, 2^255-1, small-order points, s >= l) -- through both kernel paths, every 576-byte record
compared with the oracle.  usage: python scripts/stress_ed25519.py [n_valid] [n_garbage]"""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blobstreamx_b200 import lib, synthetic as S   # noqa: E402
from oracle import cbind as orc                     # noqa: E402

P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493


def main():
    n_valid = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    n_garb = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
    rng = np.random.default_rng(20261017)
    base = S.ed25519_batch_inputs(2000, inactive_every=50)
    idx = rng.integers(0, 2000, n_valid)
    pks, sigs, msgs, lens, act = (np.ascontiguousarray(a[idx]) for a in base)
    # corrupt a third of the valid ones in one random bit of pk / R / s / message
    for i in range(0, n_valid, 3):
        which = rng.integers(0, 4)
        tgt = (pks, sigs, sigs, msgs)[which]
        lo, hi = ((0, 32), (0, 32), (32, 64), (0, max(1, int(lens[i]))))[which]
        tgt[i, rng.integers(lo, hi)] ^= np.uint8(1 << rng.integers(0, 8))
    g_pk = rng.integers(0, 256, (n_garb, 32), dtype=np.uint8)
    g_sig = rng.integers(0, 256, (n_garb, 64), dtype=np.uint8)
    g_msg = rng.integers(0, 256, (n_garb, 124), dtype=np.uint8)
    g_len = rng.integers(0, 125, n_garb).astype(np.uint32)
    edge_y = [0, 1, 2, P - 1, P, P + 1, 2**255 - 1, 2**255 - 19 + 18, 19, 2**254, (1 << 255) | 1, (1 << 255), 2**256 - 1,
              0x7a03ac9277fdc74ec6cc392cfa53202a0f67100d760b3cba4fd84d3d706a17c7]   # a small-order y
    edge_s = [0, 1, L - 1, L, L + 1, 2**252, 2**253 - 1, 2**256 - 1]
    for j in range(min(n_garb, 4000)):
        if j % 2 == 0:
            g_pk[j] = np.frombuffer((edge_y[(j // 2) % len(edge_y)] % 2**256).to_bytes(32, "little"), np.uint8)
        else:
            g_sig[j, :32] = np.frombuffer((edge_y[(j // 2) % len(edge_y)] % 2**256).to_bytes(32, "little"), np.uint8)
        g_sig[j, 32:] = np.frombuffer(edge_s[j % len(edge_s)].to_bytes(32, "little"), np.uint8)
    pks = np.concatenate([pks, g_pk]); sigs = np.concatenate([sigs, g_sig]); msgs = np.concatenate([msgs, g_msg])
    lens = np.concatenate([lens, g_len]); act = np.concatenate([act, np.ones(n_garb, np.uint8)])
    n = len(pks)
    t0 = time.time()
    want = orc.ed25519_batch(pks, sigs, msgs, lens, act, threads=os.cpu_count())
    t_cpu = time.time() - t0
    ctx = lib.Context(0)
    got_mono = ctx.ed25519_batch(pks, sigs, msgs, lens, act)                       # n > 16384: thread-per-signature kernel
    bad = np.nonzero((got_mono != want).any(axis=1))[0]
    assert bad.size == 0, ("mono", bad[:10])
    step = 8000                                                                    # quad-lane path in slices
    for lo in range(0, n, step):
        sl = slice(lo, min(n, lo + step))
        got = ctx.ed25519_batch(pks[sl], sigs[sl], msgs[sl], lens[sl], act[sl])
        bad = np.nonzero((got != want[sl]).any(axis=1))[0]
        assert bad.size == 0, ("quad", lo, bad[:10])
    flags = want[:, 520]
    print(f"stress ok: {n} records bit-exact on both paths (oracle {t_cpu:.1f} s); flags histogram:",
          {int(f): int((flags == f).sum()) for f in np.unique(flags)})


if __name__ == "__main__":
    main()

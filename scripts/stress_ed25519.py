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
    from tests._ed_cases import stress_inputs
    pks, sigs, msgs, lens, act = stress_inputs(n_valid, n_garb)
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

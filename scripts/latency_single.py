#!/usr/bin/env python
"""Latency of ONE proof's witness through the host-buffer C ABI (what an operator proving one range at a time sees):
header_range_1024 (skip + 32 x 32 map + reduce) and next_header, median of 50 calls, next to the single-thread oracle."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                        # noqa: E402
from blobstreamx_b200 import lib, synthetic as S   # noqa: E402
from oracle import cbind as orc                     # noqa: E402


def med(fn, n=50):
    fn(); fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3


def main():
    ctx = lib.Context(0)
    vs = S.ValidatorSet.make()
    m, skip, chain = S.header_range_inputs(32, 32, None, valset=vs)
    mm = bench.tile_ranges([m], 1)
    mm["n_jobs"], mm["batch"] = 32, 32
    t_hr = med(lambda: ctx.header_range([skip], mm))
    t0 = time.perf_counter()
    orc.verify_skip(skip, threads=1)
    orc.prove_data_commitment(32, 32, m.dh_leaf, m.dh_aunts, m.lb_leaf, m.lb_aunts, m.start_headers, m.end_headers, m.start_block,
                              m.start_header, m.end_block, m.end_header, threads=1)
    t_hr_cpu = (time.perf_counter() - t0) * 1e3
    t_skip = med(lambda: ctx.verify_skip([skip]))
    ed_in = S.ed25519_batch_inputs(100)
    t_ed = med(lambda: ctx.ed25519_batch(*ed_in))
    print(f"header_range_1024, one proof: {t_hr:.3f} ms (oracle, 1 thread: {t_hr_cpu:.1f} ms)   verify_skip alone: {t_skip:.3f} ms   "
          f"100 signatures: {t_ed:.3f} ms")


if __name__ == "__main__":
    main()

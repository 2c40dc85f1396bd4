#!/usr/bin/env python
"""Best-effort CPU library baseline for header_range_1024 (baseline B): OpenSSL SHA-256 through hashlib and
libsodium Ed25519 verification through PyNaCl, one process per host core.  Per range it hashes messages of the sizes the
witness schedule holds (20 969 digests: 1 024 leaves of 35 B, 1 024 of 73 B, 200 validator leaves of ~45 B, the rest 65-byte
inner nodes / tuple leaves) and verifies 100 signatures over 108-byte votes.  It produces digests and accept / reject only --
none of the EC intermediates, quotients or the request-order layout of the witness -- so it bounds what tuned CPU libraries
could do for the same inputs, not the reference's path.  Prints one JSON line.  No CUDA, no oracle."""
import argparse
import hashlib
import json
import multiprocessing as mp
import os
import time


def _range_messages():
    sizes = [35] * 1024 + [73] * 1024 + [45] * 200 + [65] * (20969 - 2248)
    blob = os.urandom(sum(sizes))
    out, at = [], 0
    for s in sizes:
        out.append(blob[at:at + s])
        at += s
    return out


def _worker(n_ranges: int) -> float:
    from nacl.signing import SigningKey
    msgs = _range_messages()
    keys = [SigningKey.generate() for _ in range(100)]
    votes = [os.urandom(108) for _ in range(100)]
    sigs = [k.sign(v).signature for k, v in zip(keys, votes)]
    vks = [k.verify_key for k in keys]
    sha = hashlib.sha256
    t0 = time.perf_counter()
    for _ in range(n_ranges):
        for m in msgs:
            sha(m).digest()
        for vk, v, s in zip(vks, votes, sigs):
            vk.verify(v, s)
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranges-per-core", type=int, default=40)
    args = ap.parse_args()
    cores = len(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    with mp.Pool(cores) as pool:
        pool.map(_worker, [1] * cores)                      # start the processes, import libsodium
        t0 = time.perf_counter()
        busy = pool.map(_worker, [args.ranges_per_core] * cores)
        wall = time.perf_counter() - t0
    ranges = args.ranges_per_core * cores
    print(json.dumps({"value": ranges * 1024 / wall, "unit": "headers/s", "cores": cores, "kind": "library",
                      "sample": f"{ranges} ranges x (20 969 hashlib SHA-256 digests + 100 PyNaCl verifications), {cores} processes, "
                                f"{wall:.2f} s wall, {max(busy):.2f} s in the slowest worker",
                      "note": "digests and accept/reject only: no EC intermediates, no quotients, no witness layout"}))


if __name__ == "__main__":
    main()

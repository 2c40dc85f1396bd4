"""PCIe copy bandwidth on the GPU box with pinned host memory: H2D alone, D2H alone, both at once on two streams --
for ONE GPU, or (--ranks N) for N GPUs AT THE SAME TIME, one process per GPU, each bound to the CPUs next to its GPU before
it allocates its pinned buffers (as bench.py does).  The N-rank numbers are the ceiling of bench.py's end-to-end leg:
every rank moves its inputs up and its witness down over its own x16 link, but all links end in the same host memory.
usage: python scripts/ubench/pcie.py [--ranks N] [--no-bind]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def worker(rank, world, bind, barrier, q):
    import torch
    torch.cuda.set_device(rank)
    numa = None
    if bind:
        import bench
        numa = bench.bind_to_gpu_numa(torch, rank)
    n = 192 << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(1); h_out.fill_(2)                       # first touch on this rank's NUMA node
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    def both():
        h2d(); d2h()

    res = {"rank": rank, "cpus": numa}
    for name, fn, nbytes in (("h2d", h2d, n), ("d2h", d2h, n), ("duplex", both, 2 * n)):
        fn(); torch.cuda.synchronize()
        barrier.wait()                                  # every rank starts the same copy at the same time
        t0 = time.perf_counter()
        for _ in range(8):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 8
        barrier.wait()
        res[name + "_GBps"] = nbytes / dt / 1e9
    q.put(res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", type=int, default=1)
    ap.add_argument("--no-bind", action="store_true")
    a = ap.parse_args()
    import torch.multiprocessing as mp
    mp.set_start_method("spawn")
    barrier, q = mp.Barrier(a.ranks), mp.Queue()
    ps = [mp.Process(target=worker, args=(r, a.ranks, not a.no_bind, barrier, q)) for r in range(a.ranks)]
    for p in ps:
        p.start()
    out = sorted((q.get() for _ in ps), key=lambda r: r["rank"])
    for p in ps:
        p.join()
    agg = {k: sum(r[k] for r in out) for k in ("h2d_GBps", "d2h_GBps", "duplex_GBps")}
    print(json.dumps({"ranks": a.ranks, "bound_to_gpu_numa": not a.no_bind, "host_cpus": len(os.sched_getaffinity(0)),
                      "aggregate": agg, "per_rank": out}))


if __name__ == "__main__":
    main()

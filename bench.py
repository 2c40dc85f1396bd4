#!/usr/bin/env python
"""bench.py -- header_range_1024 witness-generation throughput (headers/sec) on N B200s.

A "step" is one pass of the hot path over one batch of R independent synthetic header ranges
(32 map jobs x 32 headers each = BASELINE config 2), all hint-level witness values produced:
every SHA-256 digest of the map circuits in Curta request order, the 32 subchain records, the
31 reduce nodes and the data commitment per range (and, once `--skip` is on, the verify_skip
digests + Ed25519 records of the target header).

  value      whole-job headers/s, inputs resident in HBM, CUDA events, max over ranks
  e2e        same metric through the host-buffer C-ABI entry point (pinned host buffers, H2D + D2H inside)
  roofline   the dominant kernel (prove_subchain_kernel<32>) against the measured HBM peak
  cpu_baseline  the CPU oracle ("port" of the reference's witness path) on a bounded sample, 1 thread

`--impl reference` times the CPU restatement of the reference's path on all host cores
(the reference itself is Rust + un-vendored git deps and cannot be built here).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_JOBS, BATCH = 32, 32                      # header_range_1024 (BX/bin/header_range_1024.rs:7-16)
HEADERS_PER_RANGE = N_JOBS * BATCH
METRIC = "headers/sec, header_range_1024 witness-gen"
UNIT = "headers/s"


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def make_ranges(n_distinct: int, n_jobs: int = N_JOBS, batch: int = BATCH):
    """n_distinct independent synthetic chains (seeded, hash-linked); host-side generation only."""
    from blobstreamx_b200 import synthetic as S
    vs = S.ValidatorSet.make(S.SEED)
    return [S.header_range_inputs(n_jobs, batch, None, start=1_000_000 + 10_000 * r, seed=S.SEED + r, valset=vs,
                                  with_skip=False)[0] for r in range(n_distinct)]


FIELDS = ("dh_leaf", "dh_aunts", "lb_leaf", "lb_aunts", "start_headers", "end_headers")


def tile_ranges(ms, R):
    """Concatenate R ranges (cycling over the distinct chains) into the flat arrays of the C ABI."""
    pick = [ms[r % len(ms)] for r in range(R)]
    a = {f: np.concatenate([getattr(m, f) for m in pick]) for f in FIELDS}
    a["start_blocks"] = np.array([m.start_block for m in pick], np.uint64)
    a["end_blocks"] = np.array([m.end_block for m in pick], np.uint64)
    a["start_header"] = np.stack([m.start_header for m in pick])
    a["end_header"] = np.stack([m.end_header for m in pick])
    return a


def out_shapes(R, J=N_JOBS, B=BATCH):
    return dict(map_digests=(R, J, 20 * B - 1, 32), map_subchains=(R, J, 128), reduce_digests=(R, J - 1, 32),
                reduce_nodes=(R, J - 1, 128), data_commitments=(R, 32))


def algorithmic_bytes_map(R, J=N_JOBS, B=BATCH):
    """SHA-256 = 64*blk + 32*digests; per map job (39B-2) blk and (20B-1) digests."""
    return R * J * (64 * (39 * B - 2) + 32 * (20 * B - 1))


# ------------------------------------------------------------------------------------------------
# clocks (nvml, sampled DURING the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self._nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle)
# ------------------------------------------------------------------------------------------------
def cpu_ranges_per_sec(ms, n_ranges: int, threads: int) -> float:
    from oracle import cbind as orc
    t0 = time.perf_counter()
    for r in range(n_ranges):
        m = ms[r % len(ms)]
        w = orc.prove_data_commitment(m.n_jobs, m.batch_size, m.dh_leaf, m.dh_aunts, m.lb_leaf, m.lb_aunts, m.start_headers,
                                      m.end_headers, m.start_block, m.start_header, m.end_block, m.end_header, threads=threads)
        assert w["fail"] == 0
    return n_ranges / (time.perf_counter() - t0)


def run_reference(args):
    """The reference's CPU path (restated: oracle/), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cbind as orc
    threads = min(orc.max_threads(), len(os.sched_getaffinity(0)))
    ms = make_ranges(2)
    sample = args.cpu_ranges
    for _ in range(args.warmup):
        cpu_ranges_per_sec(ms, 1, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_ranges_per_sec(ms, sample, threads)
    dt = time.perf_counter() - t0
    v = args.steps * sample * HEADERS_PER_RANGE / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"header_range_1024 witness-gen (32 map jobs x 32 headers + reduce), {sample} ranges/step on CPU"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} ranges x 1024 headers per step, OpenMP over map jobs"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from blobstreamx_b200 import lib
    from blobstreamx_b200.lib import ptr, u32

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    R = args.ranges
    ctx = lib.Context(local)
    ms = make_ranges(args.distinct)
    host = tile_ranges(ms, R)
    stream = torch.cuda.current_stream().cuda_stream

    d_in = {k: torch.from_numpy(v.view(np.uint8).reshape(-1)).to(dev) for k, v in host.items()}
    shapes = out_shapes(R)
    d_out = {k: torch.zeros(int(np.prod(s)), dtype=torch.uint8, device=dev) for k, s in shapes.items()}
    d_fail = torch.zeros(R, dtype=torch.int32, device=dev)
    P = lambda t: ptr(t.data_ptr())

    def step_dev():
        ctx.call_dev("bsx_prove_data_commitment_dev", stream, u32(R), u32(N_JOBS), u32(BATCH), P(d_in["dh_leaf"]),
                     P(d_in["dh_aunts"]), P(d_in["lb_leaf"]), P(d_in["lb_aunts"]), P(d_in["start_headers"]),
                     P(d_in["end_headers"]), P(d_in["start_blocks"]), P(d_in["start_header"]), P(d_in["end_blocks"]),
                     P(d_in["end_header"]), P(d_out["map_digests"]), P(d_out["map_subchains"]), P(d_out["reduce_digests"]),
                     P(d_out["reduce_nodes"]), P(d_out["data_commitments"]), P(d_fail))

    # explicit per-job scalars for the map-only launch (the dominant kernel timed alone)
    jb = (host["start_blocks"][:, None] + np.arange(N_JOBS, dtype=np.uint64)[None, :] * np.uint64(BATCH)).reshape(-1)
    d_bs = torch.from_numpy(jb.view(np.uint8)).to(dev)
    d_be = torch.from_numpy((jb + np.uint64(BATCH)).view(np.uint8)).to(dev)
    d_ge = torch.from_numpy(np.repeat(host["end_blocks"], N_JOBS).view(np.uint8)).to(dev)
    d_geh = torch.from_numpy(np.repeat(host["end_header"], N_JOBS, axis=0).reshape(-1)).to(dev)

    def map_only():
        ctx.call_dev("bsx_prove_subchain_batch_dev", stream, u32(BATCH), u32(R * N_JOBS), P(d_in["dh_leaf"]),
                     P(d_in["dh_aunts"]), P(d_in["lb_leaf"]), P(d_in["lb_aunts"]), P(d_in["start_headers"]),
                     P(d_in["end_headers"]), P(d_bs), P(d_be), P(d_ge), P(d_geh), P(d_out["map_digests"]),
                     P(d_out["map_subchains"]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness gate before timing: rank 0 checks range 0 against the oracle ----
    step_dev()
    torch.cuda.synchronize()
    if int(d_fail.abs().sum().item()) != 0:
        raise SystemExit("bench.py: circuit assertions failed on the synthetic workload")
    if rank == 0 and not args.no_check:
        from oracle import cbind as orc
        m = ms[0]
        w = orc.prove_data_commitment(N_JOBS, BATCH, m.dh_leaf, m.dh_aunts, m.lb_leaf, m.lb_aunts, m.start_headers,
                                      m.end_headers, m.start_block, m.start_header, m.end_block, m.end_header)
        g = d_out["map_digests"][: N_JOBS * (20 * BATCH - 1) * 32].cpu().numpy().reshape(N_JOBS, 20 * BATCH - 1, 32)
        assert (g == w["map_digests"]).all(), "GPU map digests differ from the oracle"
        assert d_out["data_commitments"][:32].cpu().numpy().tobytes() == w["data_commitment"]

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        step_dev()
    barrier()
    l0 = ctx.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with ClockSampler(local) as clk:
        ev[0].record()
        for _ in range(args.steps):
            step_dev()
        ev[1].record()
        barrier()
    launches = ctx.launch_count - l0
    ms_total = ev[0].elapsed_time(ev[1])
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * R * HEADERS_PER_RANGE / (ms_step * 1e-3)

    # ---- dominant kernel alone (roofline) ----
    for _ in range(2):
        map_only()
    torch.cuda.synchronize()
    ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev2[0].record()
    for _ in range(args.steps):
        map_only()
    ev2[1].record()
    torch.cuda.synchronize()
    k_ms = ev2[0].elapsed_time(ev2[1]) / args.steps
    alg = algorithmic_bytes_map(R)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg / (k_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            traffic = json.load(f).get("prove_subchain_kernel<32>", {}).get("dram_bytes_per_launch")
    except Exception:
        pass

    # ---- end to end through the host-buffer C ABI (pinned host memory, H2D + kernels + D2H) ----
    Re = min(R, args.e2e_ranges)
    hp = {k: torch.from_numpy(v.view(np.uint8).reshape(-1)[: v.view(np.uint8).size * Re // R].copy()).pin_memory()
          for k, v in host.items()}
    ho = {k: torch.zeros(int(np.prod(s)) * Re // R, dtype=torch.uint8).pin_memory() for k, s in shapes.items()}
    hfail = torch.zeros(Re, dtype=torch.int32).pin_memory()
    HP = lambda t: ptr(t.data_ptr())

    def step_e2e():
        ctx._call("bsx_prove_data_commitment", u32(Re), u32(N_JOBS), u32(BATCH), HP(hp["dh_leaf"]), HP(hp["dh_aunts"]),
                  HP(hp["lb_leaf"]), HP(hp["lb_aunts"]), HP(hp["start_headers"]), HP(hp["end_headers"]), HP(hp["start_blocks"]),
                  HP(hp["start_header"]), HP(hp["end_blocks"]), HP(hp["end_header"]), HP(ho["map_digests"]),
                  HP(ho["map_subchains"]), HP(ho["reduce_digests"]), HP(ho["reduce_nodes"]), HP(ho["data_commitments"]),
                  HP(hfail))

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    assert int(hfail.abs().sum().item()) == 0
    e2e = world * Re * HEADERS_PER_RANGE * args.steps / e2e_dt
    h2d = sum(t.numel() for t in hp.values())
    d2h = sum(t.numel() for t in ho.values()) + hfail.numel() * 4

    # ---- CPU oracle beside it (rank 0, N=1 only, bounded sample) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rps = cpu_ranges_per_sec(ms, args.cpu_ranges, 1)
        cpu = {"value": rps * HEADERS_PER_RANGE, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{args.cpu_ranges} ranges x 1024 headers (same generator), single thread like the reference's witness loop"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": f"header_range_1024 witness-gen: {R} independent ranges/step/GPU x (32 map jobs x 32 headers "
                                   f"+ 31 reduce nodes), all {N_JOBS * (20 * BATCH - 1) + N_JOBS - 1} SHA-256 digests per range written",
                       "ranges_per_step_per_gpu": R, "distinct_chains": args.distinct,
                       "l2": f"inputs+outputs per step = {(sum(t.numel() for t in d_in.values()) + sum(t.numel() for t in d_out.values())) / 1e6:.0f} MB > 126 MB L2 (no flush needed)"},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": {"kernel": "prove_subchain_kernel<32>", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "algorithmic_bytes_per_launch": alg, "kernel_ms": k_ms,
                         "note": "SHA-256 is ~25 int ops/byte: the int32 ALU pipe, not HBM, is the physical bound (DESIGN.md)"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ranges_per_step": Re},
            "cpu_baseline": cpu,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_ed25519(args):
    """Config 5: Ed25519 witness batch, sigs/s (device-resident inputs), next to the CPU oracle."""
    import torch
    from blobstreamx_b200 import lib, synthetic as S
    from blobstreamx_b200.lib import ptr, u32
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    ctx = lib.Context(0)
    stream = torch.cuda.current_stream().cuda_stream
    base = S.ed25519_batch_inputs(min(args.sigs, 2000))
    rep = -(-args.sigs // len(base[0]))
    pks, sigs, msgs, lens, active = (np.concatenate([a] * rep)[: args.sigs] for a in base)
    d = [torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev) for a in (pks, sigs, msgs, lens, active)]
    out = torch.zeros(args.sigs * 576, dtype=torch.uint8, device=dev)
    P = lambda t: ptr(t.data_ptr())

    def step():
        ctx.call_dev("bsx_ed25519_batch_dev", stream, u32(args.sigs), P(d[0]), P(d[1]), P(d[2]), u32(124), P(d[3]), P(d[4]), P(out))

    step()
    torch.cuda.synchronize()
    rec = out.cpu().numpy().reshape(-1, 576)
    assert (rec[:, 520] == 0xF).all(), "signatures did not verify on the GPU"
    if not args.no_check:
        from oracle import cbind as orc
        k = min(args.sigs, 200)
        want = orc.ed25519_batch(pks[:k], sigs[:k], msgs[:k], lens[:k], active[:k], threads=8)
        assert (rec[:k] == want).all(), "GPU Ed25519 records differ from the oracle"
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with ClockSampler(0) as clk:
        ev[0].record()
        for _ in range(args.steps):
            step()
        ev[1].record()
        torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / args.steps
    cpu = None
    if not args.no_cpu:
        from oracle import cbind as orc
        k = min(args.sigs, 200)
        t0 = time.perf_counter()
        orc.ed25519_batch(pks[:k], sigs[:k], msgs[:k], lens[:k], active[:k], threads=1)
        cpu = {"value": k / (time.perf_counter() - t0), "unit": "sigs/s", "cores": 1, "kind": "port", "sample": f"{k} signatures"}
    print(json.dumps({"metric": "sigs/sec, Ed25519 witness batch", "value": args.sigs / (ms * 1e-3), "unit": "sigs/s",
                      "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                      "dtype": "i32x10 limbs / i64 accumulate", "data": "synthetic",
                      "config": {"workload": f"ed25519 witness batch, {args.sigs} signatures/step (CanonicalVote messages, 1% dummy lanes)"},
                      "gpu_launches": args.steps, "clocks": clk.summary(), "cpu_baseline": cpu}))


def run_gates(args):
    """constraints/sec: U32ArithmeticGate (3 ops/row, 114 wires, 108 constraints) over 2^k rows, HBM roofline."""
    import torch
    from blobstreamx_b200 import lib
    from blobstreamx_b200.lib import ptr, u32
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    ctx = lib.Context(0)
    stream = torch.cuda.current_stream().cuda_stream
    rows, gate, p0, p1 = args.rows, 0, 3, 0
    nw, ncn = ctx.gate_num_wires(gate, p0, p1), ctx.gate_num_constraints(gate, p0, p1)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    wires = torch.zeros(nw * rows, dtype=torch.int64, device=dev)
    wv = wires.view(nw, rows)
    for i in range(p0):   # random u32 inputs of valid operations; the generator kernel fills the rest
        wv[6 * i:6 * i + 3] = torch.randint(0, 2**32, (3, rows), generator=g, device=dev, dtype=torch.int64)
    cons = torch.zeros(ncn * rows, dtype=torch.int64, device=dev)
    P = lambda t: ptr(t.data_ptr())
    ctx.call_dev("bsx_gl_gate_witness_dev", stream, u32(gate), u32(p0), u32(p1), P(wires), u32(rows))

    def step():
        ctx.call_dev("bsx_gl_gate_eval_dev", stream, u32(gate), u32(p0), u32(p1), P(wires), u32(rows), P(cons))

    step()
    torch.cuda.synchronize()
    assert int(cons.abs().max().item()) == 0, "valid witness must satisfy every constraint"
    if not args.no_check:
        from oracle import cbind as orc
        k = 512
        sub = wv[:, :k].contiguous().cpu().numpy().view(np.uint64)
        sub[5, 7] += np.uint64(1)  # one broken row so the comparison is not all zeros
        got = ctx.gl_gate_eval(gate, p0, p1, sub)
        assert (got == orc.gate_eval(gate, p0, p1, sub, threads=4)).all() and got[:, 7].any()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with ClockSampler(0) as clk:
        ev[0].record()
        for _ in range(args.steps):
            step()
        ev[1].record()
        torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / args.steps
    alg = 8 * (nw + ncn) * rows
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    cpu = None
    if not args.no_cpu:
        from oracle import cbind as orc
        k = 1 << 14
        sub = wv[:, :k].contiguous().cpu().numpy().view(np.uint64)
        t0 = time.perf_counter()
        orc.gate_eval(gate, p0, p1, sub, threads=1)
        cpu = {"value": ncn * k / (time.perf_counter() - t0), "unit": "constraints/s", "cores": 1, "kind": "port", "sample": f"{k} rows"}
    print(json.dumps({"metric": "constraints/sec, U32ArithmeticGate eval_unfiltered_base_batch", "value": ncn * rows / (ms * 1e-3),
                      "unit": "constraints/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                      "higher_is_better": True, "dtype": "u64 mod 2^64-2^32+1", "data": "synthetic",
                      "config": {"workload": f"U32ArithmeticGate num_ops=3, {rows} rows x 114 wires -> 108 constraints/row",
                                 "l2": f"{alg / 1e6:.0f} MB per step > 126 MB L2"},
                      "gpu_launches": args.steps, "clocks": clk.summary(),
                      "roofline": {"kernel": "gl_gate_eval_kernel<0>", "bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak,
                                   "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                                   "algorithmic_bytes_per_launch": alg},
                      "cpu_baseline": cpu}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="header_range", choices=["header_range", "ed25519", "gates"])
    ap.add_argument("--rows", type=int, default=1 << 20)
    ap.add_argument("--sigs", type=int, default=100000)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ranges", type=int, default=256, help="independent header ranges per step per GPU")
    ap.add_argument("--distinct", type=int, default=8, help="distinct synthetic chains (tiled to --ranges)")
    ap.add_argument("--e2e-ranges", type=int, default=64)
    ap.add_argument("--cpu-ranges", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "ed25519":
        run_ed25519(args)
    elif args.mode == "gates":
        run_gates(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

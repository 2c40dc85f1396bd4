"""Multi-GPU header_range: the map jobs of prove_data_commitment sharded over ranks, ONE all-gather of the
per-job MapReduceSubchainVariable records, then the reduce tree.

Reference shape: `MapReduceGenerator::run_once` proves the NB_MAP_JOBS map circuits independently
(PX/frontend/mapreduce/generator.rs:97-111) -- each consumes only the shared ctx (start/end block + header) and its
own BATCH_SIZE headers -- and then log2(jobs) reduce layers (generator.rs:113-151, closure BX/circuits/builder.rs:
337-395).  The reference's only multi-worker story is `PROVER=remote` (HTTP, PX/backend/prover/remote.rs:98-153).
Here rank r of W owns jobs [r*J/W, (r+1)*J/W) of EVERY range in flight (contiguous job slice = contiguous header
range) and reduces ranges [r*R/W, (r+1)*R/W).  The exchange between the two is 128 bytes per job:
  * on GPUs the engine is libbsx's `bsx_shard_*` C ABI (csrc/k_shard.cu; this module only sets it up and calls
    `bsx_shard_step_dev`): the kernel that computes a record stores it straight into the reducing rank's gathered array
    through NVLink peer memory, its last CTA publishes a per-(rank, step) flag on every peer and the reduce kernel
    acquires the flags -- compute, exchange and synchronisation are the map and reduce kernels themselves; no collective
    and no barrier kernel competes with the Ed25519 CTAs for an SM (the NCCL all-gather could only start once an SM had
    drained: 8 GPUs 3.20 ms per step; the separate barrier kernel of round 1 made the step time unstable by +-30 %).
    The peers' exchange buffers are mapped with CUDA IPC handles (exchanged once over the process group); if that is
    refused, with torch symmetric memory;
  * if neither works, and in the CPU tests (gloo), one `all_gather_into_tensor`.
No other collective is on the data path.

The compute backend is injected: `CudaBackend` (below) drives libbsx through device pointers on torch's current
stream; the CPU tests pass an oracle-backed stand-in that lives under tests/ (the product has no CPU path).
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch
import torch.distributed as dist

SUBCHAIN_BYTES = 128
MAP_FIELDS = ("dh_leaf", "dh_aunts", "lb_leaf", "lb_aunts", "start_headers", "end_headers")
_ROW = {"dh_leaf": 34, "dh_aunts": 128, "lb_leaf": 72, "lb_aunts": 128}


def job_slice(n_jobs: int, rank: int, world: int) -> slice:
    if n_jobs % world:
        raise ValueError(f"NB_MAP_JOBS={n_jobs} is not divisible by world size {world}")
    per = n_jobs // world
    return slice(rank * per, (rank + 1) * per)


def shard_map_inputs(host: Dict[str, np.ndarray], n_ranges: int, n_jobs: int, batch: int, rank: int, world: int):
    """host: the flat arrays of ALL `n_ranges` ranges (range-major, job-major; see include/bsx.h
    bsx_prove_data_commitment).  Returns this rank's job slice of every range plus the explicit per-job scalars
    bsx_prove_subchain_batch wants."""
    js = job_slice(n_jobs, rank, world)
    per = js.stop - js.start
    out = {}
    for f in ("dh_leaf", "dh_aunts", "lb_leaf", "lb_aunts"):
        a = np.ascontiguousarray(host[f], np.uint8).reshape(n_ranges, n_jobs, batch, _ROW[f])
        out[f] = np.ascontiguousarray(a[:, js]).reshape(n_ranges * per * batch, _ROW[f])
    for f in ("start_headers", "end_headers"):
        a = np.ascontiguousarray(host[f], np.uint8).reshape(n_ranges, n_jobs, 32)
        out[f] = np.ascontiguousarray(a[:, js]).reshape(n_ranges * per, 32)
    sb = np.ascontiguousarray(host["start_blocks"], np.uint64).reshape(n_ranges)
    eb = np.ascontiguousarray(host["end_blocks"], np.uint64).reshape(n_ranges)
    j = np.arange(js.start, js.stop, dtype=np.uint64)
    bs = (sb[:, None] + j[None, :] * np.uint64(batch)).reshape(-1)
    out["batch_start"] = bs
    out["batch_end"] = bs + np.uint64(batch)
    out["global_end"] = np.repeat(eb, per)
    out["global_end_header"] = np.repeat(np.ascontiguousarray(host["end_header"], np.uint8).reshape(n_ranges, 32), per, axis=0)
    return out


class CudaBackend:
    """libbsx through the `_dev` entry points on torch's current CUDA stream."""

    def __init__(self, device_index: int):
        from . import lib
        self.lib = lib
        self.ctx = lib.Context(device_index)
        self.device = torch.device("cuda", device_index)

    def tensor(self, a: np.ndarray) -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(self.device)

    def empty(self, nbytes: int) -> torch.Tensor:
        return torch.zeros(nbytes, dtype=torch.uint8, device=self.device)

    def shard(self, rank, world, n_ranges, n_jobs, batch, exchange_buf=0):
        return self.lib.Shard(self.ctx, rank, world, n_ranges, n_jobs, batch, exchange_buf)

    def map(self, B, n_jobs, t, digests, subchains):
        P = lambda x: self.lib.ptr(x.data_ptr())
        self.ctx.call_dev("bsx_prove_subchain_batch_dev", torch.cuda.current_stream().cuda_stream, self.lib.u32(B), self.lib.u32(n_jobs),
                          P(t["dh_leaf"]), P(t["dh_aunts"]), P(t["lb_leaf"]), P(t["lb_aunts"]), P(t["start_headers"]),
                          P(t["end_headers"]), P(t["batch_start"]), P(t["batch_end"]), P(t["global_end"]),
                          P(t["global_end_header"]), P(digests), P(subchains))

    def map_p2p(self, B, n_jobs, t, digests, peer_ptrs, rank, per, total_jobs, ranges_per_owner):
        """The map stage with every subchain record stored into its reducing rank's array (peer memory)."""
        P = lambda x: self.lib.ptr(x.data_ptr())
        bases = np.asarray(peer_ptrs, np.uint64)
        u32 = self.lib.u32
        self.ctx.call_dev("bsx_prove_subchain_batch_p2p_dev", torch.cuda.current_stream().cuda_stream, u32(B), u32(n_jobs),
                          P(t["dh_leaf"]), P(t["dh_aunts"]), P(t["lb_leaf"]), P(t["lb_aunts"]), P(t["start_headers"]),
                          P(t["end_headers"]), P(t["batch_start"]), P(t["batch_end"]), P(t["global_end"]),
                          P(t["global_end_header"]), P(digests), self.lib.ptr(bases.ctypes.data), u32(len(bases)), u32(rank),
                          u32(per), u32(total_jobs), u32(ranges_per_owner))

    def reduce(self, n_ranges, n_jobs, B, subchains, t, reduce_digests, reduce_nodes, dcs, fail):
        P = lambda x: self.lib.ptr(x.data_ptr())
        self.ctx.call_dev("bsx_reduce_subchains_dev", torch.cuda.current_stream().cuda_stream, self.lib.u32(n_ranges), self.lib.u32(n_jobs),
                          P(subchains), P(t["start_blocks"]), P(t["start_header"]), P(t["end_blocks"]), P(t["end_header"]),
                          self.lib.u32(B), P(reduce_digests), P(reduce_nodes), P(dcs), P(fail))


class ShardedHeaderRange:
    """One step = all map jobs of `n_ranges` ranges (this rank's slice), all-gather, reduce of this rank's ranges."""

    def __init__(self, backend, n_ranges: int, n_jobs: int, batch: int, rank: int = 0, world: int = 1, group=None):
        if n_ranges % world:
            raise ValueError("the number of ranges in flight must be divisible by the world size")
        self.be, self.R, self.J, self.B, self.rank, self.world, self.group = backend, n_ranges, n_jobs, batch, rank, world, group
        self.per = n_jobs // world
        self.own = slice(rank * (n_ranges // world), (rank + 1) * (n_ranges // world))
        R, per, B = n_ranges, self.per, batch
        e = backend.empty
        self.map_digests = e(R * per * (20 * B - 1) * 32)
        self.local_sub = e(R * per * SUBCHAIN_BYTES)
        Ro = R // world
        self.shard = None        # bsx_shard (CUDA backend, world > 1): peer stores + in-kernel flags
        self._symm = None
        self.exchange = "none" if world == 1 else "all_gather"
        if world > 1 and hasattr(backend, "shard") and os.environ.get("BSX_EXCHANGE", "p2p") == "p2p":
            self.shard = self._setup_shard()
            if self.shard:
                self.map_subchains = e(Ro * n_jobs * SUBCHAIN_BYTES)
        self.gathered = e(world * R * per * SUBCHAIN_BYTES) if world > 1 and not self.shard else None
        self.all_sub = (e(R * n_jobs * SUBCHAIN_BYTES) if not self.shard else None) if world > 1 else self.local_sub
        self.reduce_digests = e(Ro * max(n_jobs - 1, 1) * 32)
        self.reduce_nodes = e(Ro * max(n_jobs - 1, 1) * SUBCHAIN_BYTES)
        self.data_commitments = e(Ro * 32)
        self.fail = e(Ro * 4)
        self.t: Optional[dict] = None
        self._sio = None

    def _all_ok(self, ok: bool) -> bool:
        """True only if every rank succeeded (all ranks must take the same exchange path)."""
        flags = [None] * self.world
        dist.all_gather_object(flags, bool(ok), group=self.group)
        return all(flags)

    def _setup_shard(self):
        """bsx_shard on every rank + the peers' exchange buffers: CUDA IPC handles first, torch symmetric memory second."""
        import warnings
        R, J, B, W = self.R, self.J, self.B, self.world
        mode = os.environ.get("BSX_SHARD_MAP", "ipc")        # "ipc" | "symm": how the peers' buffers are mapped
        if mode == "ipc":
            sh, err = None, None
            try:
                sh = self.be.shard(self.rank, W, R, J, B)
                handles = [None] * W
                dist.all_gather_object(handles, sh.ipc_handle(), group=self.group)
                for w in range(W):
                    if w != self.rank:
                        sh.open_peer(w, handles[w])
            except Exception as exc:
                err = exc
            if self._all_ok(err is None):
                self.exchange = "peer stores + in-kernel flags (CUDA IPC)"
                dist.barrier(group=self.group)
                return sh
            if sh:
                sh.close()
            warnings.warn(f"CUDA IPC mapping of the exchange buffers failed on some rank ({err}); trying symmetric memory")
        sh, err = None, None
        try:
            import torch.distributed._symmetric_memory as symm
            nbytes = self.be.lib.Shard.exchange_bytes(W, R, J)
            buf = symm.empty(nbytes, dtype=torch.uint8, device=self.be.device)
            hdl = symm.rendezvous(buf, self.group or dist.group.WORLD)
            sh = self.be.shard(self.rank, W, R, J, B, exchange_buf=buf.data_ptr())
            for w in range(W):
                if w != self.rank:
                    sh.set_peer(w, int(hdl.buffer_ptrs[w]))
            self._symm = (buf, hdl)
        except Exception as exc:
            err = exc
        if self._all_ok(err is None):
            self.exchange = "peer stores + in-kernel flags (symmetric memory)"
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            return sh
        warnings.warn(f"no peer mapping of the exchange buffers ({type(err).__name__}: {err}); using all_gather_into_tensor")
        return None

    def load(self, host: Dict[str, np.ndarray]):
        """host arrays of ALL ranges -> device-resident shard + this rank's public inputs for the reduce."""
        sh = shard_map_inputs(host, self.R, self.J, self.B, self.rank, self.world)
        t = {k: self.be.tensor(v) for k, v in sh.items()}
        o = self.own
        t["start_blocks"] = self.be.tensor(np.ascontiguousarray(host["start_blocks"], np.uint64)[o])
        t["end_blocks"] = self.be.tensor(np.ascontiguousarray(host["end_blocks"], np.uint64)[o])
        t["start_header"] = self.be.tensor(np.ascontiguousarray(host["start_header"], np.uint8).reshape(self.R, 32)[o])
        t["end_header"] = self.be.tensor(np.ascontiguousarray(host["end_header"], np.uint8).reshape(self.R, 32)[o])
        self.t = t

    def step(self):
        R, J, per, W = self.R, self.J, self.per, self.world
        if self.shard:
            if self._sio is None:
                lib_ = self.be.lib
                t = self.t
                sin = lib_.fill_struct(lib_.ShardIn(), **{k: t[k].data_ptr() for k in (
                    "dh_leaf", "dh_aunts", "lb_leaf", "lb_aunts", "start_headers", "end_headers", "batch_start", "batch_end",
                    "global_end", "global_end_header", "start_blocks", "end_blocks", "start_header", "end_header")})
                sout = lib_.fill_struct(lib_.ShardOut(), map_digests=self.map_digests.data_ptr(), map_subchains=self.map_subchains.data_ptr(),
                                        reduce_digests=self.reduce_digests.data_ptr(), reduce_nodes=self.reduce_nodes.data_ptr(),
                                        data_commitments=self.data_commitments.data_ptr(), fail=self.fail.data_ptr())
                self._sio = (sin, sout)
            self.shard.step_dev(torch.cuda.current_stream().cuda_stream, *self._sio)
            return
        self.be.map(self.B, R * per, self.t, self.map_digests, self.local_sub)
        if W > 1:
            dist.all_gather_into_tensor(self.gathered, self.local_sub, group=self.group)
            # [W, R, per, 128] -> [R, W*per, 128]: job j of range r lives on rank j // per
            g = self.gathered.view(W, R, per * SUBCHAIN_BYTES).permute(1, 0, 2)
            self.all_sub.view(R, W, per * SUBCHAIN_BYTES).copy_(g)
        o = self.own
        sub = self.all_sub.view(R, J * SUBCHAIN_BYTES)[o].reshape(-1)
        self.be.reduce(R // W, J, self.B, sub, self.t, self.reduce_digests, self.reduce_nodes, self.data_commitments, self.fail)

    def exchange_text(self) -> str:
        return {"all_gather": "one all_gather_into_tensor", "none": "nothing (one rank)"}.get(
            self.exchange, f"{self.exchange}: bsx_shard_step_dev, no collective and no barrier launch")

    def results(self):
        Ro = self.R // self.world
        return dict(data_commitments=self.data_commitments.cpu().numpy().reshape(Ro, 32),
                    fail=self.fail.cpu().numpy().view(np.uint32).reshape(Ro),
                    reduce_nodes=self.reduce_nodes.cpu().numpy().reshape(Ro, max(self.J - 1, 1), SUBCHAIN_BYTES),
                    map_subchains=(self.map_subchains.cpu().numpy().reshape(Ro, self.J, SUBCHAIN_BYTES) if self.shard else
                                   self.all_sub.cpu().numpy().reshape(self.R, self.J, SUBCHAIN_BYTES)[self.own]),
                    local_map_digests=self.map_digests.cpu().numpy().reshape(self.R, self.per, 20 * self.B - 1, 32))

"""world_size-2 (and 4) gloo test of the sharded header_range host logic (blobstreamx_b200/distributed.py):
job partitioning, the single all-gather of subchain records, record re-ordering and the per-rank reduce.
The compute backend here is an ORACLE stand-in defined in this test file (CPU); on GPUs the same class runs with
CudaBackend over NCCL (tests/test_gpu_distributed.py, bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class OracleBackend:
    """tests-only: the map / reduce stages on the CPU oracle, same interface as distributed.CudaBackend."""

    def tensor(self, a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy())

    def empty(self, nbytes):
        return torch.zeros(nbytes, dtype=torch.uint8)

    def map(self, B, n_jobs, t, digests, subchains):
        from oracle import cbind as orc
        n = lambda k, dt=np.uint8: t[k].numpy().view(dt)
        dhl, dha = n("dh_leaf").reshape(n_jobs, B * 34), n("dh_aunts").reshape(n_jobs, B * 128)
        lbl, lba = n("lb_leaf").reshape(n_jobs, B * 72), n("lb_aunts").reshape(n_jobs, B * 128)
        sh, eh, geh = n("start_headers").reshape(n_jobs, 32), n("end_headers").reshape(n_jobs, 32), n("global_end_header").reshape(n_jobs, 32)
        bs, be, ge = n("batch_start", np.uint64), n("batch_end", np.uint64), n("global_end", np.uint64)
        dg = digests.numpy().reshape(n_jobs, 20 * B - 1, 32)
        sb = subchains.numpy().reshape(n_jobs, 128)
        for j in range(n_jobs):
            dg[j], sb[j] = orc.prove_subchain(B, dhl[j], dha[j], lbl[j], lba[j], sh[j], eh[j], int(bs[j]), int(be[j]), int(ge[j]), geh[j])

    def reduce(self, n_ranges, n_jobs, B, subchains, t, reduce_digests, reduce_nodes, dcs, fail):
        from oracle import cbind as orc
        sub = subchains.numpy().reshape(n_ranges, n_jobs, 128)
        sb, eb = t["start_blocks"].numpy().view(np.uint64), t["end_blocks"].numpy().view(np.uint64)
        shd, ehd = t["start_header"].numpy().reshape(n_ranges, 32), t["end_header"].numpy().reshape(n_ranges, 32)
        for r in range(n_ranges):
            w = orc.reduce_subchains(n_jobs, B, sub[r], int(sb[r]), shd[r], int(eb[r]), ehd[r])
            reduce_nodes.numpy().reshape(n_ranges, max(n_jobs - 1, 1), 128)[r] = w["reduce_nodes"]
            reduce_digests.numpy().reshape(n_ranges, max(n_jobs - 1, 1), 32)[r] = w["reduce_digests"]
            dcs.numpy().reshape(n_ranges, 32)[r] = np.frombuffer(w["data_commitment"], np.uint8)
            fail.numpy().view(np.uint32)[r] = w["fail"]


def _workload(R, J, B):
    import bench
    from blobstreamx_b200 import synthetic as S
    vs = S.ValidatorSet.make(S.SEED, n=4)
    ms = [S.header_range_inputs(J, B, nb, start=1_000_000 + 1000 * r, seed=S.SEED + r, valset=vs, with_skip=False)[0]
          for r, nb in zip(range(R), [None, J * B - 3, 5, 1][:R] + [None] * R)]
    return ms, bench.tile_ranges(ms, R)


def _worker(rank, world, port, R, J, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from blobstreamx_b200.distributed import ShardedHeaderRange
        ms, host = _workload(R, J, B)
        eng = ShardedHeaderRange(OracleBackend(), R, J, B, rank, world)
        eng.load(host)
        eng.step()
        eng.step()   # idempotent: buffers are reused
        res = eng.results()
        q.put((rank, res["data_commitments"], res["fail"], res["reduce_nodes"], res["map_subchains"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,R,J,B", [(2, 4, 4, 8), (4, 4, 8, 4), (2, 2, 2, 16), (8, 8, 8, 2)])
def test_sharded_header_range_gloo(world, R, J, B):
    from oracle import cbind as orc
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, R, J, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, dcs, fail, nodes, subs = q.get(timeout=180)
        got[rank] = (dcs, fail, nodes, subs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ms, _ = _workload(R, J, B)
    per = R // world
    for rank in range(world):
        dcs, fail, nodes, subs = got[rank]
        for k in range(per):
            m = ms[rank * per + k]
            want = orc.prove_data_commitment(J, B, m.dh_leaf, m.dh_aunts, m.lb_leaf, m.lb_aunts, m.start_headers, m.end_headers,
                                             m.start_block, m.start_header, m.end_block, m.end_header)
            assert want["fail"] == 0 and fail[k] == 0
            assert dcs[k].tobytes() == want["data_commitment"]
            assert (nodes[k] == want["reduce_nodes"]).all()
            assert (subs[k] == want["map_subchains"]).all()


def test_shard_map_inputs_partition():
    """Every job of every range lands on exactly one rank, in order."""
    from blobstreamx_b200.distributed import job_slice, shard_map_inputs
    R, J, B, W = 3, 8, 4, 4
    _, host = _workload(R, J, B)
    seen = np.zeros((R, J), int)
    for r in range(W):
        js = job_slice(J, r, W)
        sh = shard_map_inputs(host, R, J, B, r, W)
        seen[:, js] += 1
        full = np.asarray(host["dh_leaf"]).reshape(R, J, B, 34)
        assert (sh["dh_leaf"].reshape(R, js.stop - js.start, B, 34) == full[:, js]).all()
        bs = sh["batch_start"].reshape(R, -1)
        assert (bs[:, 0] == np.asarray(host["start_blocks"]) + np.uint64(js.start * B)).all()
    assert (seen == 1).all()
    with pytest.raises(ValueError):
        job_slice(6, 0, 4)

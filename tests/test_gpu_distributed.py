"""ShardedHeaderRange with the CUDA backend on one GPU (world 1) and, when the box has 2+ GPUs, over NCCL with two
ranks -- results equal the oracle's.  The N>1 host logic is also covered on CPU by tests/test_distributed_gloo.py."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_world1_cuda():
    import bench
    from blobstreamx_b200 import synthetic as S
    from blobstreamx_b200.distributed import CudaBackend, ShardedHeaderRange
    from oracle import cbind as orc
    R, J, B = 3, 8, 16
    vs = S.ValidatorSet.make(S.SEED, n=4)
    ms = [S.header_range_inputs(J, B, nb, start=5_000_000 + 1000 * r, seed=S.SEED + r, valset=vs, with_skip=False)[0]
          for r, nb in enumerate((None, 100, 3))]
    eng = ShardedHeaderRange(CudaBackend(0), R, J, B)
    eng.load(bench.tile_ranges(ms, R))
    eng.step()
    res = eng.results()
    for r, m in enumerate(ms):
        w = orc.prove_data_commitment(J, B, m.dh_leaf, m.dh_aunts, m.lb_leaf, m.lb_aunts, m.start_headers, m.end_headers,
                                      m.start_block, m.start_header, m.end_block, m.end_header, threads=4)
        assert res["fail"][r] == 0 == w["fail"]
        assert res["data_commitments"][r].tobytes() == w["data_commitment"]
        assert (res["reduce_nodes"][r] == w["reduce_nodes"]).all() and (res["local_map_digests"][r] == w["map_digests"]).all()


def test_bench_two_ranks_nccl():
    """bench.py under torchrun with 2 ranks (its correctness gate compares every rank's outputs with the oracle)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "3",
                          "--warmup", "3", "--ranges", "16", "--e2e-ranges", "4", "--distinct", "2"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    assert '"n_gpus": 2' in out.stdout


def test_peer_store_map_emulated_two_ranks():
    """bsx_prove_subchain_batch_p2p_dev on ONE GPU: two "ranks" run one after the other, each storing its jobs' subchain
    records into the array of the rank that reduces the range (both arrays local here).  The gathered arrays must equal the
    records of the plain map stage, and the reduce over them the oracle's."""
    import bench
    import torch
    from blobstreamx_b200 import synthetic as S
    from blobstreamx_b200.distributed import SUBCHAIN_BYTES, CudaBackend, shard_map_inputs
    from oracle import cbind as orc
    R, J, B, W = 4, 8, 4, 2
    ms = [S.header_range_inputs(J, B, nb, start=6_000_000 + 1000 * r, seed=S.SEED + r, with_skip=False)[0]
          for r, nb in enumerate((None, 13, 5, 32))]
    host = bench.tile_ranges(ms, R)
    be = CudaBackend(0)
    per, Ro = J // W, R // W
    bufs = [be.empty(Ro * J * SUBCHAIN_BYTES) for _ in range(W)]
    ptrs = [int(b.data_ptr()) for b in bufs]
    for rank in range(W):
        t = {k: be.tensor(v) for k, v in shard_map_inputs(host, R, J, B, rank, W).items()}
        dig = be.empty(R * per * (20 * B - 1) * 32)
        be.map_p2p(B, R * per, t, dig, ptrs, rank, per, J, Ro)
        plain = be.empty(R * per * SUBCHAIN_BYTES)
        be.map(B, R * per, t, dig, plain)
        torch.cuda.synchronize()
        got = torch.cat(bufs).cpu().numpy().reshape(R, J, SUBCHAIN_BYTES)[:, rank * per:(rank + 1) * per]
        assert (got == plain.cpu().numpy().reshape(R, per, SUBCHAIN_BYTES)).all()
    gathered = torch.cat(bufs).cpu().numpy().reshape(R, J, SUBCHAIN_BYTES)
    for r, m in enumerate(ms):
        w = orc.prove_data_commitment(J, B, m.dh_leaf, m.dh_aunts, m.lb_leaf, m.lb_aunts, m.start_headers, m.end_headers,
                                      m.start_block, m.start_header, m.end_block, m.end_header)
        assert (gathered[r] == w["map_subchains"]).all()

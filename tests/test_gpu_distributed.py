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


def _shard_io(lib, be, host, R, J, B, rank, W):
    """device tensors + bsx_shard_in / bsx_shard_out of one rank"""
    from blobstreamx_b200.distributed import SUBCHAIN_BYTES, shard_map_inputs
    per, Ro = J // W, R // W
    own = slice(rank * Ro, (rank + 1) * Ro)
    t = {k: be.tensor(v) for k, v in shard_map_inputs(host, R, J, B, rank, W).items()}
    t["start_blocks"] = be.tensor(np.ascontiguousarray(host["start_blocks"], np.uint64)[own])
    t["end_blocks"] = be.tensor(np.ascontiguousarray(host["end_blocks"], np.uint64)[own])
    t["start_header"] = be.tensor(np.ascontiguousarray(host["start_header"], np.uint8).reshape(R, 32)[own])
    t["end_header"] = be.tensor(np.ascontiguousarray(host["end_header"], np.uint8).reshape(R, 32)[own])
    o = dict(map_digests=be.empty(R * per * (20 * B - 1) * 32), map_subchains=be.empty(Ro * J * SUBCHAIN_BYTES),
             reduce_digests=be.empty(Ro * (J - 1) * 32), reduce_nodes=be.empty(Ro * (J - 1) * SUBCHAIN_BYTES),
             data_commitments=be.empty(Ro * 32), fail=be.empty(Ro * 4))
    sin = lib.fill_struct(lib.ShardIn(), **{k: v.data_ptr() for k, v in t.items()})
    sout = lib.fill_struct(lib.ShardOut(), **{k: v.data_ptr() for k, v in o.items()})
    return t, o, sin, sout


def test_shard_c_abi_two_ranks_on_one_gpu():
    """The bsx_shard_* C ABI with two ranks emulated on ONE GPU (two ctxs, two streams, peers set by pointer): the map kernel
    of each rank stores its records into the reducing rank's array and publishes its step flag, the reduce kernels acquire
    both flags.  Five steps (both parities of the alternating arrays, reuse after two steps); every step's outputs equal
    the oracle's, and the gathered records equal the plain map stage's."""
    import bench
    import torch
    from blobstreamx_b200 import lib, synthetic as S
    from blobstreamx_b200.distributed import SUBCHAIN_BYTES, CudaBackend
    from oracle import cbind as orc
    R, J, B, W = 4, 8, 4, 2
    ms = [S.header_range_inputs(J, B, nb, start=6_000_000 + 1000 * r, seed=S.SEED + r, with_skip=False)[0]
          for r, nb in enumerate((None, 13, 5, 32))]
    host = bench.tile_ranges(ms, R)
    bes = [CudaBackend(0) for _ in range(W)]
    shards = [lib.Shard(bes[r].ctx, r, W, R, J, B) for r in range(W)]
    for r in range(W):
        for w in range(W):
            if w != r:
                shards[r].set_peer(w, shards[w].exchange_buffer()[0])
    io = [_shard_io(lib, bes[r], host, R, J, B, r, W) for r in range(W)]
    streams = [torch.cuda.Stream() for _ in range(W)]
    torch.cuda.synchronize()
    want = [orc.prove_data_commitment(J, B, m.dh_leaf, m.dh_aunts, m.lb_leaf, m.lb_aunts, m.start_headers, m.end_headers,
                                      m.start_block, m.start_header, m.end_block, m.end_header) for m in ms]
    per, Ro = J // W, R // W
    for step in range(5):
        for r in (range(W) if step % 2 == 0 else reversed(range(W))):      # either rank may be first
            shards[r].step_dev(streams[r].cuda_stream, io[r][2], io[r][3])
        torch.cuda.synchronize()
        for r in range(W):
            o = io[r][1]
            assert not o["fail"].cpu().numpy().view(np.uint32).any(), (step, r)
            sub = o["map_subchains"].cpu().numpy().reshape(Ro, J, SUBCHAIN_BYTES)
            dig = o["map_digests"].cpu().numpy().reshape(R, per, 20 * B - 1, 32)
            for k in range(Ro):
                w = want[r * Ro + k]
                assert (sub[k] == w["map_subchains"]).all(), (step, r, k)
                assert o["data_commitments"].cpu().numpy().reshape(Ro, 32)[k].tobytes() == w["data_commitment"]
                assert (o["reduce_nodes"].cpu().numpy().reshape(Ro, J - 1, SUBCHAIN_BYTES)[k] == w["reduce_nodes"]).all()
            for g in range(R):
                assert (dig[g] == want[g]["map_digests"][r * per:(r + 1) * per]).all()
            # poison the outputs so that the next step must rewrite them
            for v in o.values():
                v.fill_(0xEE)
        torch.cuda.synchronize()
    for s in shards:
        s.close()


def test_shard_missing_peer_times_out_instead_of_hanging():
    """A rank whose peer never runs its step gets BSX_FAIL_EXCHANGE_TIMEOUT (2048) after 4 s -- never a hung GPU."""
    import time
    import bench
    import torch
    from blobstreamx_b200 import lib, synthetic as S
    from blobstreamx_b200.distributed import CudaBackend
    R, J, B, W = 2, 4, 2, 2
    ms = [S.header_range_inputs(J, B, None, start=7_000_000 + 1000 * r, seed=S.SEED + r, with_skip=False)[0] for r in range(R)]
    host = bench.tile_ranges(ms, R)
    bes = [CudaBackend(0) for _ in range(W)]
    shards = [lib.Shard(bes[r].ctx, r, W, R, J, B) for r in range(W)]
    shards[0].set_peer(1, shards[1].exchange_buffer()[0])
    t, o, sin, sout = _shard_io(lib, bes[0], host, R, J, B, 0, W)
    t0 = time.time()
    shards[0].step_dev(torch.cuda.current_stream().cuda_stream, sin, sout)
    torch.cuda.synchronize()
    assert 3.0 < time.time() - t0 < 20.0
    assert (o["fail"].cpu().numpy().view(np.uint32) & 2048).all()
    with pytest.raises(lib.BsxError):        # rank 1 has no mapping of rank 0's buffer yet
        shards[1].step_dev(torch.cuda.current_stream().cuda_stream, sin, sout)
    for s in shards:
        s.close()

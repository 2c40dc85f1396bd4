#!/bin/bash
# r02c (2 GPUs): bsx_shard_* C ABI -- emulated two-rank test, real two-rank bench over CUDA IPC peer stores + in-kernel flags
OUT=gpurun_out/r02c
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== pytest distributed"; timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_ed25519_builds.py -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest.log
echo "== bench N=1"; timeout 600 python bench.py --no-cpu 2> $OUT/bench1.err | tee $OUT/bench_n1.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value']/1e6, 'ms', d['ms_per_step'], d['step_ms'], 'e2e', d['e2e']['value']/1e6, '2048:', d['header_range_2048']['value']/1e6, d['header_range_2048']['ms_per_step'])"
for ex in p2p nccl; do
echo "== bench N=2 exchange=$ex"; BSX_EXCHANGE=$ex timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no-cpu 2> $OUT/bench2_$ex.err | tee $OUT/bench_n2_$ex.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value']/1e6, 'ms', d['ms_per_step'], d['step_ms'], d['config']['sharding'], 'e2e', d['e2e']['value']/1e6, '2048:', d['header_range_2048']['value']/1e6, d['header_range_2048']['ms_per_step'], d['header_range_2048']['step_ms'])"
tail -3 $OUT/bench2_$ex.err
done
echo "== bench N=2 symm"; BSX_SHARD_MAP=symm timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --no-cpu --no-2048 --steps 20 2> $OUT/bench2_symm.err | tee $OUT/bench_n2_symm.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value']/1e6, 'ms', d['ms_per_step'], d['step_ms'], d['config']['sharding'])"
tail -3 $OUT/bench2_symm.err

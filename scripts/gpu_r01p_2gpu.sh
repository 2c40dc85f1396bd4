#!/bin/bash
# r01p: 2-GPU check of the sharded path with the signed-window Ed25519 kernel and the wave-fill arrangement
OUT=gpurun_out/r01p_2gpu
mkdir -p $OUT
nvidia-smi -L | tee $OUT/gpus.txt
echo "== pytest distributed"; timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -3 | tee $OUT/pytest_gpu.log
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu 2> $OUT/bench2.err | tee $OUT/bench_n2.json | cut -c1-260
tail -3 $OUT/bench2.err
echo "== bench N=1"; timeout 600 python bench.py --no-cpu 2> $OUT/bench1.err | tee $OUT/bench_n1.json | cut -c1-200

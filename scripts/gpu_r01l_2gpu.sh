#!/bin/bash
# r01l: 2-GPU check of the sharded path after the Ed25519 / pipeline / carveout changes
OUT=gpurun_out/r01l
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
nvidia-smi -L | tee $OUT/gpus.txt
echo "== pytest distributed"; timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu.log
echo "== bench N=1"; timeout 600 python bench.py --no-cpu 2> $OUT/bench1.err | tee $OUT/bench_n1.json | cut -c1-200
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 2> $OUT/bench2.err | tee $OUT/bench_n2.json | cut -c1-200
tail -3 $OUT/bench2.err
echo "== reference arm N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>> $OUT/bench2.err | tee $OUT/bench_ref_n2.json | cut -c1-200

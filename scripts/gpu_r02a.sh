#!/bin/bash
# r02a: tunables in the ctx, every Ed25519 build under pytest, bench with the 2048 block / step distribution / e2e oracle gate
OUT=gpurun_out/r02a
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-300; tail -5 $OUT/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> $OUT/bench.err | tee $OUT/bench_ref.json | cut -c1-200

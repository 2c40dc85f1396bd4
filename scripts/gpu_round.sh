#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, ncu launch list, ncu full capture of the top kernel.
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh [tag] [kernel-regex]'
TAG=${1:-r01}
KRE=${2:-prove_subchain_kernel}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== bench"; timeout 600 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> $OUT/bench.err | tee $OUT/bench_ref.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --ranges 64 --e2e-ranges 16 --no-cpu --no-check > $OUT/ncu_bench.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 4 -c 2 -f -o $OUT/prof \
    python bench.py --steps 3 --warmup 3 --ranges 64 --e2e-ranges 16 --no-cpu --no-check > $OUT/ncu_full.log 2>&1
ls -la $OUT

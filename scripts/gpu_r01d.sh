#!/bin/bash
OUT=gpurun_out/r01d
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json; tail -5 $OUT/bench.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --ranges 64 --e2e-ranges 16 --no-cpu --no-check > $OUT/ncu_bench.log 2>&1
tail -3 $OUT/ncu_bench.log

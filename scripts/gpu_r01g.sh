#!/bin/bash
OUT=gpurun_out/r01g
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
echo "== gates"; timeout 300 python bench.py --mode gates --steps 10 --warmup 3 2>> $OUT/bench.err | tee $OUT/gates.json | cut -c1-500
echo "== ncu gates"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gl_gate_eval_kernel -s 3 -c 1 -f -o $OUT/prof_gates \
    python bench.py --mode gates --steps 2 --warmup 3 --no-cpu --no-check > $OUT/ncu_gates.log 2>&1
tail -3 $OUT/bench.err

#!/bin/bash
# timeline of the device-resident step (HR_TRACE) under different Ed25519 register budgets
OUT=gpurun_out/${1:-hr_trace2}
mkdir -p $OUT
run() { local r=$1; shift; echo "== ranges=$r $*"; env "$@" BSX_HR_TRACE=1 timeout 300 python bench.py --ranges $r --no-cpu --no-check --e2e-threads 1 --e2e-ranges 8 --no-2048 --steps 3 --warmup 3 2>&1 >/dev/null | grep "bsx header_range" | tail -3; }
run 378 BSX_X=0
run 378 BSX_ED_OCC=8
run 378 BSX_ED_OCC=8 BSX_HR_HASH_STREAM=0
run 757 BSX_X=0
run 757 BSX_ED_OCC=8 BSX_ED_RESIDENT=4

#!/bin/bash
# timeline of the pipelined host call (BSX_PIPE_TRACE) under different Ed25519 register budgets / skip-half orders
OUT=gpurun_out/${1:-pipe_trace}
mkdir -p $OUT
for cfg in "occ4 A=1" "occ6 BSX_ED_OCC=6" "occ8 BSX_ED_OCC=8" "occ4_last BSX_PIPE_ED=2" "occ6_last BSX_ED_OCC=6 BSX_PIPE_ED=2"; do
  set -- $cfg; tag=$1; shift
  echo "== $tag"
  env "$@" BSX_PIPE_TRACE=1 timeout 300 python bench.py --no-cpu --no-check --steps 4 --warmup 3 --e2e-threads 1 2> $OUT/trace_$tag.log | tee $OUT/bench_$tag.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('single_call'))"
  grep "bsx pipe" $OUT/trace_$tag.log | tail -12
done

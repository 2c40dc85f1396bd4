#!/bin/bash
# when do the two halves of the device-resident header_range step finish (BSX_HR_TRACE), by Ed25519 register budget
OUT=gpurun_out/${1:-hr_trace}
mkdir -p $OUT
for cfg in "occ4 A=1" "occ6 BSX_ED_OCC=6" "occ8 BSX_ED_OCC=8" "occ4_proofs6 BSX_PROOFS_OCC=6" "occ8_proofs6 BSX_ED_OCC=8 BSX_PROOFS_OCC=6"; do
  set -- $cfg; tag=$1; shift
  echo "== $tag"
  env "$@" BSX_HR_TRACE=1 timeout 300 python bench.py --no-cpu --no-check --steps 4 --warmup 3 --e2e-threads 1 2> $OUT/trace_$tag.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
  grep "bsx header_range" $OUT/trace_$tag.log | tail -3
done

#!/bin/bash
# same-box A/B of the r01p step changes at other batch sizes
OUT=gpurun_out/${1:-ab_r256}
mkdir -p $OUT
for r in ${RANGES:-256 512}; do
for cfg in "default A=1" "uncapped BSX_ED_REGS=0" "hash_main BSX_HR_HASH_STREAM=0" "old BSX_ED_REGS=0 BSX_HR_HASH_STREAM=0" "default_b A=1"; do
  set -- $cfg; tag=$1; shift
  echo "== ranges=$r $tag: $(env "$@" timeout 300 python bench.py --ranges $r --no-cpu --no-check --e2e-threads 1 --steps 20 --warmup 5 2>> $OUT/err.log | tee $OUT/bench_${tag}_r$r.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])")"
done
done

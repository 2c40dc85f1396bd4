#!/bin/bash
# same-box A/B: compact vs inlined point arithmetic in the 192-register Ed25519 build beside the hash kernels
OUT=gpurun_out/${1:-ab_inl192}
mkdir -p $OUT
for r in 378 756; do
for cfg in "compact A=1" "inlined BSX_ED_INLINE=1" "compact_b A=1" "inlined_b BSX_ED_INLINE=1"; do
  set -- $cfg; tag=$1; shift
  echo "== ranges=$r $tag: $(env "$@" timeout 300 python bench.py --ranges $r --no-cpu --no-check --e2e-threads 1 --steps 20 --warmup 5 2>> $OUT/err.log | tee $OUT/bench_${tag}_r$r.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernels_alone_ms'])")"
done
done

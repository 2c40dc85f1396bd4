#!/bin/bash
# end-to-end host path: Ed25519 128-register build (default for the pipelined call) vs BSX_ED_OCC=4
OUT=gpurun_out/${1:-e2e_occ}
mkdir -p $OUT
hr() { tag=$1; shift; envs=(); args=()
  for x in "$@"; do case $x in *=*) envs+=("$x");; *) args+=("$x");; esac; done
  echo "== $tag"; env "${envs[@]}" timeout 300 python bench.py --no-cpu "${args[@]}" 2>> $OUT/err.log | tee $OUT/bench_$tag.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['single_call'])"
}
for rep in 1 2 3; do
  for r in 378 512; do hr corun_r${r}_$rep A=1 --ranges $r; hr occ4_r${r}_$rep BSX_ED_OCC=4 --ranges $r; done
done
tail -3 $OUT/err.log

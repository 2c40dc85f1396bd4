#!/bin/bash
# r02g: FP64 Ed25519 register budgets at batch sizes that fill their own waves (occ 6 -> 568 ranges, occ 8 -> 757 ranges)
OUT=gpurun_out/r02g
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
run() {  # ranges env...
  local r=$1; shift
  echo "== ranges=$r $*"
  env "$@" timeout 600 python bench.py --no-cpu --no-2048 --steps 20 --warmup 5 --ranges $r --e2e-ranges 64 2>> $OUT/bench.err | tee $OUT/bench_r${r}_$(echo "$*" | tr ' =' '__').json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value', round(d['value']/1e6,2), 'M headers/s, ms', round(d['ms_per_step'],3), 'alone', {k[:14]: round(v,3) for k,v in d['kernels_alone_ms'].items()})"
}
run 378 BSX_X=0
run 568 BSX_X=0
run 568 BSX_ED_OCC=6
run 757 BSX_X=0
run 757 BSX_ED_OCC=8
run 757 BSX_ED_OCC=6
for n in 56832 75776; do
for occ in 0 6 8; do
  echo "== ed25519 n=$n occ=$occ"
  BSX_ED_OCC=$occ BSX_ED_INLINE=0 timeout 300 python bench.py --mode ed25519 --sigs $n --steps 10 --warmup 3 --no-cpu 2>> $OUT/bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), 'M sig/s', round(d['ms_per_step'],3), 'ms')"
done
done
tail -3 $OUT/bench.err

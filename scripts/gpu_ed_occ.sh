#!/bin/bash
# Ed25519 register budgets (2 / 3 / 4 warps per SM sub-partition) alone and inside the header_range step
OUT=gpurun_out/${1:-ed_occ}
mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_ed25519.py tests/test_gpu_verify.py -m gpu -x -q 2>&1 | tail -2
run() { tag=$1; shift
  for n in 37888 100000 400000; do
    echo "== $tag ed n=$n"; env "$@" timeout 300 python bench.py --mode ed25519 --sigs $n --steps 5 --warmup 3 --no-cpu 2>> $OUT/err.log | tee $OUT/ed_${tag}_$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
  done
}
hr() { tag=$1; shift; envs=(); args=()
  for x in "$@"; do case $x in *=*) envs+=("$x");; *) args+=("$x");; esac; done
  echo "== $tag header_range"; env "${envs[@]}" timeout 300 python bench.py --no-cpu "${args[@]}" 2>> $OUT/err.log | tee $OUT/bench_$tag.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d.get('kernels_alone_ms'), d['e2e']['value'], d['e2e']['single_call'])"
}
run occ4inl A=1
run occ4 BSX_ED_INLINE=0
run occ6 BSX_ED_OCC=6
run occ8 BSX_ED_OCC=8
hr occ4 A=1
hr occ8 BSX_ED_OCC=8
hr occ4b A=1
hr occ8b BSX_ED_OCC=8
for r in 256 512 756; do hr occ4_r$r A=1 --ranges $r; hr occ8_r$r BSX_ED_OCC=8 --ranges $r; done
tail -3 $OUT/err.log

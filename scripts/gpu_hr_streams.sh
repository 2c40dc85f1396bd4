#!/bin/bash
# device-resident header_range step: skip hash kernel on its own stream, Ed25519 register caps (value first, then the BSX_HR_TRACE timeline)
OUT=gpurun_out/${1:-hr_streams}
mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_verify.py tests/test_gpu_ed25519.py -m gpu -x -q 2>&1 | tail -2
for cfg in "old_order BSX_HR_HASH_STREAM=0" "side A=1" "side_r192 BSX_ED_REGS=192" "side_r176 BSX_ED_REGS=176" "side_r160 BSX_ED_REGS=160" "side_occ6 BSX_ED_OCC=6" "side_occ8 BSX_ED_OCC=8" "side_b A=1" "side_r192_b BSX_ED_REGS=192"; do
  set -- $cfg; tag=$1; shift
  echo "== $tag"
  env "$@" timeout 300 python bench.py --no-cpu --no-check --steps 20 --warmup 5 --e2e-threads 1 2> $OUT/err_$tag.log | tee $OUT/bench_$tag.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
  env "$@" BSX_HR_TRACE=1 timeout 300 python bench.py --no-cpu --no-check --steps 3 --warmup 3 --e2e-threads 1 2>&1 >/dev/null | grep "bsx header_range" | tail -1
done

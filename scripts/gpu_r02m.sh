#!/bin/bash
# r02m: key-table window width A/B (this build) -- Ed25519 tests, step at 378 / 757 ranges, Ed25519 alone
OUT=gpurun_out/${1:-r02m}
mkdir -p $OUT
echo "== pytest ed25519"; timeout 1500 python -m pytest tests/test_gpu_ed25519_builds.py tests/test_gpu_ed25519.py -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_ed.log
run() { local r=$1; shift
  env "$@" timeout 300 python bench.py --ranges $r --no-cpu --e2e-threads 1 --e2e-ranges 64 --no-2048 --steps 20 --warmup 5 2>> $OUT/err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ranges=$r $*', round(d['value']/1e6,1), 'M headers/s', round(d['ms_per_step'],3), 'ms', {k[:14]: round(v,3) for k,v in d['kernels_alone_ms'].items()}, 'general', round(d['general_path']['value']/1e6,1))"
}
run 757 BSX_X=0
run 757 BSX_ED_KOCC=4
run 757 BSX_ED_KOCC=6
run 757 BSX_ED_KOCC=4 BSX_ED_INLINE=1
run 1514 BSX_X=0
run 1514 BSX_ED_KOCC=4
run 378 BSX_X=0
run 1135 BSX_X=0
for n in 37888 75776; do
  timeout 300 python bench.py --mode ed25519 --sigs $n --steps 5 --warmup 3 --no-cpu 2>> $OUT/err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ed25519 n=$n', round(d['value']/1e6,2), 'Msig/s', round(d['ms_per_step'],3), 'ms')"
done
tail -3 $OUT/err.log

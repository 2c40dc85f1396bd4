#!/bin/bash
# r02l: per-key window tables for h*A: GPU tests, step at 378 / 757 ranges per Ed25519 register budget, Ed25519 alone
OUT=gpurun_out/r02l
mkdir -p $OUT
echo "== pytest ed25519"; timeout 1500 python -m pytest tests/test_gpu_ed25519_builds.py tests/test_gpu_ed25519.py tests/test_gpu_verify.py -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_ed.log
run() { local r=$1; shift
  env "$@" timeout 300 python bench.py --ranges $r --no-cpu --e2e-threads 1 --e2e-ranges 64 --no-2048 --steps 20 --warmup 5 2>> $OUT/err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ranges=$r $*', round(d['value']/1e6,1), 'M headers/s', round(d['ms_per_step'],3), 'ms', {k[:14]: round(v,3) for k,v in d['kernels_alone_ms'].items()})"
}
run 757 BSX_ED_KEYTAB=0
run 757 BSX_X=0
run 757 BSX_ED_OCC=6
run 757 BSX_ED_OCC=4
run 378 BSX_X=0
run 378 BSX_ED_OCC=8
run 568 BSX_X=0
run 1135 BSX_X=0
run 1514 BSX_X=0
for n in 37888 75776 100000; do for occ in 0 6 8; do
  BSX_ED_OCC=$occ timeout 300 python bench.py --mode ed25519 --sigs $n --steps 5 --warmup 3 --no-cpu 2>> $OUT/err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ed25519 n=$n occ=$occ', round(d['value']/1e6,2), 'Msig/s', round(d['ms_per_step'],3), 'ms')"
done; done
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu.log
tail -3 $OUT/err.log

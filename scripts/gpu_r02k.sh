#!/bin/bash
# r02k: R-decompression shortcut + SHA-256 additions on the FMA pipe: GPU tests, step at 378 / 757 ranges, Ed25519 alone, trees, sweeps
OUT=gpurun_out/r02k
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_gpu.log
run() { local r=$1; shift
  env "$@" timeout 300 python bench.py --ranges $r --no-cpu --e2e-threads 1 --e2e-ranges 64 --no-2048 --steps 20 --warmup 5 2>> $OUT/err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ranges=$r $*', round(d['value']/1e6,1), 'M headers/s', round(d['ms_per_step'],3), 'ms', {k[:14]: round(v,3) for k,v in d['kernels_alone_ms'].items()})"
}
run 757 BSX_X=0
run 378 BSX_X=0
run 568 BSX_X=0
for n in 37888 75776 100000; do
  timeout 300 python bench.py --mode ed25519 --sigs $n --steps 5 --warmup 3 --no-cpu 2>> $OUT/err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ed25519 n=$n', round(d['value']/1e6,2), 'Msig/s', round(d['ms_per_step'],3), 'ms')"
done
timeout 300 python bench.py --mode tree --trees 4096 --steps 10 --warmup 3 --no-cpu 2>> $OUT/err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tree4096', round(d['ms_per_step'],3), 'ms')"
timeout 300 python bench.py --mode shape --steps 10 --warmup 3 --no-cpu 2>> $OUT/err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('shape', round(d['ms_per_step'],3), 'ms')"
python scripts/latency_single.py 2>> $OUT/err.log | tail -5
tail -3 $OUT/err.log

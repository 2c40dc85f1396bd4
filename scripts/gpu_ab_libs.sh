#!/bin/bash
# A/B of library builds under csrc/build/variants (BSX_LIB_PATH) against the default build: Ed25519 batches + the header_range step
OUT=gpurun_out/${1:-ab}
mkdir -p $OUT
for lib in default $(ls blobstreamx_b200/csrc/build/variants/*.so 2>/dev/null); do
  if [ $lib = default ]; then unset BSX_LIB_PATH; tag=default; else export BSX_LIB_PATH=$PWD/$lib; tag=$(basename $lib .so); fi
  for n in 37800 100000; do
    timeout 300 python bench.py --mode ed25519 --sigs $n --steps 5 --warmup 3 --no-cpu 2>> $OUT/err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag n=$n', round(d['value']/1e6,2), 'Msig/s', round(d['ms_per_step'],3), 'ms')"
  done
  timeout 300 python bench.py --no-cpu --e2e-threads 1 2>> $OUT/err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag header_range', round(d['value']/1e6,1), round(d['ms_per_step'],3), d['kernels_alone_ms'], round(d['e2e']['value']/1e6,1))"
done
tail -2 $OUT/err.log

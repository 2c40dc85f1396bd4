#!/bin/bash
# A/B of library builds under csrc/build/variants (BSX_LIB_PATH) on the SHA-256-bound paths: header_range step (map stage alone), 2048-leaf trees, input shaping
OUT=gpurun_out/${1:-ab_sha}
mkdir -p $OUT
for rep in 1 2; do
for lib in default $(ls blobstreamx_b200/csrc/build/variants/*.so 2>/dev/null); do
  if [ $lib = default ]; then unset BSX_LIB_PATH; tag=default; else export BSX_LIB_PATH=$PWD/$lib; tag=$(basename $lib .so); fi
  timeout 300 python bench.py --no-cpu --e2e-threads 1 --no-2048 --steps 20 --warmup 5 2>> $OUT/err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag header_range', round(d['value']/1e6,1), round(d['ms_per_step'],3), d['kernels_alone_ms'])"
  timeout 300 python bench.py --mode tree --trees 4096 --steps 10 --warmup 3 --no-cpu 2>> $OUT/err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag tree4096', round(d['ms_per_step'],3), 'ms')"
  timeout 300 python bench.py --mode shape --steps 10 --warmup 3 --no-cpu 2>> $OUT/err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag shape', round(d['ms_per_step'],3), 'ms')"
done
done 2>&1 | tee $OUT/ab.txt
tail -2 $OUT/err.log

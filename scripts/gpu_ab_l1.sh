#!/bin/bash
# is L1 capacity what the two halves of the step fight over?  carveout sweep + the SHA-256 build without the L1-resident tail table
OUT=gpurun_out/${1:-ab_l1}
mkdir -p $OUT
run() { local r=$1; shift
  env "$@" timeout 300 python bench.py --ranges $r --no-cpu --no-check --e2e-threads 1 --e2e-ranges 8 --no-2048 --steps 20 --warmup 5 2>> $OUT/err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ranges=$r $*', round(d['value']/1e6,1), 'M headers/s', round(d['ms_per_step'],3), 'ms', {k[:14]: round(v,3) for k,v in d['kernels_alone_ms'].items()})"
}
S1=BSX_LIB_PATH=$PWD/blobstreamx_b200/csrc/build/variants/libbsx_sha1.so
for c in 50 30 15; do
  run 378 BSX_CARVEOUT=$c
  run 378 BSX_CARVEOUT=$c BSX_ED_OCC=8
  run 757 BSX_CARVEOUT=$c
done
run 378 $S1
run 378 $S1 BSX_ED_OCC=8
run 757 $S1
run 378 $S1 BSX_CARVEOUT=15 BSX_ED_OCC=8

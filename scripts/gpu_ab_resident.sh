#!/bin/bash
# co-residency experiment: the 128-register Ed25519 build limited to k CTAs per SM (ED_RESIDENT) so that the SHA-256 kernels share the SMs with it
OUT=gpurun_out/${1:-ab_resident}
mkdir -p $OUT
run() { # ranges, env...
  local r=$1; shift
  env "$@" timeout 300 python bench.py --ranges $r --no-cpu --no-check --e2e-threads 1 --e2e-ranges 64 --no-2048 --steps 20 --warmup 5 2>> $OUT/err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ranges=$r $*', round(d['value']/1e6,1), 'M headers/s', round(d['ms_per_step'],3), 'ms', {k[:14]: round(v,3) for k,v in d['kernels_alone_ms'].items()})"
}
for lib in default blobstreamx_b200/csrc/build/variants/libbsx_fma13.so; do
  if [ $lib = default ]; then L=BSX_X=0; else L=BSX_LIB_PATH=$PWD/$lib; fi
  run 757 $L
  run 757 $L BSX_ED_OCC=8 BSX_ED_RESIDENT=4
  run 757 $L BSX_ED_OCC=8 BSX_ED_RESIDENT=6
  run 568 $L BSX_ED_OCC=8 BSX_ED_RESIDENT=6
  run 473 $L BSX_ED_OCC=8 BSX_ED_RESIDENT=5
  run 378 $L BSX_ED_OCC=8 BSX_ED_RESIDENT=4
  run 378 $L BSX_ED_OCC=8
  run 378 $L
done 2>&1 | tee $OUT/ab.txt
tail -3 $OUT/err.log

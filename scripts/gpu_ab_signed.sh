#!/bin/bash
# same-box A/B: signed 4-bit windows (8-entry h*A table) vs the unsigned 15-entry table (variants/libbsx_unsigned.so)
OUT=gpurun_out/${1:-ab_signed}
mkdir -p $OUT
echo "== pytest ed25519 + verify"; timeout 900 python -m pytest tests/test_gpu_ed25519.py tests/test_gpu_verify.py tests/test_gpu_header_range.py -m gpu -x -q 2>&1 | tail -3
run() { # tag, env...
  tag=$1; shift
  for n in 37888 100000; do
    echo "== $tag ed n=$n"; env "$@" timeout 300 python bench.py --mode ed25519 --sigs $n --steps 5 --warmup 3 --no-cpu 2>> $OUT/err.log | tee $OUT/ed_${tag}_$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
  done
  echo "== $tag header_range"; env "$@" timeout 300 python bench.py --no-cpu 2>> $OUT/err.log | tee $OUT/bench_$tag.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d.get('kernels_alone_ms'), d['e2e'])"
}
OLD=$PWD/blobstreamx_b200/csrc/build/variants/libbsx_unsigned.so
run signed A=1
run unsigned BSX_LIB_PATH=$OLD
run signed_occ6 BSX_ED_OCC=6
run signed2 A=1
run unsigned2 BSX_LIB_PATH=$OLD
tail -3 $OUT/err.log

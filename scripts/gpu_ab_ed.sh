#!/bin/bash
# A/B of Ed25519 field-multiplication variants: default lib vs libs under csrc/build/variants (BSX_LIB_PATH)
OUT=gpurun_out/${1:-ab_ed}
mkdir -p $OUT
echo "== pytest ed25519 + verify"; timeout 600 python -m pytest tests/test_gpu_ed25519.py tests/test_gpu_verify.py -m gpu -x -q 2>&1 | tail -3
for lib in default $(ls blobstreamx_b200/csrc/build/variants/*.so 2>/dev/null); do
  if [ $lib = default ]; then unset BSX_LIB_PATH; tag=default; else export BSX_LIB_PATH=$PWD/$lib; tag=$(basename $lib .so); fi
  for n in 25600 100000; do
    echo "== $tag n=$n"; timeout 300 python bench.py --mode ed25519 --sigs $n --steps 5 --warmup 3 --no-cpu 2>> $OUT/err.log | tee $OUT/ed_${tag}_$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
  done
done
unset BSX_LIB_PATH
echo "== header_range"; timeout 300 python bench.py --no-cpu 2>> $OUT/err.log | tee $OUT/bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernels_alone_ms'], d['e2e'])"
tail -3 $OUT/err.log

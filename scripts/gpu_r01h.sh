#!/bin/bash
OUT=gpurun_out/r01h
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
echo "== pytest goldilocks"; timeout 600 python -m pytest tests/test_gpu_goldilocks.py -m gpu -x -q 2>&1 | tail -4
echo "== gates"; timeout 300 python bench.py --mode gates --steps 10 --warmup 3 --no-cpu 2>> $OUT/bench.err | tee $OUT/gates.json | cut -c1-250
for occ in 4 6 8; do
  echo "== ed25519 occ=$occ n=100000"; BSX_ED_OCC=$occ timeout 300 python bench.py --mode ed25519 --sigs 100000 --steps 5 --warmup 3 --no-cpu 2>> $OUT/bench.err | tee $OUT/ed_occ${occ}_100000.json | cut -c1-200
  echo "== ed25519 occ=$occ n=25600"; BSX_ED_OCC=$occ timeout 300 python bench.py --mode ed25519 --sigs 25600 --steps 5 --warmup 3 --no-cpu 2>> $OUT/bench.err | tee $OUT/ed_occ${occ}_25600.json | cut -c1-200
  echo "== header_range occ=$occ"; BSX_ED_OCC=$occ timeout 300 python bench.py --no-cpu 2>> $OUT/bench.err | tee $OUT/bench_occ$occ.json | cut -c1-200
done
for T in 16 256 4096; do
  echo "== tree T=$T"; timeout 300 python bench.py --mode tree --trees $T --steps 10 --warmup 3 2>> $OUT/bench.err | tee $OUT/tree_$T.json | cut -c1-250
done
tail -3 $OUT/bench.err

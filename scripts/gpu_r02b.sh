#!/bin/bash
# r02b: Ed25519 field arithmetic on the FP64 pipe -- parity of every build, then A/B against the IMAD.WIDE form
OUT=gpurun_out/r02b
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
echo "== pytest builds"; timeout 900 python -m pytest tests/test_gpu_ed25519_builds.py tests/test_gpu_ed25519.py -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_builds.log
for fp in 0 1; do
  for inl in 0 1; do
    for n in 37888 100000; do
      echo "== ed25519 n=$n fp64=$fp inline=$inl"
      BSX_ED_FP64=$fp BSX_ED_INLINE=$inl timeout 300 python bench.py --mode ed25519 --sigs $n --steps 10 --warmup 3 --no-cpu 2>> $OUT/bench.err | tee $OUT/ed_fp${fp}_inl${inl}_$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e6, 'M sig/s', d['ms_per_step'], 'ms')"
    done
  done
done
for fp in 0 1; do
  echo "== header_range fp64=$fp"
  BSX_ED_FP64=$fp timeout 600 python bench.py --no-cpu --steps 20 --warmup 5 2>> $OUT/bench.err | tee $OUT/bench_fp$fp.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value']/1e6, 'ms', d['ms_per_step'], 'alone', d['kernels_alone_ms'], 'e2e', d['e2e']['value']/1e6, d['e2e']['single_call']/1e6, '2048:', d['header_range_2048']['value']/1e6, d['header_range_2048']['ms_per_step'])"
done
for occ in 6 8; do
  echo "== header_range fp64=1 occ=$occ"
  BSX_ED_FP64=1 BSX_ED_OCC=$occ timeout 600 python bench.py --no-cpu --steps 20 --warmup 5 --no-2048 2>> $OUT/bench.err | tee $OUT/bench_fp1_occ$occ.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value']/1e6, 'ms', d['ms_per_step'], 'alone', d['kernels_alone_ms'])"
done
echo "== header_range fp64=1 regs=0 (216/240-register build)"
BSX_ED_FP64=1 BSX_ED_REGS=0 timeout 600 python bench.py --no-cpu --steps 20 --warmup 5 --no-2048 2>> $OUT/bench.err | tee $OUT/bench_fp1_regs0.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value']/1e6, 'ms', d['ms_per_step'], 'alone', d['kernels_alone_ms'])"
tail -5 $OUT/bench.err

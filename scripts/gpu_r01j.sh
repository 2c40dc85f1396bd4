#!/bin/bash
# r01j: pair-Karatsuba fe_mul + 8-bit s*G windows + chunked copy/compute pipeline in bsx_header_range
OUT=gpurun_out/r01j
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.log
for n in 25600 100000; do
  echo "== ed n=$n"; timeout 300 python bench.py --mode ed25519 --sigs $n --steps 5 --warmup 3 --no-cpu 2>> $OUT/err.log | tee $OUT/ed_$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
done
for ch in 0 8 16 32 64 256; do
  for er in 64 256; do
  echo "== header_range chunk=$ch e2e_ranges=$er"; BSX_PIPE_CHUNK=$ch timeout 300 python bench.py --no-cpu --e2e-ranges $er 2>> $OUT/err.log | tee $OUT/bench_chunk${ch}_e$er.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernels_alone_ms'], d['e2e'])"
  done
done
tail -3 $OUT/err.log

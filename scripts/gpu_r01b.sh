#!/bin/bash
# r01b: Ed25519 parity + throughput, SHA-256 compress variants A/B
OUT=gpurun_out/r01b
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
V=blobstreamx_b200/csrc/build/variants
for v in sha0 sha1; do
  echo "== bench variant $v"; BSX_LIB_PATH=$V/libbsx_$v.so timeout 300 python bench.py --no-cpu 2>> $OUT/bench.err | tee $OUT/bench_$v.json
done
echo "== bench default (sha2)"; timeout 300 python bench.py 2>> $OUT/bench.err | tee $OUT/bench_sha2.json
for n in 100 1000 10000 100000; do
  echo "== ed25519 n=$n"; timeout 300 python bench.py --mode ed25519 --sigs $n --steps 5 --warmup 3 2>> $OUT/bench.err | tee $OUT/ed_$n.json
done
echo "== ncu ed25519"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ed25519_batch_kernel -s 1 -c 1 -f -o $OUT/prof_ed \
    python bench.py --mode ed25519 --sigs 37888 --steps 1 --warmup 3 --no-cpu --no-check > $OUT/ncu_ed.log 2>&1
echo "== ncu map kernel (sha2 variant, R=256)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prove_subchain_kernel -s 4 -c 1 -f -o $OUT/prof_map \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-check > $OUT/ncu_map.log 2>&1
tail -3 $OUT/bench.err
ls -la $OUT

#!/bin/bash
# r01p: signed-window Ed25519, 192-register cap beside the hash kernels, skip hashes on a third stream, device-side encoders
OUT=gpurun_out/r01p
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench"; timeout 600 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-200; tail -3 $OUT/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> $OUT/bench.err | tee $OUT/bench_ref.json | cut -c1-200
for r in 256 512 756; do
  echo "== bench ranges=$r"; timeout 300 python bench.py --ranges $r --no-cpu 2>> $OUT/bench.err | tee $OUT/bench_r$r.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['single_call'])"
done
for n in 100 10000 37888 100000; do
  echo "== ed25519 n=$n"; timeout 300 python bench.py --mode ed25519 --sigs $n --steps 5 --warmup 3 --no-cpu 2>> $OUT/bench.err | tee $OUT/ed_$n.json | cut -c1-160
done
echo "== encode"; timeout 300 python bench.py --mode encode --steps 10 --warmup 3 2>> $OUT/bench.err | tee $OUT/encode.json | cut -c1-160
echo "== shape"; timeout 300 python bench.py --mode shape --steps 10 --warmup 3 --no-cpu 2>> $OUT/bench.err | tee $OUT/shape.json | cut -c1-160
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-check --e2e-threads 1 > $OUT/ncu_bench.log 2>&1
echo "== ncu ed25519 capped"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ed25519_batch_kernel_capped -s 2 -c 1 -f -o $OUT/prof_ed_capped \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-check --e2e-threads 1 > $OUT/ncu_ed.log 2>&1
echo "== ncu encode"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_headers_kernel -s 2 -c 1 -f -o $OUT/prof_encode \
    python bench.py --mode encode --steps 2 --warmup 3 --no-cpu > $OUT/ncu_encode.log 2>&1
echo "== sanitizer (encoders)"
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_encode.py -m gpu -x -q 2>&1 | tail -4 | tee $OUT/sanitizer_encode.log
tail -3 $OUT/bench.err

#!/bin/bash
# r02h: full GPU test suite, smoke, default bench (one wave of the 8-CTA Ed25519 build = 757 ranges), reference arm, ncu of this build
OUT=gpurun_out/r02h
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench"; /usr/bin/time -v timeout 900 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value']/1e6, 'ms', d['ms_per_step'], d['step_ms'], d['clocks']); print('alone', d['kernels_alone_ms']); print('e2e', d['e2e']['value']/1e6, d['e2e']['single_call']/1e6, 'lat', d['latency_single_range_ms']); print('2048:', d['header_range_2048']['value']/1e6, d['header_range_2048']['ms_per_step']); print('roofline', d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['pipe'], d['roofline']['capture_stale']); print('cpu', d['cpu_baseline'], d['cpu_library_baseline'])"
grep -E "Elapsed|Maximum resident" $OUT/bench.err; grep -v "Elapsed\|Maximum\|^\s" $OUT/bench.err | tail -5
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> $OUT/bench.err | tee $OUT/bench_ref.json | cut -c1-200
ARGS="--steps 2 --warmup 3 --no-cpu --no-check --e2e-threads 1 --no-2048"
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py $ARGS > $OUT/ncu_bench.log 2>&1
cap() { echo "== ncu full $1"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $OUT/prof_$1 python bench.py $ARGS > $OUT/ncu_$1.log 2>&1; ls -la $OUT/prof_$1.ncu-rep 2>/dev/null | awk '{print $5}'; }
cap ed25519 "ed25519_batch_kernel" 2
cap subchain_proofs "subchain_proofs_kernel" 4
cap subchain_commit "subchain_commit_kernel" 4

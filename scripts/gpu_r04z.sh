#!/bin/bash
# r04z: closing single-GPU run: tests, smoke, default bench + reference arm, sweeps / ed25519 / plonk / trace modes, launch list, (ncu: scripts/gpu_r02p_ncu.sh)
OUT=gpurun_out/r04z
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench"; S=$(date +%s); timeout 900 python bench.py 2> $OUT/bench.err > $OUT/bench.json; echo "wall $(( $(date +%s) - S )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r04z/bench.json').read())
print('value', d['value']/1e6, 'ms', d['ms_per_step'], d['step_ms'], d['clocks']); print('next_header', d['next_header'])
print('general path', d['general_path']['value']/1e6, d['general_path']['ms_per_step'])
print('alone', d['kernels_alone_ms'])
print('e2e', d['e2e']['value']/1e6, d['e2e'].get('single_call'), 'lat', d['latency_single_range_ms'], d['latency_single_range_c_abi_pinned_ms'])
print('2048:', d['header_range_2048']['value']/1e6, d['header_range_2048']['ms_per_step'])
print('roofline', d['roofline']['frac'], d['roofline']['signatures_per_s'], d['roofline']['capture_stale'], 'map', d['roofline_map']['frac'])
print('constraints', d['constraints']['value']/1e9, d['constraints']['roofline']['frac'])
print('cpu', d['cpu_baseline'], d['cpu_library_baseline'])
PY
tail -3 $OUT/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> $OUT/bench.err | tee $OUT/bench_ref.json | cut -c1-260
echo "== modes"
for m in sweeps ed25519 tree shape plonk trace gates poseidon; do
  timeout 600 python bench.py --mode $m 2>> $OUT/bench.err > $OUT/mode_$m.json; python -c "
import json; d=json.loads(open('$OUT/mode_$m.json').read()); print('$m', d.get('metric'), d.get('value'), d.get('unit'), d.get('ms_per_step'))"
done
ARGS="--steps 2 --warmup 3 --no-cpu --no-check --e2e-threads 1 --e2e-ranges 64 --no-2048"
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py $ARGS > $OUT/ncu_bench.log 2>&1

#!/bin/bash
# r02i: default bench line of this build (wall-clocked), reference arm, ncu summaries refreshed from the r02h captures' successor
OUT=gpurun_out/r02i
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/nproc.txt
echo "== bench"; T0=$(date +%s.%N)
timeout 900 python bench.py 2> $OUT/bench.err > $OUT/bench.json
T1=$(date +%s.%N); echo "wall $(echo "$T1 - $T0" | bc) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02i/bench.json').read())
print('value', d['value']/1e6, 'ms', d['ms_per_step'], d['step_ms'], d['clocks'])
print('alone', d['kernels_alone_ms'])
print('e2e', d['e2e']['value']/1e6, d['e2e'].get('single_call'), 'lat', d['latency_single_range_ms'])
print('2048:', d['header_range_2048']['value']/1e6, d['header_range_2048']['ms_per_step'])
print('roofline', d['roofline'])
print('cpu', d['cpu_baseline'], d['cpu_library_baseline'])
print('config', d['config'])
PY
tail -5 $OUT/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> $OUT/bench.err | tee $OUT/bench_ref.json | cut -c1-300

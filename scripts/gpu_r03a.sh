#!/bin/bash
# quad-lane path with the decompression and inversion chains on the FP64 pipe: tests, single-range latency, small batches
OUT=gpurun_out/${1:-r03a}
mkdir -p $OUT
echo "== pytest ed25519 + verify"; timeout 1500 python -m pytest tests/test_gpu_ed25519_builds.py tests/test_gpu_ed25519.py tests/test_gpu_verify.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu --no-2048 --steps 20 --warmup 5 2>> $OUT/err.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value', round(d['value']/1e6,1), 'lat python', round(d['latency_single_range_ms'],3), 'lat C ABI pinned', round(d['latency_single_range_c_abi_pinned_ms'],3), 'next_header', round(d['next_header']['latency_ms'],3))"
python scripts/latency_single.py 2>> $OUT/err.log | tail -2
timeout 600 python bench.py --mode sweeps --steps 20 --warmup 5 2>> $OUT/err.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print([(p['signatures'], round(p['ms'],3)) for p in d['ed25519_sweep']['points']])"
tail -2 $OUT/err.log

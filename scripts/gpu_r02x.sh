#!/bin/bash
# r02x: verify_kernel with the word-wise trusted-set compare: verify tests + the step
OUT=gpurun_out/r02x
mkdir -p $OUT
echo "== pytest verify"; timeout 900 python -m pytest tests/test_gpu_verify.py tests/test_gpu_encode.py tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -3
run() { local r=$1; shift
  env "$@" timeout 300 python bench.py --ranges $r --no-cpu --e2e-threads 1 --e2e-ranges 64 --no-2048 --steps 20 --warmup 5 2>> $OUT/err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ranges=$r $*', round(d['value']/1e6,1), 'M headers/s', round(d['ms_per_step'],3), 'ms', {k[:14]: round(v,3) for k,v in d['kernels_alone_ms'].items()}, 'general', round(d['general_path']['value']/1e6,1), 'lat', round(d['latency_single_range_ms'],3))"
}
run 757 BSX_X=0
run 757 BSX_X=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:verify_kernel -c 6 --csv python bench.py --steps 2 --warmup 3 --no-cpu --no-check --e2e-threads 1 --e2e-ranges 64 --no-2048 2>/dev/null | grep verify_kernel | cut -d, -f5,12- | tail -3
tail -2 $OUT/err.log

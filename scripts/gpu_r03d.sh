#!/bin/bash
# pairing rule re-checked on the final build (word stores changed the balance)
OUT=gpurun_out/r03d
mkdir -p $OUT
run() { local r=$1; shift
  env "$@" timeout 300 python bench.py --ranges $r --no-cpu --no-check --e2e-threads 1 --e2e-ranges 16 --no-2048 --steps 20 --warmup 5 2>> $OUT/err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ranges=$r $*', round(d['value']/1e6,1), 'M headers/s', round(d['ms_per_step'],3), 'ms', {k[:14]: round(v,3) for k,v in d['kernels_alone_ms'].items()})"
}
for r in 568 757 946 1135; do run $r BSX_ED_PAIR=0; run $r BSX_ED_PAIR=1; done
tail -2 $OUT/err.log

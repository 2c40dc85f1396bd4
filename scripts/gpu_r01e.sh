#!/bin/bash
OUT=gpurun_out/r01e
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
echo "== bench default (split map, occ8)"; timeout 600 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-220; 
echo "== bench split occ6"; BSX_PROOFS_OCC=6 timeout 600 python bench.py --no-cpu 2>> $OUT/bench.err | tee $OUT/bench_occ6.json | cut -c1-220
echo "== bench fused"; BSX_SUBCHAIN_FUSED=1 timeout 600 python bench.py --no-cpu 2>> $OUT/bench.err | tee $OUT/bench_fused.json | cut -c1-220
for f in bench bench_occ6 bench_fused; do python - <<PY
import json
d=json.load(open("$OUT/$f.json"))
print("$f", "ms/step %.3f"%d["ms_per_step"], "value %.1fM"%(d["value"]/1e6), d["kernels_alone_ms"], "e2e %.1fM"%(d["e2e"]["value"]/1e6))
PY
done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-check > $OUT/ncu_bench.log 2>&1
echo "== ncu full proofs kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:subchain_proofs_kernel -s 4 -c 1 -f -o $OUT/prof_proofs \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-check > $OUT/ncu_full.log 2>&1
tail -3 $OUT/bench.err

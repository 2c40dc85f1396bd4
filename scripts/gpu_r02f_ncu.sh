#!/bin/bash
# r02f: ncu launch list of the bench command + full captures of the hot kernels of THIS build -> profiles/ncu_summary.json
OUT=gpurun_out/r02f
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
ARGS="--steps 2 --warmup 3 --no-cpu --no-check --e2e-threads 1 --no-2048"
echo "== pytest plonk"; timeout 600 python -m pytest tests/test_gpu_plonk.py -m gpu -q 2>&1 | tail -4 | tee $OUT/pytest_plonk.log
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py $ARGS > $OUT/ncu_bench.log 2>&1
grep -c "gpu__time_duration" $OUT/launches.csv
cap() {  # name regex skip
  echo "== ncu full $1"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $OUT/prof_$1 python bench.py $ARGS > $OUT/ncu_$1.log 2>&1
  ls -la $OUT/prof_$1.ncu-rep 2>/dev/null | awk '{print $5}'
}
cap ed25519 "ed25519_batch_kernel" 2
cap subchain_proofs "subchain_proofs_kernel" 4
cap subchain_commit "subchain_commit_kernel" 4
echo "== ncu plonk"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ntt_dif_strided_kernel|ntt_dif_contig_kernel|gl_merkle_leaves_kernel|gl_gate_quotient_kernel" -s 12 -c 6 -f -o $OUT/prof_plonk python bench.py --mode plonk --log-rows 18 --steps 2 --no-cpu --no-check > $OUT/ncu_plonk.log 2>&1
echo "== ncu trace"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sha256_trace_kernel -s 40 -c 1 -f -o $OUT/prof_trace python bench.py --mode trace --steps 2 --no-cpu --no-check > $OUT/ncu_trace.log 2>&1
ls -la $OUT

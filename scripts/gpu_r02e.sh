#!/bin/bash
# r02e: prover inner loops (NTT / LDE / Merkle caps / quotient / FRI fold), sign-bit assertions, latest-block input shaping, sweeps
OUT=gpurun_out/r02e
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
echo "== pytest plonk"; timeout 900 python -m pytest tests/test_gpu_plonk.py -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_plonk.log
echo "== pytest all"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
echo "== plonk bench 2^18"; timeout 600 python bench.py --mode plonk --log-rows 18 --steps 10 2>> $OUT/bench.err | tee $OUT/plonk_18.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step']); [print('  ', k, round(v['ms'],3), 'ms', round(v['GBps']), 'GB/s', round(v['frac_of_hbm_peak'],3)) for k,v in d['stages'].items()]"
echo "== plonk bench 2^20"; timeout 600 python bench.py --mode plonk --log-rows 20 --steps 5 --no-cpu 2>> $OUT/bench.err | tee $OUT/plonk_20.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step']); [print('  ', k, round(v['ms'],3), 'ms', round(v['GBps']), 'GB/s', round(v['frac_of_hbm_peak'],3)) for k,v in d['stages'].items()]"
echo "== sweeps"; timeout 900 python bench.py --mode sweeps --steps 10 2>> $OUT/bench.err | tee $OUT/sweeps.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); [print('  T', p['trees'], round(p['ms'],3), 'ms', round(p['algorithmic_GBps']), 'GB/s') for p in d['data_commitment_tree_sweep']['points']]; [print('  n', p['signatures'], round(p['ms'],3), 'ms', round(p['sigs_per_s']/1e6,2), 'M sig/s') for p in d['ed25519_sweep']['points']]"
tail -5 $OUT/bench.err
for ct in 128 256 512 1024; do
  echo "== header_range COMMIT_THREADS=$ct"
  BSX_COMMIT_THREADS=$ct timeout 600 python bench.py --no-cpu --steps 20 --warmup 5 2>> $OUT/bench.err | tee $OUT/bench_ct$ct.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value']/1e6, 'ms', d['ms_per_step'], 'alone', d['kernels_alone_ms'], '2048:', d['header_range_2048']['ms_per_step'], d['header_range_2048']['roofline_map']['kernel_ms'])"
done
echo "== trace"; timeout 600 python bench.py --mode trace --steps 10 2>> $OUT/bench.err | tee $OUT/trace.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e6, 'M rows/s', d['ms_per_step'], 'ms', d['roofline']['achieved'], 'GB/s', d['roofline']['frac'])"
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
tail -5 $OUT/bench.err

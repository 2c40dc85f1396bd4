#!/bin/bash
# Re-entry round: full GPU parity suite, smoke, both bench arms, gates / Ed25519 occupancy / tree sweeps, launch list, ncu captures.
OUT=gpurun_out/r01i
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench"; timeout 600 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-400; tail -3 $OUT/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> $OUT/bench.err | tee $OUT/bench_ref.json | cut -c1-300
echo "== gates"; timeout 300 python bench.py --mode gates --steps 10 --warmup 3 2>> $OUT/bench.err | tee $OUT/gates.json | cut -c1-400
for occ in 4 6 8; do
  echo "== ed25519 occ=$occ n=100000"; BSX_ED_OCC=$occ timeout 300 python bench.py --mode ed25519 --sigs 100000 --steps 5 --warmup 3 --no-cpu 2>> $OUT/bench.err | tee $OUT/ed_occ${occ}_100000.json | cut -c1-200
  echo "== ed25519 occ=$occ n=25600"; BSX_ED_OCC=$occ timeout 300 python bench.py --mode ed25519 --sigs 25600 --steps 5 --warmup 3 --no-cpu 2>> $OUT/bench.err | tee $OUT/ed_occ${occ}_25600.json | cut -c1-200
  echo "== header_range occ=$occ"; BSX_ED_OCC=$occ timeout 300 python bench.py --no-cpu 2>> $OUT/bench.err | tee $OUT/bench_occ$occ.json | cut -c1-200
done
for T in 16 256 4096; do
  echo "== tree T=$T"; timeout 300 python bench.py --mode tree --trees $T --steps 10 --warmup 3 2>> $OUT/bench.err | tee $OUT/tree_$T.json | cut -c1-300
done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --ranges 64 --e2e-ranges 16 --no-cpu --no-check > $OUT/ncu_bench.log 2>&1
echo "== ncu gates"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gl_gate_eval_kernel -s 3 -c 1 -f -o $OUT/prof_gates \
    python bench.py --mode gates --steps 2 --warmup 3 --no-cpu --no-check > $OUT/ncu_gates.log 2>&1
echo "== ncu ed25519"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ed25519_batch_kernel -s 2 -c 1 -f -o $OUT/prof_ed \
    python bench.py --mode ed25519 --sigs 100000 --steps 2 --warmup 3 --no-cpu --no-check > $OUT/ncu_ed.log 2>&1
tail -3 $OUT/bench.err
ls -la $OUT

#!/bin/bash
# r01k: full round after the Ed25519 / pipeline / gate-kernel changes
OUT=gpurun_out/r01k
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench"; timeout 600 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-300; tail -3 $OUT/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> $OUT/bench.err | tee $OUT/bench_ref.json | cut -c1-300
echo "== gates"; timeout 300 python bench.py --mode gates --steps 10 --warmup 3 2>> $OUT/bench.err | tee $OUT/gates.json | cut -c1-300
for n in 100 1000 10000 25600 100000; do
  echo "== ed25519 n=$n"; timeout 300 python bench.py --mode ed25519 --sigs $n --steps 5 --warmup 3 --no-cpu 2>> $OUT/bench.err | tee $OUT/ed_$n.json | cut -c1-200
done
for T in 16 256 4096; do
  echo "== tree T=$T"; timeout 300 python bench.py --mode tree --trees $T --steps 10 --warmup 3 --no-cpu 2>> $OUT/bench.err | tee $OUT/tree_$T.json | cut -c1-200
done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-check > $OUT/ncu_bench.log 2>&1
echo "== ncu gates"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gl_gate_eval_kernel -s 3 -c 1 -f -o $OUT/prof_gates \
    python bench.py --mode gates --steps 2 --warmup 3 --no-cpu --no-check > $OUT/ncu_gates.log 2>&1
echo "== ncu ed25519"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ed25519_batch_kernel -s 2 -c 1 -f -o $OUT/prof_ed \
    python bench.py --mode ed25519 --sigs 25600 --steps 2 --warmup 3 --no-cpu --no-check > $OUT/ncu_ed.log 2>&1
echo "== ncu map proofs"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:subchain_proofs_kernel -s 2 -c 1 -f -o $OUT/prof_proofs \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-check > $OUT/ncu_proofs.log 2>&1
tail -3 $OUT/bench.err
ls -la $OUT | head -40

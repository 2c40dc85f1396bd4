#!/bin/bash
# r01o: final round of this session (wave-aligned batch, Poseidon MDS, inlined Ed25519 for standalone batches, attestation proofs)
OUT=gpurun_out/r01o
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench"; timeout 600 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-200; tail -3 $OUT/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> $OUT/bench.err | tee $OUT/bench_ref.json | cut -c1-200
for g in arithmetic add_many subtraction comparison range_check; do
  echo "== gates $g"; timeout 300 python bench.py --mode gates --gate $g --steps 10 --warmup 3 2>> $OUT/bench.err | tee $OUT/gates_$g.json | cut -c1-160
done
for n in 100 1000 10000 25600 100000; do
  echo "== ed25519 n=$n"; timeout 300 python bench.py --mode ed25519 --sigs $n --steps 5 --warmup 3 --no-cpu 2>> $OUT/bench.err | tee $OUT/ed_$n.json | cut -c1-160
done
for T in 16 256 4096; do
  echo "== tree T=$T"; timeout 300 python bench.py --mode tree --trees $T --steps 10 --warmup 3 --no-cpu 2>> $OUT/bench.err | tee $OUT/tree_$T.json | cut -c1-160
done
for L in 8 64; do
  echo "== poseidon len=$L"; timeout 300 python bench.py --mode poseidon --hash-len $L --steps 10 --warmup 3 2>> $OUT/bench.err | tee $OUT/poseidon_$L.json | cut -c1-160
done
echo "== shape"; timeout 300 python bench.py --mode shape --steps 10 --warmup 3 2>> $OUT/bench.err | tee $OUT/shape.json | cut -c1-160
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-check --e2e-threads 1 > $OUT/ncu_bench.log 2>&1
echo "== ncu map proofs"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:subchain_proofs_kernel -s 2 -c 1 -f -o $OUT/prof_proofs \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-check --e2e-threads 1 > $OUT/ncu_proofs.log 2>&1
echo "== ncu ed25519 quad"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ed25519_quad_kernel -s 2 -c 1 -f -o $OUT/prof_ed_quad \
    python bench.py --mode ed25519 --sigs 10000 --steps 2 --warmup 3 --no-cpu --no-check > $OUT/ncu_ed_quad.log 2>&1
echo "== ncu shape"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:range_inputs_kernel -s 2 -c 1 -f -o $OUT/prof_shape \
    python bench.py --mode shape --steps 2 --warmup 3 --no-cpu > $OUT/ncu_shape.log 2>&1
tail -3 $OUT/bench.err

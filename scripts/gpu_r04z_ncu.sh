#!/bin/bash
# r04z: ncu --set full captures of the bench command's top kernels (one launch each; gpurun brings back at most 64 MiB, so
# only the Ed25519 capture carries the source view)
OUT=gpurun_out/r04z
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
ARGS="--steps 2 --warmup 3 --no-cpu --no-check --e2e-threads 1 --e2e-ranges 64 --no-2048"
cap() { echo "== ncu full $1"; timeout 600 ncu --set full --clock-control none $4 -k regex:$2 -s $3 -c 1 -f -o $OUT/prof_$1 python bench.py $ARGS > $OUT/ncu_$1.log 2>&1; ls -la $OUT/prof_$1.ncu-rep 2>/dev/null | awk '{print $5}'; }
cap ed25519 "ed25519_keyed_kernel" 2 "--import-source on"
cap subchain_proofs "subchain_proofs_kernel" 4
cap subchain_commit "subchain_commit_kernel" 4
cap key_bases "ed25519_key_bases_kernel" 2
cap key_table "ed25519_key_table_kernel" 2
echo "== ncu full ed_trace (rows + chain8 kernels, bench.py --mode trace)"
timeout 600 ncu --set full --clock-control none -k regex:ed_trace_rows -s 2 -c 1 -f -o $OUT/prof_ed_trace_rows python bench.py --mode trace --steps 3 --no-cpu --no-check --trace-jobs 1 > $OUT/ncu_ed_trace_rows.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:ed_trace_chain8 -s 2 -c 1 -f -o $OUT/prof_ed_trace_chain8 python bench.py --mode trace --steps 3 --no-cpu --no-check --trace-jobs 1 > $OUT/ncu_ed_trace_chain8.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:sha256_trace -s 2 -c 1 -f -o $OUT/prof_sha256_trace python bench.py --mode trace --steps 3 --no-cpu --no-check --ed-trace-circuits 1 > $OUT/ncu_sha256_trace.log 2>&1
for k in ed_trace_rows ed_trace_chain8 sha256_trace; do python scripts/ncu_summary.py $OUT/prof_$k.ncu-rep profiles/r04z_${k}_ncu_full.csv; rm -f $OUT/prof_$k.ncu-rep; done
# summaries are made here (ncu is on the box); only the Ed25519 report itself travels back (64 MiB limit)
TAG=r04z python scripts/ncu_capture.py ed25519=$OUT/prof_ed25519.ncu-rep:75700:signatures subchain_proofs=$OUT/prof_subchain_proofs.ncu-rep:757:ranges \
    subchain_commit=$OUT/prof_subchain_commit.ncu-rep:757:ranges key_bases=$OUT/prof_key_bases.ncu-rep:100:keys key_table=$OUT/prof_key_table.ncu-rep:100:keys > $OUT/ncu_capture.log 2>&1
tail -3 $OUT/ncu_capture.log
mkdir -p $OUT/profiles; cp profiles/r04z_*_ncu_full.csv profiles/ncu_summary.json $OUT/profiles/
rm -f $OUT/prof_subchain_proofs.ncu-rep $OUT/prof_subchain_commit.ncu-rep $OUT/prof_key_bases.ncu-rep $OUT/prof_key_table.ncu-rep
du -sh gpurun_out

#!/bin/bash
# r02p: ncu --set full captures of the bench command's top kernels (one launch each; gpurun brings back at most 64 MiB, so
# only the Ed25519 capture carries the source view)
OUT=gpurun_out/r03c
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
ARGS="--steps 2 --warmup 3 --no-cpu --no-check --e2e-threads 1 --e2e-ranges 64 --no-2048"
cap() { echo "== ncu full $1"; timeout 600 ncu --set full --clock-control none $4 -k regex:$2 -s $3 -c 1 -f -o $OUT/prof_$1 python bench.py $ARGS > $OUT/ncu_$1.log 2>&1; ls -la $OUT/prof_$1.ncu-rep 2>/dev/null | awk '{print $5}'; }
cap ed25519 "ed25519_keyed_kernel" 2 "--import-source on"
cap subchain_proofs "subchain_proofs_kernel" 4
cap subchain_commit "subchain_commit_kernel" 4
cap key_bases "ed25519_key_bases_kernel" 2
cap key_table "ed25519_key_table_kernel" 2
# summaries are made here (ncu is on the box); only the Ed25519 report itself travels back (64 MiB limit)
TAG=r03c python scripts/ncu_capture.py ed25519=$OUT/prof_ed25519.ncu-rep:75700:signatures subchain_proofs=$OUT/prof_subchain_proofs.ncu-rep:757:ranges \
    subchain_commit=$OUT/prof_subchain_commit.ncu-rep:757:ranges key_bases=$OUT/prof_key_bases.ncu-rep:100:keys key_table=$OUT/prof_key_table.ncu-rep:100:keys > $OUT/ncu_capture.log 2>&1
tail -3 $OUT/ncu_capture.log
mkdir -p $OUT/profiles; cp profiles/r03c_*_ncu_full.csv profiles/ncu_summary.json $OUT/profiles/
rm -f $OUT/prof_subchain_proofs.ncu-rep $OUT/prof_subchain_commit.ncu-rep $OUT/prof_key_bases.ncu-rep $OUT/prof_key_table.ncu-rep
du -sh gpurun_out

#!/bin/bash
# every build of the thread-per-signature Ed25519 kernel against the oracle on the stress inputs (scripts/stress_ed25519.py)
OUT=gpurun_out/${1:-stress_variants}
mkdir -p $OUT
for cfg in "inlined_216 A=1" "compact_216 BSX_ED_INLINE=0" "capped_192 BSX_ED_REGS=1" "build_168 BSX_ED_OCC=6" "build_128 BSX_ED_OCC=8"; do
  set -- $cfg; tag=$1; shift
  echo "== $tag: $(env "$@" timeout 600 python scripts/stress_ed25519.py 30000 90000 2>&1 | tail -1)" | tee -a $OUT/stress.txt
done

#!/bin/bash
# r02y: compute-sanitizer over the new Ed25519 kernels (key tables, table-path kernel, word stores) and the verify / header_range paths that call them
OUT=gpurun_out/r02y
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
echo "== memcheck: per-key table path, overflow fallback, R shortcut edges"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ed25519_builds.py -m gpu -x -q -k "per_key or host_path" 2>&1 | tail -6 | tee $OUT/memcheck_ed.txt
echo "== memcheck: verify_* and header_range"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_verify.py -m gpu -x -q 2>&1 | tail -6 | tee $OUT/memcheck_verify.txt
echo "== racecheck: key assignment (atomicCAS table) and table kernels"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_ed25519_builds.py -m gpu -x -q -k "per_key_table_path and not repeat" 2>&1 | tail -6 | tee $OUT/racecheck_ed.txt

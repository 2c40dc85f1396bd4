#!/bin/bash
# r02d (8 GPUs): host-side PCIe ceiling with 8 concurrent ranks, then the sharded bench at N = 8 and N = 4
OUT=gpurun_out/r02d
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
nproc > $OUT/nproc.txt; nvidia-smi topo -m > $OUT/topo.txt 2>&1; numactl -H > $OUT/numa.txt 2>&1
echo "== pcie 1 rank"; timeout 120 python scripts/ubench/pcie.py --ranks 1 2>&1 | tail -1 | tee $OUT/pcie_1.json | cut -c1-300
echo "== pcie 8 ranks (bound)"; timeout 180 python scripts/ubench/pcie.py --ranks 8 2>&1 | tail -1 | tee $OUT/pcie_8.json | cut -c1-300
echo "== pcie 8 ranks (not bound)"; timeout 180 python scripts/ubench/pcie.py --ranks 8 --no-bind 2>&1 | tail -1 | tee $OUT/pcie_8_nobind.json | cut -c1-300
for n in 8 4; do
echo "== bench N=$n"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-cpu 2> $OUT/bench$n.err | tee $OUT/bench_n$n.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value']/1e6, 'ms', d['ms_per_step'], d['step_ms'], d['config']['sharding'][:120], 'e2e', d['e2e']['value']/1e6, '2048:', d['header_range_2048']['value']/1e6, d['header_range_2048']['ms_per_step'], d['header_range_2048']['step_ms'])"
tail -2 $OUT/bench$n.err
done

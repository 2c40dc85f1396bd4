#!/bin/bash
# r02r (N GPUs): the sharded bench of this build exactly as the driver launches it, at every N the box has, plus the distributed GPU tests
N=${1:-2}
OUT=gpurun_out/r04z_multi_$N
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
nproc > $OUT/nproc.txt
echo "== distributed GPU tests"; timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -3
for n in $N $((N/2)) 1; do
[ $n -ge 1 ] || continue
echo "== bench N=$n"
if [ $n -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n"; fi
timeout 900 $L bench.py --gpus $n --steps 20 --warmup 5 2> $OUT/bench$n.err | tee $OUT/bench_n$n.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value']/1e6, 'ms', d['ms_per_step'], d['step_ms'], d['config']['sharding'][:100], 'general', d['general_path']['value']/1e6, 'e2e', d['e2e']['value']/1e6, '2048:', d['header_range_2048']['value']/1e6, d['header_range_2048']['ms_per_step'])"
tail -2 $OUT/bench$n.err
[ $n -eq 1 ] && break
done
echo "== reference arm N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29699 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>> $OUT/ref.err | tail -1 | cut -c1-200

#!/bin/bash
# Multi-GPU round: bench at N ranks (torchrun) for the default config and config C
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in B C; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 --config $cfg > gpurun_out/bench_${cfg}_n$N.log 2>&1
echo "rc=$?" >> gpurun_out/bench_${cfg}_n$N.log
grep '^{' gpurun_out/bench_${cfg}_n$N.log | cut -c1-260
done

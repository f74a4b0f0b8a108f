#!/bin/bash
# Multi-GPU round on N ranks (torchrun): bench lines for the configs given (default B), written to gpurun_out/
N=${1:-2}; shift
CFGS=${@:-B}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in $CFGS; do
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 --config $cfg > gpurun_out/bench_${cfg}_n$N.log 2>&1
echo "rc=$?" >> gpurun_out/bench_${cfg}_n$N.log
grep '^{' gpurun_out/bench_${cfg}_n$N.log | cut -c1-200
tail -2 gpurun_out/bench_${cfg}_n$N.log | cut -c1-300
done

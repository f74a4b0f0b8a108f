#!/bin/bash
# One GPU-box round: parity tests, smoke, bench (both arms), ncu launch list + full capture of the fused kernel.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r01}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/${TAG}_bench.log 2>&1
timeout 600 python bench.py --config Bc --cpu-seconds 0 > gpurun_out/${TAG}_bench_concat.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.log 2>&1
BENCH="python bench.py --steps 2 --warmup 3 --user-block 1024 --cpu-seconds 0"
timeout 300 $BENCH > gpurun_out/${TAG}_bench_small.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_fused -s 3 -c 1 -o gpurun_out/${TAG}_prof_fused $BENCH > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -n 3 gpurun_out/${TAG}_pytest_gpu.log gpurun_out/${TAG}_smoke.log
tail -c 1500 gpurun_out/${TAG}_bench.log

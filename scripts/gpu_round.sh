#!/bin/bash
# One GPU-box round: parity tests, smoke, bench, ncu launch list + full capture of the fused kernel.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.log 2>&1
BENCH="python bench.py --steps 2 --warmup 3 --user-block 1024 --cpu-seconds 0"
timeout 300 $BENCH > gpurun_out/bench_small.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
timeout 300 $BENCH > gpurun_out/bench_small2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_gated -s 3 -c 1 -o gpurun_out/prof_gated $BENCH > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log
tail -c 1500 gpurun_out/bench.log

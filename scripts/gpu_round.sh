#!/bin/bash
# One GPU-box round: parity tests, smoke, bench (both arms + extra configs), stage benches, ncu launch list + full captures.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r01}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/${TAG}_bench.log 2>&1
timeout 600 python bench.py --config Bc --cpu-seconds 0 > gpurun_out/${TAG}_bench_concat.log 2>&1
timeout 600 python bench.py --config C --steps 8 --cpu-seconds 0 > gpurun_out/${TAG}_bench_attention.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.log 2>&1
for f in gated concatenate attention; do timeout 120 python scripts/bench_stages.py --fusion $f; done > gpurun_out/${TAG}_stages.jsonl 2>&1
BENCH="python bench.py --steps 2 --warmup 3 --cpu-seconds 0"
timeout 300 $BENCH > gpurun_out/${TAG}_bench_small.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_fused -s 3 -c 1 -o gpurun_out/${TAG}_prof_fused $BENCH > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_fused -s 3 -c 1 -o gpurun_out/${TAG}_prof_fused_attention $BENCH --config C > gpurun_out/${TAG}_ncu_full_attention.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"gemm3x|merge_topk|metrics_warp" -c 6 -o gpurun_out/${TAG}_prof_stages python scripts/bench_stages.py --fusion gated > gpurun_out/${TAG}_ncu_stages.log 2>&1
tail -n 3 gpurun_out/${TAG}_pytest_gpu.log gpurun_out/${TAG}_smoke.log
tail -c 600 gpurun_out/${TAG}_bench.log

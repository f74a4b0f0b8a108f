#!/bin/bash
# Final check of a round: GPU tests, smoke, the bench lines (both arms), stage benches, sweep.  No ncu.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r01f}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/${TAG}_bench.log 2>&1
timeout 600 python bench.py --config Bc --cpu-seconds 0 > gpurun_out/${TAG}_bench_concat.log 2>&1
timeout 600 python bench.py --config C --steps 8 --cpu-seconds 0 > gpurun_out/${TAG}_bench_attention.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.log 2>&1
for f in gated concatenate attention; do timeout 120 python scripts/bench_stages.py --fusion $f; done > gpurun_out/${TAG}_stages.jsonl 2>/dev/null
timeout 600 python scripts/sweep.py > gpurun_out/${TAG}_sweep.jsonl 2>/dev/null
tail -n 3 gpurun_out/${TAG}_pytest_gpu.log gpurun_out/${TAG}_smoke.log
grep '^{' gpurun_out/${TAG}_bench.log | cut -c1-200
# stage kernels (K1/K4/K5) under ncu, after the plain runs above
if [ -n "$NCU_STAGES" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"merge_topk|metrics_warp|gemm3x" -c 12 -o gpurun_out/${TAG}_prof_stages -f python scripts/bench_stages.py > gpurun_out/${TAG}_ncu_stages.log 2>&1
ncu -i gpurun_out/${TAG}_prof_stages.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_stages_raw.csv 2>/dev/null
fi

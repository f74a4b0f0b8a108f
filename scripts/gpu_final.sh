#!/bin/bash
# Final check of a round on one GPU: GPU tests, smoke, bench lines (both arms), stage benches, sweeps of the round's new axes,
# ncu launch list of the bench command + full captures of the fused kernel and of the K5 metrics kernel.
# A number printed by a run under ncu is never used as a bench value.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02f}
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -80 > gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/${TAG}_bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.log 2>&1
for f in gated concatenate attention; do timeout 120 python scripts/bench_stages.py --fusion $f; done > gpurun_out/${TAG}_stages.jsonl 2>/dev/null
timeout 400 python scripts/sweep.py --fusion gated --items 100000 --batch 4096 --activation relu gelu tanh leaky_relu silu --top-k 50 100 200 > gpurun_out/${TAG}_sweep_act_topk.jsonl 2>/dev/null
timeout 300 python scripts/sweep.py --fusion concatenate attention --items 100000 --batch 4096 --activation silu --top-k 50 128 >> gpurun_out/${TAG}_sweep_act_topk.jsonl 2>/dev/null
if [ -z "$NO_NCU" ]; then
BENCH="python bench.py --steps 2 --warmup 3 --no-also --cpu-seconds 0 --literal-seconds 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_fused -s 3 -c 1 -o gpurun_out/${TAG}_prof_fused $BENCH > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 300 ncu -i gpurun_out/${TAG}_prof_fused.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_fused_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:"metrics_warp|merge_topk_reg|gemm3x" -c 8 -o gpurun_out/${TAG}_prof_stages python scripts/bench_stages.py --fusion gated --eager > gpurun_out/${TAG}_ncu_stages.log 2>&1
timeout 300 ncu -i gpurun_out/${TAG}_prof_stages.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_stages_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_prof_fused.ncu-rep gpurun_out/${TAG}_prof_stages.ncu-rep
fi
tail -n 3 gpurun_out/${TAG}_pytest_gpu.log gpurun_out/${TAG}_smoke.log
grep '^{' gpurun_out/${TAG}_bench.log | cut -c1-300

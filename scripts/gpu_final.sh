#!/bin/bash
# Final check of a round on one GPU: GPU tests, smoke, bench lines (both arms), stage benches.  No ncu.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02f}
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/${TAG}_bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.log 2>&1
for f in gated concatenate attention; do timeout 120 python scripts/bench_stages.py --fusion $f; done > gpurun_out/${TAG}_stages.jsonl 2>/dev/null
tail -n 3 gpurun_out/${TAG}_pytest_gpu.log gpurun_out/${TAG}_smoke.log
grep '^{' gpurun_out/${TAG}_bench.log | cut -c1-200

#!/bin/bash
# Final check of the round-2 tree on one GPU: every GPU test, smoke, the default bench line (both legs of cpu_baseline inside).
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02g}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -40 > gpurun_out/${TAG}_pytest_gpu.log
tail -n 3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
tail -n 4 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.log 2>&1
grep '^{' gpurun_out/${TAG}_bench.log | cut -c1-400

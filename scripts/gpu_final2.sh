#!/bin/bash
# Final check of the round-2 tree on one GPU: every GPU test (all failures reported), optionally smoke + the default bench line.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02g}
timeout 1000 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -60 > gpurun_out/${TAG}_pytest_gpu.log
tail -n 6 gpurun_out/${TAG}_pytest_gpu.log
if [ -n "$WITH_BENCH" ]; then
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
tail -n 4 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.log 2>&1
grep '^{' gpurun_out/${TAG}_bench.log | cut -c1-400
fi

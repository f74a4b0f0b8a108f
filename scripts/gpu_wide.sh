#!/bin/bash
# Wide gated front end (F_GATEDW): parity tests (one small case first, under a short timeout: a pipeline bug would hang),
# embedding-dim sweep, regression subset of the neighbouring front ends.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02w}
python -c "import torch; print(torch.cuda.get_device_name(0))"
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gated_any_embedding_dim and 128-40-700-50-kw0" 2>&1 | tail -5 > gpurun_out/${TAG}_pytest_first.log
cat gpurun_out/${TAG}_pytest_first.log
grep -q " passed" gpurun_out/${TAG}_pytest_first.log || exit 1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "gated_any_embedding_dim or gated_wide" 2>&1 | grep -v "^$" | tail -60 > gpurun_out/${TAG}_pytest_wide.log
tail -n 5 gpurun_out/${TAG}_pytest_wide.log
timeout 200 python scripts/sweep.py --fusion gated --items 100000 --dims 128 512 --batch 64 4096 > gpurun_out/${TAG}_sweep_gated_dims.jsonl 2> gpurun_out/${TAG}_sweep.err
cut -c1-420 gpurun_out/${TAG}_sweep_gated_dims.jsonl; tail -n 3 gpurun_out/${TAG}_sweep.err

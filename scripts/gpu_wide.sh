#!/bin/bash
# Wide gated front end (F_GATEDW): parity tests, regression subset of the neighbouring front ends, embedding-dim sweep.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02w}
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "gated_any_embedding_dim or gated_wide" 2>&1 | grep -v "^$" | tail -60 > gpurun_out/${TAG}_pytest_wide.log
tail -n 5 gpurun_out/${TAG}_pytest_wide.log
timeout 300 python scripts/sweep.py --fusion gated --items 100000 --dims 128 256 512 --batch 64 4096 > gpurun_out/${TAG}_sweep_gated_dims.jsonl 2> gpurun_out/${TAG}_sweep.err
cut -c1-400 gpurun_out/${TAG}_sweep_gated_dims.jsonl; tail -n 3 gpurun_out/${TAG}_sweep.err
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "matches_emulated_and_exact_oracle or concat_any_embedding_dim or small_batch or smaller_mlp" 2>&1 | grep -v "^$" | tail -15 > gpurun_out/${TAG}_pytest_regress.log
tail -n 4 gpurun_out/${TAG}_pytest_regress.log

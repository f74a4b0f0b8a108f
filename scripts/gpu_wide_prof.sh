#!/bin/bash
# Wide gated front end: embedding-dim x catalogue x batch sweep, then ncu (launch list + one full capture) of one sweep point.
# A number printed by a run under ncu is never used as a bench value.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02g}
timeout 240 python scripts/sweep.py --fusion gated --items 100000 1000000 --dims 64 128 256 512 --batch 1 64 4096 > gpurun_out/${TAG}_sweep_gated_dims.jsonl 2> gpurun_out/${TAG}_sweep.err
python - <<'PY'
import json
for l in open("gpurun_out/r02g_sweep_gated_dims.jsonl"):
    d = json.loads(l); print(d["n_items"], d["embedding_dim"], d["user_batch"], round(d["ms_per_call"], 3), round(d["kernel_ms"], 3), round(d["pairs_per_s"] / 1e9, 3), round(d["frac_of_bf16_peak"], 3), d["path"])
PY
tail -n 2 gpurun_out/${TAG}_sweep.err
PT="python scripts/sweep.py --fusion gated --items 100000 --dims 128 --batch 4096 --min-ms 1"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_wide_launches.csv $PT > gpurun_out/${TAG}_wide_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:score_fused -s 3 -c 1 -o gpurun_out/${TAG}_prof_wide $PT > gpurun_out/${TAG}_wide_ncu_full.log 2>&1
timeout 120 ncu -i gpurun_out/${TAG}_prof_wide.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_wide_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_prof_wide.ncu-rep
tail -n 2 gpurun_out/${TAG}_wide_ncu_full.log

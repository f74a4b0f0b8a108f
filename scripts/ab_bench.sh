#!/bin/bash
# usage: ab_bench.sh <variant|base> ...   prints G pairs/s, TFLOP/s, frac, clocks for each
cd "$(dirname "$0")/.."
for v in "$@"; do
  if [ "$v" = base ]; then unset PXR_LIB; else export PXR_LIB=$PWD/pixelrec_multimodal_b200/variants/$v/libpxr.so; fi
  echo -n "$v: "
  timeout 300 python bench.py --steps 8 --warmup 3 --cpu-seconds 0 ${BENCH_ARGS} 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']/1e9,3), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['clocks'])"
done

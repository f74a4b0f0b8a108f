#!/bin/bash
# build a variant of libpxr.so with extra nvcc defines into pixelrec_multimodal_b200/variants/<name>/libpxr.so
name=$1; shift
cd "$(dirname "$0")/../pixelrec_multimodal_b200"
mkdir -p variants/$name
for f in pxr_api simt_kernels score_tc items_tc sampling novelty diversity; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DPXR_PRECISE_MATH -Xcompiler -fPIC -I ../include -I csrc "$@" -c csrc/$f.cu -o variants/$name/$f.o &
done
for a in 1 2 3 4 5; do   # the per-activation / fp16 objects of score_tc.cu (build.py UNITS)
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DPXR_PRECISE_MATH -Xcompiler -fPIC -I ../include -I csrc "$@" -DPXR_TC_TU=$a -c csrc/score_tc.cu -o variants/$name/score_tc_act$a.o &
done
wait
nvcc -shared -o variants/$name/libpxr.so variants/$name/*.o -gencode arch=compute_100a,code=sm_100a && echo built variants/$name

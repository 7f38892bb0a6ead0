#!/bin/bash
# Compile (on the box: nvcc is in the image) and run the two round-2 experiments.   gpurun --timeout 300 -- 'bash tools/gpu_experiments.sh tag'
TAG=${1:-exp}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for e in umma_shifted_window tma_box_rate; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I airpose_b200/csrc experiments/$e.cu -lcuda -o /tmp/$e > $OUT/$e.build.log 2>&1 || { echo "$e: build failed"; tail -5 $OUT/$e.build.log; continue; }
  timeout 120 /tmp/$e > $OUT/$e.log 2>&1; echo "$e exit $?"; cat $OUT/$e.log
done

#!/bin/bash
# round-2 call 1: the two sm_100a experiments + the stock-PyTorch GPU baseline (the 10x denominator)
TAG=${1:-r02e}; OUT=gpurun_out/$TAG; mkdir -p $OUT
bash tools/gpu_experiments.sh $TAG
timeout 600 python tools/torch_gpu_baseline.py 64 10 > $OUT/torch_gpu_baseline.json 2> $OUT/torch_gpu_baseline.err
echo "torch baseline exit $?"; cat $OUT/torch_gpu_baseline.json; tail -3 $OUT/torch_gpu_baseline.err

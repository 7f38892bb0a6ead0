#!/bin/bash
# Trunk schedule sweep (one 128-image forward): chunk size of stage A x one/two streams.   gpurun -- 'bash tools/gpu_sweep.sh tag'
TAG=${1:-sweep}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for cfg in "" "AIRPOSE_TRUNK_CHUNK=128" "AIRPOSE_TRUNK_CHUNK=128 AIRPOSE_TRUNK_ONE_STREAM=1" "AIRPOSE_TRUNK_CHUNK=32" "AIRPOSE_TRUNK_CHUNK=64 AIRPOSE_TRUNK_ONE_STREAM=1" "AIRPOSE_TRUNK_GROUP=64" "AIRPOSE_SK_MINKB=18" "AIRPOSE_SK_MINKB=72"; do
  echo "== $cfg" | tee -a $OUT/sweep.log
  env $cfg timeout 120 python tools/gpu_probe.py --one trunk 2>&1 | grep "n=128" | tee -a $OUT/sweep.log
done

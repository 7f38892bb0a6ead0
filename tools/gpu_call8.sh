#!/bin/bash
TAG=${1:-r02t}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for c in 32 64 128; do echo "== chunk $c"; AIRPOSE_TRUNK_CHUNK=$c timeout 300 python tools/gpu_probe.py trunk 2>&1 | grep "n=128"; done
echo "== chunk 128 group 256 (n=256)"; 
bash tools/gpu_ncu_tail.sh $TAG > /dev/null 2>&1
python tools/region_stalls.py $OUT/prof_tail_source.csv 36000 w_full,slab_full,slab_empty,acc2_full,stg_full,stg_empty,acc3_full0,acc3_full1,acc3_empty0,acc3_empty1,res_full0,res_full1,res_empty0,res_empty1

#!/bin/bash
# ncu --set full of the fused tail kernel inside a 128-image trunk pass
TAG=${1:-r02l}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bneck_tail -s 6 -c 3 -o $OUT/prof_tail python tools/run_once.py trunk 128 2 > $OUT/ncu_tail.log 2>&1
echo "ncu exit $?"
ncu -i $OUT/prof_tail.ncu-rep --page raw --csv > $OUT/prof_tail_raw.csv 2>/dev/null
ncu -i $OUT/prof_tail.ncu-rep --page source --csv > $OUT/prof_tail_source.csv 2>/dev/null
ls -la $OUT

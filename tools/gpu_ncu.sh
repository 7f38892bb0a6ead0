#!/bin/bash
# ncu-only GPU-box pass: launch list of the bench command + full captures of the trunk GEMMs (one whole
# 128-image forward) and the SMPL-X kernels.  CSV pages are extracted on the box; reports larger than
# 24 MiB are dropped so gpurun_out/ stays under its 64 MiB return limit.
#   gpurun --timeout 1200 -- 'bash tools/gpu_ncu.sh [tag]'
TAG=${1:-ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none -k regex:gemm_ -s 77 -c 77 \
    -o $OUT/prof_trunk python tools/run_once.py trunk 128 2 > $OUT/ncu_trunk.log 2>&1
echo "ncu trunk exit $?"
ncu -i $OUT/prof_trunk.ncu-rep --page raw --csv > $OUT/prof_trunk_raw.csv 2>/dev/null
# source-level view of three representative layers: layer1 conv3 (+residual), layer3 conv2 (3x3), layer4 conv2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_ -s 80 -c 3 \
    -o $OUT/prof_trunk_l1 python tools/run_once.py trunk 128 2 > $OUT/ncu_trunk_l1.log 2>&1
ncu -i $OUT/prof_trunk_l1.ncu-rep --page source --csv > $OUT/prof_trunk_l1_source.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smplx_ -s 3 -c 3 \
    -o $OUT/prof_lbs python tools/run_once.py lbs 8192 2 > $OUT/ncu_lbs.log 2>&1
echo "ncu lbs exit $?"
ncu -i $OUT/prof_lbs.ncu-rep --page raw --csv > $OUT/prof_lbs_raw.csv 2>/dev/null
ncu -i $OUT/prof_lbs.ncu-rep --page source --csv > $OUT/prof_lbs_source.csv 2>/dev/null
find $OUT -name "*.ncu-rep" -size +24M -delete
du -sh $OUT; ls -la $OUT

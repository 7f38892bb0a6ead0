#!/bin/bash
# round 2, call F: whole GPU suite + full ncu captures of one 128-image trunk forward and lbs(8192) for profiles/traffic.json
TAG=${1:-r02f2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x -s 2>&1 > $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log; grep -n "gradient vs\|   conv weights\|^   [a-z]" $OUT/pytest_gpu.log | head -14
# one forward = 74 launches (stem pack + stem + 52 conv launches of which 6 fused tails ... ) : skip the first forward (warm-up)
timeout 900 ncu --set full --clock-control none -k regex:'gemm_|bneck_|stem_|avgpool' -s 75 -c 75 -o $OUT/prof_trunk python tools/run_once.py trunk 128 2 > $OUT/ncu_trunk.log 2>&1
echo "ncu trunk exit $?"
ncu -i $OUT/prof_trunk.ncu-rep --page raw --csv > $OUT/prof_trunk_raw.csv 2>/dev/null
python tools/ncu_reduce.py $OUT/prof_trunk_raw.csv $OUT/ncu_full_trunk_128img.csv
timeout 600 ncu --set full --clock-control none -k regex:smplx_ -s 3 -c 3 -o $OUT/prof_lbs python tools/run_once.py lbs 8192 2 > $OUT/ncu_lbs.log 2>&1
ncu -i $OUT/prof_lbs.ncu-rep --page raw --csv > $OUT/prof_lbs_raw.csv 2>/dev/null
python tools/ncu_reduce.py $OUT/prof_lbs_raw.csv $OUT/ncu_full_lbs_b8192.csv
find $OUT -name "*.ncu-rep" -size +20M -delete
find $OUT -name "prof_trunk_raw.csv" -size +30M -delete
ls -la $OUT

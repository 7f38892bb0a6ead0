#!/bin/bash
# round 2, call C: mesh-lane SMPL-X kernel: parity at B >= 256, timing, ncu
TAG=${1:-r02ml}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "smplx or lbs or twoview_full" 2>&1 > $OUT/pytest_smplx.log; tail -15 $OUT/pytest_smplx.log
echo "== lbs"; timeout 300 python tools/gpu_probe.py lbs 2>&1 | tail -8
echo "== lbs old kernel"; AIRPOSE_SMPLX_NO_ML=1 timeout 300 python tools/gpu_probe.py lbs 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smplx_ -s 3 -c 3 \
    -o $OUT/prof_lbs python tools/run_once.py lbs 8192 2 > $OUT/ncu_lbs.log 2>&1
echo "ncu lbs exit $?"
ncu -i $OUT/prof_lbs.ncu-rep --page raw --csv > $OUT/prof_lbs_raw.csv 2>/dev/null
ncu -i $OUT/prof_lbs.ncu-rep --page source --csv > $OUT/prof_lbs_source.csv 2>/dev/null
find $OUT -name "*.ncu-rep" -size +24M -delete

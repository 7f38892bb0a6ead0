#!/bin/bash
# round-2 dev call: fused-kernel parity, trunk parity, trunk timing with / without the fused kernels, launch list of one trunk pass
TAG=${1:-r02f}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q -s > $OUT/pytest_fused.log 2>&1; echo "fused exit $?"; tail -15 $OUT/pytest_fused.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -s -k "trunk or stem or twoview_end_to_end" > $OUT/pytest_trunk.log 2>&1; echo "trunk exit $?"; tail -8 $OUT/pytest_trunk.log
timeout 300 python tools/gpu_probe.py trunk > $OUT/probe_fused.log 2>&1; cat $OUT/probe_fused.log
AIRPOSE_NO_FUSED_TAIL=1 AIRPOSE_NO_FUSED_STEM=1 timeout 300 python tools/gpu_probe.py trunk > $OUT/probe_unfused.log 2>&1; cat $OUT/probe_unfused.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_trunk128.csv python tools/run_once.py trunk 128 2 > $OUT/ncu_launches.log 2>&1
echo "ncu exit $?"
python tools/launch_summary.py $OUT/launches_trunk128.csv | tail -40

#!/bin/bash
TAG=${1:-r02e2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_trunk_batch.py -q -x -s 2>&1 > $OUT/pytest.log; grep -n "passed\|failed\|teacher-forced\|trunk at" $OUT/pytest.log | cut -c1-300
timeout 300 python tools/gpu_probe.py trunk 2>&1 | grep "n="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_trunk128.csv python tools/run_once.py trunk 128 2 > $OUT/ncu_launches.log 2>&1
python tools/launch_summary.py $OUT/launches_trunk128.csv | tail -14

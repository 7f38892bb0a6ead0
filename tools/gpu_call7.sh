#!/bin/bash
# full GPU parity suite + trunk timing + launch list
TAG=${1:-r02o}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_gpu.log
timeout 300 python tools/gpu_probe.py trunk > $OUT/probe_fused.log 2>&1; cat $OUT/probe_fused.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_trunk128.csv python tools/run_once.py trunk 128 2 > $OUT/ncu_launches.log 2>&1
python tools/launch_summary.py $OUT/launches_trunk128.csv | tail -14

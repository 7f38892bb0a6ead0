#!/bin/bash
TAG=${1:-r02h2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trunk_batch.py tests/test_gpu_fused.py -q -x -k "gemm or conv or trunk or stem or bneck" 2>&1 | tail -3
echo "== resident B"; timeout 300 python tools/gpu_probe.py trunk 2>&1 | grep "n="
echo "== streamed B"; AIRPOSE_NO_BRES=1 timeout 300 python tools/gpu_probe.py trunk 2>&1 | grep "n="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_trunk128.csv python tools/run_once.py trunk 128 2 > $OUT/ncu_launches.log 2>&1
python tools/launch_summary.py $OUT/launches_trunk128.csv | tail -12

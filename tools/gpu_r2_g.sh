#!/bin/bash
TAG=${1:-r02g2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_trunk_batch.py tests/test_gpu_dropin.py -q -x 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "autograd_training or training_step or trunk_train or trunk_backward" 2>&1 | tail -3
for cap in 0 74 100; do echo "== stage A grid cap $cap"; AIRPOSE_STAGEA_GRID=$cap timeout 300 python tools/gpu_probe.py trunk 2>&1 | grep "n=128"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_trunk128.csv python tools/run_once.py trunk 128 2 > $OUT/ncu_launches.log 2>&1
python tools/launch_summary.py $OUT/launches_trunk128.csv | grep -E "bneck|stem_pool|total"

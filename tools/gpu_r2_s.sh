#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_fused.py -q -x -s -k "slab" 2>&1 | grep -E "conv3x3 slab|passed|failed|Error|timeout|assert" | head -20
timeout 300 python -m pytest tests/test_gpu_trunk_batch.py -q -x 2>&1 | tail -2
echo "== slab"; timeout 200 python tools/gpu_probe.py trunk 2>&1 | grep "n=128"
echo "== im2col"; AIRPOSE_NO_SLAB_CONV=1 timeout 200 python tools/gpu_probe.py trunk 2>&1 | grep "n=128"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file /tmp/l.csv python tools/run_once.py trunk 128 2 > /dev/null 2>&1; python tools/launch_summary.py /tmp/l.csv | grep -E "slab|gemm_tma|total"

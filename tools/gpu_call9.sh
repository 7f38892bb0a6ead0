#!/bin/bash
TAG=${1:-r02u}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== default"; timeout 300 python tools/gpu_probe.py trunk 2>&1 | grep "n="
echo "== split B"; AIRPOSE_TRUNK_SPLIT_B=1 timeout 300 python tools/gpu_probe.py trunk 2>&1 | grep "n="
echo "== split B chunk 32"; AIRPOSE_TRUNK_SPLIT_B=1 AIRPOSE_TRUNK_CHUNK=32 timeout 300 python tools/gpu_probe.py trunk 2>&1 | grep "n="
AIRPOSE_TRUNK_SPLIT_B=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "trunk or twoview_end_to_end" 2>&1 | tail -3
AIRPOSE_TRUNK_SPLIT_B=1 python tools/diag_determinism.py 70 2>&1 | tail -10

#!/bin/bash
TAG=${1:-r02k}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python tools/diag_determinism.py 70 2>&1 | tail -10
python tools/diag_determinism.py 70 2>&1 | tail -10
timeout 600 python -m pytest tests/test_gpu_fused.py -q 2>&1 | tail -3
timeout 300 python tools/gpu_probe.py trunk 2>&1 | tail -4

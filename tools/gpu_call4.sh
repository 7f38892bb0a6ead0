#!/bin/bash
TAG=${1:-r02i}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python tools/diag_determinism.py 70 2>&1 | tail -12
echo "== ONE_STREAM"; AIRPOSE_TRUNK_ONE_STREAM=1 python tools/diag_determinism.py 70 2>&1 | tail -12

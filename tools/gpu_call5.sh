#!/bin/bash
TAG=${1:-r02j}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== v2, NO_PDL + ONE_STREAM"; AIRPOSE_NO_PDL=1 AIRPOSE_TRUNK_ONE_STREAM=1 python tools/diag_determinism.py 70 2>&1 | tail -10
echo "== v2, unfused"; AIRPOSE_NO_FUSED_TAIL=1 python tools/diag_determinism.py 70 2>&1 | tail -10
cp experiments/bneck_v1.cu.txt airpose_b200/csrc/bneck.cu && python -m airpose_b200.build --force > $OUT/build_v1.log 2>&1; tail -1 $OUT/build_v1.log
echo "== v1"; python tools/diag_determinism.py 70 2>&1 | tail -10

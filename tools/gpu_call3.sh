#!/bin/bash
TAG=${1:-r02h}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_fused.py -q -s > $OUT/pytest_fused.log 2>&1; echo "fused exit $?"; tail -25 $OUT/pytest_fused.log
for v in "" "AIRPOSE_NO_PDL=1" "AIRPOSE_TRUNK_ONE_STREAM=1" "AIRPOSE_NO_FUSED_TAIL=1"; do
  echo "== env: $v"
  env $v timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "trunk_pair_entry" 2>&1 | tail -3
done

#!/bin/bash
# Quick GPU iteration: selected tests + probe sections.  usage: gpu_quick.sh <tag> "<pytest -k expr>" "<probe sections>"
TAG=${1:-q}; KEXPR=${2:-"gemm or conv or trunk"}; SECS=${3:-"trunk twoview"}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q -s -k "$KEXPR" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
grep -E "passed|failed|error|Error|timeout|exit" $OUT/pytest.log | tail -8
timeout 300 python tools/gpu_probe.py $SECS > $OUT/probe.log 2>&1
cat $OUT/probe.log

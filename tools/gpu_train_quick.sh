#!/bin/bash
# Training-path iteration: selected tests, the timed training step, a short bench (JSON contract check).
TAG=${1:-tq}; KEXPR=${2:-"training or autograd or pair or backward"}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 700 python -m pytest tests -m gpu -x -q -s -k "$KEXPR" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
grep -E "passed|failed|error|Error|exit|pair vs|autograd step|parameter gradients|full training" $OUT/pytest.log | tail -12
timeout 300 python tools/train_demo.py --full --pairs 32 --steps 10 --warmup 3 > $OUT/train_full.json 2> $OUT/train_full.err
cat $OUT/train_full.json; tail -3 $OUT/train_full.err
if [ -n "$3" ]; then
  timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
fi

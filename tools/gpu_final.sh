#!/bin/bash
# What the driver runs at round end, in its order: parity tests, smoke(), the reference arm, the bench.   gpurun -- 'bash tools/gpu_final.sh tag'
TAG=${1:-final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref exit $?"; cut -c1-200 $OUT/bench_ref.json
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json; tail -3 $OUT/bench.err

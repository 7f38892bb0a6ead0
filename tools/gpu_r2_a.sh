#!/bin/bash
# round 2, call A: new non-self parity tests, the drop-in tests, the full bench line, chunk-size probe
TAG=${1:-r02x}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_trunk_batch.py tests/test_gpu_dropin.py -q -s -x 2>&1 | tail -60 > $OUT/pytest_new.log; tail -40 $OUT/pytest_new.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; tail -c 6000 $OUT/bench.json; tail -5 $OUT/bench.err
echo "== chunk 64"; timeout 300 python tools/gpu_probe.py trunk 2>&1 | grep "n="
echo "== chunk 128"; AIRPOSE_TRUNK_CHUNK=128 timeout 300 python tools/gpu_probe.py trunk 2>&1 | grep "n="
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>/dev/null; cat $OUT/bench_reference.json

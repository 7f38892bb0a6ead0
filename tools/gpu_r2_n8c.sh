#!/bin/bash
# One N-GPU bench line with every leg (train_step included), CPU baseline skipped.  gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu_r2_n8c.sh tag 8'
TAG=${1:-r02n8c}; N=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --sustain-s 1 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "exit $?"
python - <<PY
import json
d = json.loads(open("$OUT/bench_n$N.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f pairs256 %.0f train %.2f ms (%.0f pairs/s, allreduce alone %.2f ms, identical %s)" % (
    d["value"], d["e2e"]["value"], d["pairs256"]["value"], d["train_step"]["ms_per_step"], d["train_step"]["pairs_per_s"],
    d["train_step"]["allreduce_ms"], d["train_step"]["params_identical_across_ranks"]))
PY
tail -3 $OUT/bench_n$N.err

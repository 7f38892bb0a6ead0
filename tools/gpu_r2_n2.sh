#!/bin/bash
TAG=${1:-r02n2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
N=${2:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "exit $?"; tail -c 5000 $OUT/bench_n$N.json; tail -5 $OUT/bench_n$N.err

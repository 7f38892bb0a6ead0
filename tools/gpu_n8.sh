#!/bin/bash
# N-GPU pass (one box): the bench through torchrun (the driver's launch line) and BASELINE config 4 (whole-network training, 32 pairs per GPU).
#   gpurun --gpus 8 --timeout 400 -- 'bash tools/gpu_n8.sh tag 8'
TAG=${1:-n8}; N=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "bench exit $?"; cat $OUT/bench_n$N.json; tail -2 $OUT/bench_n$N.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tools/train_demo.py --full --pairs 32 --steps 10 --warmup 3 > $OUT/train_n$N.json 2> $OUT/train_n$N.err
echo "train exit $?"; cat $OUT/train_n$N.json; tail -2 $OUT/train_n$N.err

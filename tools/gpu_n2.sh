#!/bin/bash
# Two-GPU pass: the bench through torchrun (the driver's launch line) and the data-parallel whole-network training demo.
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/gpu_n2.sh tag'
TAG=${1:-n2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err
echo "bench exit $?"; cat $OUT/bench_n2.json; tail -3 $OUT/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/train_demo.py --full --pairs 32 --steps 10 --warmup 3 > $OUT/train_n2.json 2> $OUT/train_n2.err
echo "train exit $?"; cat $OUT/train_n2.json; tail -3 $OUT/train_n2.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $OUT/bench_ref_n2.json 2> $OUT/bench_ref_n2.err
echo "ref exit $?"; cat $OUT/bench_ref_n2.json | cut -c1-200

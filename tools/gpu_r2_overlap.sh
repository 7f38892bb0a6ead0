#!/bin/bash
# Overlapped vs single gradient all-reduce of the training step on N GPUs.  gpurun --gpus N --timeout 900 -- 'bash tools/gpu_r2_overlap.sh tag N'
TAG=${1:-r02ov}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -s -k "upper_done" 2>&1 | grep -E "late_split|passed|failed|Error" | tail -4
for rep in 1 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/train_demo.py --full --pairs 32 --steps 12 --warmup 4 2> $OUT/overlap_$rep.err | tail -1 | tee -a $OUT/overlap.jsonl | cut -c1-200
  AIRPOSE_NO_OVERLAP_ALLREDUCE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/train_demo.py --full --pairs 32 --steps 12 --warmup 4 2> $OUT/single_$rep.err | tail -1 | tee -a $OUT/single.jsonl | cut -c1-200
done
tail -3 $OUT/overlap_1.err

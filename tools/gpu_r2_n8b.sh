#!/bin/bash
OUT=gpurun_out/r02n8b; mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 20 --warmup 5 --no-extra-legs --sustain-s 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f e2e %.0f e2e_fp32 %.0f host=%s' % (d['value'], d['e2e']['value'], d['e2e_fp32']['value'], d['config']['host']))"; }
echo "== pinned"; run 29521
echo "== not pinned"; AIRPOSE_BENCH_NO_PIN=1 run 29522
echo "== pinned again"; run 29523
nvidia-smi topo -m 2>/dev/null | head -14; lscpu | grep -E "NUMA|Socket|^CPU\(s\)" 

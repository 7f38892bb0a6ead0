#!/bin/bash
# Per-launch device times of one 128-image trunk forward (ncu, one pass) for the current kernel and,
# with a second argument, for an A/B environment setting.   usage: gpu_layers.sh <tag> ["ENV=1 ENV2=x"]
TAG=${1:-layers}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { # name, env
  env $2 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
     --clock-control none -k regex:gemm_ -s 77 -c 77 --csv --log-file $OUT/$1.csv python tools/run_once.py trunk 128 2 > $OUT/$1.log 2>&1
  echo "$1 exit $?"
}
run new ""
if [ -n "$2" ]; then run alt "$2"; fi
if [ -n "$3" ]; then run alt2 "$3"; fi

#!/bin/bash
# Source-level (stall sampling) captures of single trunk GEMM launches: gpu_src.sh <tag> <skip> [<skip> ...]
# skip = index of the launch among the gemm_ kernels of `run_once.py trunk 128 2` (second forward starts at 77).
TAG=$1; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
for S in "$@"; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_ -s $S -c 1 \
      -o $OUT/k$S python tools/run_once.py trunk 128 2 > $OUT/k$S.log 2>&1
  ncu -i $OUT/k$S.ncu-rep --page source --csv > $OUT/k${S}_source.csv 2>/dev/null
  ncu -i $OUT/k$S.ncu-rep --page raw --csv > $OUT/k${S}_raw.csv 2>/dev/null
  rm -f $OUT/k$S.ncu-rep
  echo "k$S done"
done

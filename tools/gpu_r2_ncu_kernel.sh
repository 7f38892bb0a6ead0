#!/bin/bash
# ncu --set full of ONE kernel of the training step.  gpurun --timeout 600 -- 'bash tools/gpu_r2_ncu_kernel.sh tag kernel_regex [skip]'
TAG=${1:-r02k}; KRE=${2:-smplx_vertex_bwd}; SKIP=${3:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -o $OUT/prof python tools/train_demo.py --full --pairs 32 --steps 2 --warmup 1 > $OUT/ncu.log 2>&1
echo "ncu exit $?"
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2>/dev/null
ncu -i $OUT/prof.ncu-rep --page source --csv > $OUT/prof_source.csv 2>/dev/null
rm -f $OUT/prof.ncu-rep
ls -la $OUT

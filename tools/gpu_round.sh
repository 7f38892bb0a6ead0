#!/bin/bash
# One GPU-box round: parity tests, bench (both arms), ncu launch list of the bench command,
# one `ncu --set full` capture of the dominant kernels.  Outputs land in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?"; cat $OUT/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
cat $OUT/bench_ref.json
timeout 300 python tools/gpu_probe.py gemm lbs trunk ief twoview > $OUT/probe.log 2>&1
cat $OUT/probe.log
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
echo "ncu launches exit $?"
# full capture: trunk GEMM kernels (a spread of layers) and the SMPL-X vertex kernel
timeout 600 ncu --set full --clock-control none -k regex:gemm_ -s 77 -c 77 \
    -o $OUT/prof_trunk python tools/run_once.py trunk 128 2 > $OUT/ncu_trunk.log 2>&1
echo "ncu trunk exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smplx_ -s 3 -c 3 \
    -o $OUT/prof_lbs python tools/run_once.py lbs 8192 2 > $OUT/ncu_lbs.log 2>&1
echo "ncu lbs exit $?"
for f in trunk lbs; do ncu -i $OUT/prof_$f.ncu-rep --page raw --csv > $OUT/prof_${f}_raw.csv 2>/dev/null; done
ncu -i $OUT/prof_lbs.ncu-rep --page source --csv > $OUT/prof_lbs_source.csv 2>/dev/null
# keep gpurun_out under the 64 MiB return limit
find $OUT -name "*.ncu-rep" -size +24M -delete
ls -la $OUT

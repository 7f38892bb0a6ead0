#!/bin/bash
# final N=1 pass: full GPU suite, bench (ours + reference arm), ncu launch list of the bench, full captures for traffic.json
TAG=${1:-r02z1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > $OUT/pytest_gpu_tail.txt; cat $OUT/pytest_gpu_tail.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; tail -c 1500 $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>/dev/null; cut -c1-300 $OUT/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs --sustain-s 0 > $OUT/bench_under_ncu.log 2>&1
python tools/launch_summary.py $OUT/launches_bench.csv > $OUT/launches_bench_summary.txt; tail -22 $OUT/launches_bench_summary.txt
timeout 900 ncu --set full --clock-control none -k regex:'gemm_|bneck_|stem_|avgpool' -s 75 -c 75 -o $OUT/prof_trunk python tools/run_once.py trunk 128 2 > $OUT/ncu_trunk.log 2>&1
ncu -i $OUT/prof_trunk.ncu-rep --page raw --csv > $OUT/prof_trunk_raw.csv 2>/dev/null
python tools/ncu_reduce.py $OUT/prof_trunk_raw.csv $OUT/ncu_full_trunk_128img.csv
timeout 600 ncu --set full --clock-control none -k regex:smplx_ -s 3 -c 3 -o $OUT/prof_lbs python tools/run_once.py lbs 8192 2 > $OUT/ncu_lbs.log 2>&1
ncu -i $OUT/prof_lbs.ncu-rep --page raw --csv > $OUT/prof_lbs_raw.csv 2>/dev/null
python tools/ncu_reduce.py $OUT/prof_lbs_raw.csv $OUT/ncu_full_lbs_b8192.csv
find $OUT -name "*.ncu-rep" -delete

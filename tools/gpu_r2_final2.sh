#!/bin/bash
# Final validation of the round: GPU suite, smoke, bench (both arms), ncu launch list of the bench command, train-step launch list.
TAG=${1:-r02z6}; OUT=gpurun_out; mkdir -p $OUT/$TAG
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > $OUT/${TAG}_pytest_gpu_tail.txt
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1; echo smoke=$? >> $OUT/${TAG}_smoke.txt
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/$TAG/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline --sustain-s 0 > $OUT/$TAG/bench_under_ncu.log 2>&1
python tools/launch_summary.py $OUT/$TAG/launches_bench.csv > $OUT/${TAG}_launches_bench_summary.txt 2>&1
rm -f $OUT/$TAG/launches_bench.csv
bash tools/gpu_r2_trainlist.sh ${TAG}_train > /dev/null 2>&1
tail -2 $OUT/${TAG}_pytest_gpu_tail.txt; tail -2 $OUT/${TAG}_smoke.txt
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench.json"))
print(d["value"], d["stage_ms"], "e2e", d["e2e"]["value"], "sust", d["sustained"]["value"], "p256", d["pairs256"]["value"], "train", d["train_step"]["ms_per_step"],
      "lbs", d["roofline_lbs"]["ms"], "x", d["torch_gpu_baseline"]["ours_over_best_torch_gpu"], "frac", d["roofline"]["frac"], d["clocks"])
PY
head -3 $OUT/${TAG}_train/train_step_launches.txt; head -c 300 $OUT/${TAG}_train/train_full.json; tail -3 $OUT/${TAG}_launches_bench_summary.txt

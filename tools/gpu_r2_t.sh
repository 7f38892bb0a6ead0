#!/bin/bash
TAG=${1:-r02t2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_train.csv python tools/train_demo.py --full --steps 1 --warmup 1 > $OUT/ncu.log 2>&1
python - <<'PY'
import csv,collections,sys
f=sys.argv[1] if len(sys.argv)>1 else None
PY
python tools/launch_summary.py $OUT/launches_train.csv | sort -k3 -n -r -t$'\t' | tail -45

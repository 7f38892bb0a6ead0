#!/bin/bash
# ncu launch list of ONE whole training step (32 pairs), reduced by kernel name.  gpurun --timeout 600 -- 'bash tools/gpu_r2_trainlist.sh tag'
TAG=${1:-trainlist}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 200 python tools/train_demo.py --full --pairs 32 --steps 10 --warmup 3 > $OUT/train_full.json 2> $OUT/train_full.err
cat $OUT/train_full.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv \
    --log-file $OUT/launches_train.csv python tools/train_demo.py --full --pairs 32 --steps 1 --warmup 1 > $OUT/train_under_ncu.log 2>&1
echo "ncu train exit $?"
python - > $OUT/train_step_launches.txt <<PY
import csv, collections
rows = [l for l in open("$OUT/launches_train.csv") if not l.startswith("==")]
agg = collections.Counter(); cnt = collections.Counter()
recs = list(csv.DictReader(rows))
adam = [i for i, r in enumerate(recs) if "adam_kernel" in r["Kernel Name"]]
recs = recs[adam[-2] + 1: adam[-1] + 1] if len(adam) >= 2 else recs      # the last whole step
for r in recs:
    k = r["Kernel Name"].split("(")[0][:70]
    agg[k] += float(r["Metric Value"].replace(",", "")) / 1e3; cnt[k] += 1
tot = sum(agg.values())
print("total us %.0f over %d launches" % (tot, sum(cnt.values())))
for k, v in agg.most_common(60):
    print("%8.1f us %5.1f%% %5d  %s" % (v, 100 * v / tot, cnt[k], k))
PY
cat $OUT/train_step_launches.txt
rm -f $OUT/launches_train.csv.tmp

"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: python tools/launch_summary.py file.csv"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = r["Kernel Name"][:70]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    v = float(r["Metric Value"].replace(",", ""))
    a[1] += v / 1e3 if r.get("Metric Unit", "ns") in ("ns", "nsecond") else v
tot = sum(t for _, t in agg.values())
for k, (n, t) in agg.items():
    print("%-72s %5d %10.1f us  %8.1f/launch" % (k, n, t, t / n))
print("total %.1f us over %d launches" % (tot, sum(n for n, _ in agg.values())))

"""Top stall sites of an `ncu --page source --csv` dump: python tools/stall_top.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print(rows[0][1][:100])
hdr = rows[1]
ci, cs = hdr.index('# Samples'), hdr.index('Source')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
body = [r for r in rows[2:] if len(r) > ci and r[ci].isdigit()]
tot = sum(int(r[ci]) for r in body)
print('total samples', tot)
agg = {}
for r in body:
    for i in stall:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for idx, r in sorted(enumerate(body), key=lambda t: -int(t[1][ci]))[:N]:
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall), reverse=True)[:2]
    print('%5d %5s  %-90s %s' % (idx, r[ci], r[cs][:90], st))

"""Group the warp-stall samples of an `ncu --page source --csv` (SASS) dump into regions delimited by mbarrier waits.
    python tools/region_stalls.py file.csv <bar_base_hex> name0,name1,...   (barrier names in 8-byte order from bar_base)"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
base = int(sys.argv[2], 16)
names = sys.argv[3].split(",")
hdr = rows[1]
ci, cs = hdr.index('# Samples'), hdr.index('Source')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
body = [r for r in rows[2:] if len(r) > ci and r[ci].isdigit()]
tot = sum(int(r[ci]) for r in body)
cur, regions, order = 'prologue', {}, []
for r in body:
    m = re.search(r'TRYWAIT.*\+0x([0-9a-f]+)\]', r[cs])
    if m:
        off = int(m.group(1), 16) - base
        if 0 <= off < 8 * len(names):
            cur = 'after ' + names[off // 8] + ' @' + r[0][-5:]
    if cur not in regions:
        regions[cur] = [0, {}]; order.append(cur)
    regions[cur][0] += int(r[ci])
    for i in stall:
        v = int(r[i] or 0)
        if v: regions[cur][1][hdr[i][6:]] = regions[cur][1].get(hdr[i][6:], 0) + v
print('total samples', tot)
for k in order:
    n, st = regions[k]
    top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
    print('%-34s %6d %5.1f%%  %s' % (k, n, 100.0 * n / tot, top))

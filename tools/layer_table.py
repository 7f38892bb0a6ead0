"""Per-layer table from `tools/gpu_layers.sh` output: python tools/layer_table.py gpurun_out/<tag>/new.csv [alt.csv]"""
import csv, sys, collections

SEQ_A = ['stem', 'l1b0c1', 'l1b0c2', 'l1b0ds', 'l1b0c3'] + ['l1b%dc%d' % (b, c) for b in (1, 2) for c in (1, 2, 3)] + \
        ['l2b0c1', 'l2b0c2', 'l2b0ds', 'l2b0c3'] + ['l2b%dc%d' % (b, c) for b in (1, 2, 3) for c in (1, 2, 3)]
SEQ_B = ['l3b0c1', 'l3b0c2', 'l3b0ds', 'l3b0c3'] + ['l3b%dc%d' % (b, c) for b in range(1, 6) for c in (1, 2, 3)] + \
        ['l4b0c1', 'l4b0c2', 'l4b0ds', 'l4b0c3'] + ['l4b%dc%d' % (b, c) for b in (1, 2) for c in (1, 2, 3)]
LABELS = ['A:' + s for s in SEQ_A] * 2 + ['B:' + s for s in SEQ_B]


def load(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        per.setdefault(int(r['ID']), {})[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
    ids = sorted(per)
    return [per[i] for i in ids]


def main():
    tabs = [load(p) for p in sys.argv[1:]]
    n = len(tabs[0])
    agg = collections.OrderedDict()
    for i in range(n):
        lab = LABELS[i] if n == len(LABELS) else str(i)
        key = lab[:-2] + 'bN' + lab[-2:] if lab[4:5] == 'b' and lab[5] != '0' else lab     # merge identical blocks 1..N
        e = agg.setdefault(key, [0] + [0.0] * (3 * len(tabs)))
        e[0] += 1
        for t, tab in enumerate(tabs):
            m = tab[i]
            e[1 + 3 * t] += m['gpu__time_duration.sum'] / 1e3
            e[2 + 3 * t] += (m.get('dram__bytes_read.sum', 0) + m.get('dram__bytes_write.sum', 0)) / 1e6
            e[3 + 3 * t] += m.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0)
    tot = [0.0] * len(tabs)
    print('%-12s %3s' % ('layer', 'n') + ''.join('  %9s %9s %6s' % ('us/launch', 'dramMB', 'tens%') for _ in tabs))
    for k, e in agg.items():
        s = '%-12s %3d' % (k, e[0])
        for t in range(len(tabs)):
            s += '  %9.1f %9.1f %6.1f' % (e[1 + 3 * t] / e[0], e[2 + 3 * t] / e[0], e[3 + 3 * t] / e[0])
            tot[t] += e[1 + 3 * t]
        print(s)
    print('total us: ' + '  '.join('%.1f' % x for x in tot))


main()

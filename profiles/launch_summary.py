"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean time, share."""
import csv
import sys
from collections import OrderedDict


def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    kn, mv, mn = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
    mu = hdr.index('Metric Unit')
    agg = OrderedDict()
    for r in rows[1:]:
        if r[mn] != 'gpu__time_duration.sum':
            continue
        v = float(r[mv].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[mu], 1.0)
        name = r[kn].split('(')[0][:62]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-62s n=%3d avg=%8.1f us share=%.3f' % (name, n, t / n, t / total))


if __name__ == '__main__':
    main(sys.argv[1])

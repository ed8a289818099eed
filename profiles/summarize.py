"""Summarise ncu CSV exports kept under profiles/ (launch lists and raw pages)."""
import csv
import io
import sys
from collections import defaultdict


def launches(path):
    txt = open(path).read()
    rows = csv.DictReader(io.StringIO(txt[txt.index('"ID"'):]))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r['Metric Name'] != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r['Metric Unit'], 1e-6)
        k = r['Kernel Name'].split('(')[0]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print('| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| %s | %d | %.3f | %.3f | %.1f%% |' % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))


if __name__ == '__main__':
    launches(sys.argv[1])

#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg, tot, n = collections.OrderedDict(), 0.0, 0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3}.get(row['Metric Unit'], v)
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        name = re.sub(r'^void ', '', name)
        k = (name, row.get('Grid Size'), row.get('Block Size'))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += v; tot += v; n += 1
    print('# ncu launch list summary: %s' % path)
    print()
    print('%d launches, %.1f us total (cold-cache, serialised: compare SHARES, not absolutes)' % (n, tot))
    print()
    print('| share | total us | launches | us / launch | kernel | grid | block |')
    print('|---:|---:|---:|---:|---|---|---|')
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print('| %.1f%% | %.1f | %d | %.2f | `%s` | %s | %s |' % (100 * t / tot, t, c, t / c, k[0][:120], k[1], k[2]))


if __name__ == '__main__':
    main(sys.argv[1])

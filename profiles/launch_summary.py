#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (cold-cache, serialised: compare SHARES)."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = row['Kernel Name'].split('(')[0].replace('void ', '')
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        u = row['Metric Unit']
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else (v * 1e6 if u == 's' else v))
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print('%-44s %6s %12s %10s %7s' % ('kernel', 'n', 'total_us', 'avg_us', 'share'))
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print('%-44s %6d %12.1f %10.1f %7.3f' % (k[:44], v[0], v[1], v[1] / v[0], v[1] / tot))


if __name__ == '__main__':
    main(sys.argv[1])

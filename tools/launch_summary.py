"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1e6 if unit == 'ns' else (v / 1e3 if unit == 'us' else v)
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        name = re.sub(r'^void ', '', name).replace('<unnamed>::', '')
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print('total %.3f ms over %d launches' % (tot, sum(c for c, _ in agg.values())))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-64s %5d %9.3f ms %5.1f%%' % (k[:64], c, t, 100 * t / tot))


if __name__ == '__main__':
    main(sys.argv[1])

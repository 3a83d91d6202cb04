"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v *= {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3}.get(unit, 1.0)
        tot[row['Kernel Name']] += v
        cnt[row['Kernel Name']] += 1
    total = sum(tot.values())
    print('%10s %6s %5s %10s  kernel' % ('sum [us]', 'share', 'n', 'avg [us]'))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print('%10.1f %5.1f%% %5d %10.1f  %s' % (v, 100 * v / total, cnt[k], v / cnt[k], k[:100]))
    print('%10.1f  total' % total)


if __name__ == '__main__':
    main(sys.argv[1])

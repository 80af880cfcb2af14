"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean time, share."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[ui], 1e-3)
        name = r[ki].split('(')[0][:70]
        agg[name][0] += 1
        agg[name][1] += float(r[vi].replace(',', '')) * scale
    tot = sum(v[1] for v in agg.values())
    print("%-72s %5s %10s %7s" % ("kernel", "n", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-72s %5d %10.1f %6.1f%%" % (k, v[0], v[1] / v[0], 100 * v[1] / tot))
    print("total us: %.1f over %d launches" % (tot, sum(v[0] for v in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1])

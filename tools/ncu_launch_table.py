"""Per-kernel table (steady-state launches) from tools/ncu_launches.sh output."""
import collections
import csv
import sys


def main(path, last=3):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, vi = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split('(')[0].replace('void ', '').replace('fmhr::', '')[:40]
        agg.setdefault(name, collections.defaultdict(list))[r[mi]].append(float(r[vi].replace(',', '')))
    print("%-40s %3s %8s %11s %11s %9s %8s %8s" % ("kernel", "n", "us", "warp_inst", "lsu_wavef", "wf/clk/SM", "rd_MB", "wr_MB"))
    tot = 0.0
    for k, m in agg.items():
        f = lambda key: sum(m[key][-last:]) / max(1, len(m[key][-last:]))
        t = f('gpu__time_duration.sum') / 1000
        tot += t
        print("%-40s %3d %8.1f %11.0f %11.0f %9.2f %8.1f %8.1f" % (
            k, len(m['gpu__time_duration.sum']), t, f('smsp__inst_executed.sum'), f('l1tex__data_pipe_lsu_wavefronts.sum'),
            f('l1tex__data_pipe_lsu_wavefronts.sum') / f('sm__cycles_elapsed.max') / 148,
            f('dram__bytes_read.sum') / 1e6, f('dram__bytes_write.sum') / 1e6))
    print("sum of kernel times per iteration: %.1f us" % tot)


if __name__ == "__main__":
    main(sys.argv[1])

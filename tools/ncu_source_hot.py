"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name ...`:
hottest CUDA source lines by executed warp instructions and by stall samples."""
import csv
import sys


def main(path, top=25):
    out = []
    fname, hdr = None, None
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
            continue
        ie, ss, te = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        try:
            n, smp, tn = float(r[ie] or 0), float(r[ss] or 0), float(r[te] or 0)
        except ValueError:
            continue
        if n > 0 or smp > 0:
            out.append((n, smp, tn, fname, r[0], r[1].strip()[:105]))
    tot, tots = sum(o[0] for o in out), sum(o[1] for o in out)
    print("total warp instructions %.4g, samples %.0f, avg lanes %.1f" % (tot, tots, sum(o[2] for o in out) / max(tot, 1)))
    for title, key in (("by instructions", lambda o: -o[0]), ("by stall samples", lambda o: -o[1])):
        print("--- " + title)
        for n, smp, tn, f, ln, src in sorted(out, key=key)[:top]:
            print("%5.1f%% inst %5.1f%% smp %4.1f lanes  %s:%s  %s" % (100 * n / tot, 100 * smp / max(tots, 1), tn / max(n, 1), f, ln, src))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)

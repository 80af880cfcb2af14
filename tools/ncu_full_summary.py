"""Per-kernel extract of `ncu -i X.ncu-rep --page raw --csv` for a --set full capture (profiles/r1_final_ncu_full_summary.txt)."""
import csv
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("l1tex__t_sector_hit_rate.pct", "l1_hit"), ("lts__t_sector_hit_rate.pct", "l2_hit"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1_wavefront_pct"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"), ("smsp__inst_executed.sum", "warp_inst"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes"),
        ("lts__t_sectors_srcunit_tex_op_red.sum", "l2_red_sectors"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    print("ncu --set full --clock-control none --import-source on, one steady-state launch of every kernel of the fused")
    print("iteration (bench.py --steps 3 --warmup 3 --no-graphs, workload interhand_48x512x334; default cache control = cold")
    print("L2, so the times are above the warm-cache launch table).  tools/ncu_full.sh + tools/ncu_full_summary.py.")
    for r in rows[2:]:
        print("== " + r[ki].split('(')[0].replace("void ", "").replace("fmhr::", ""))
        for w, short in WANT:
            if w in hdr:
                v = r[hdr.index(w)]
                try:
                    v = "%.3f" % float(v)
                except ValueError:
                    pass
                print("   %-18s %s %s" % (short, v, units[hdr.index(w)]))
        st = sorted([(float(r[hdr.index(h)] or 0),
                      h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for h in stall], reverse=True)[:5]
        print("   stalls/issue       " + ", ".join("%s=%.2f" % (b, a) for a, b in st))


if __name__ == "__main__":
    main(sys.argv[1])

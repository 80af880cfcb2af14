"""Per-kernel summary of an `ncu --set full` report exported with `ncu -i X.ncu-rep --page raw --csv`."""
import csv
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit"), ("lts__t_sector_hit_rate.pct", "l2_hit"),
        ("smsp__inst_executed.sum", "warp_inst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
        ("lts__t_sectors_srcunit_tex_op_red.sum", "l2_red_sectors"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    seen = set()
    for r in rows[2:]:
        name = r[ki].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        print("== " + name)
        for m, short in WANT:
            if m in hdr:
                i = hdr.index(m)
                print("   %-16s %s %s" % (short, r[i], units[i]))


if __name__ == "__main__":
    main(sys.argv[1])

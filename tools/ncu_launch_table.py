"""Per-kernel table (steady-state launches) from tools/ncu_launches.sh output."""
import collections
import csv
import json
import sys

STAGE_OF = {"ham_vertex_prep_kernel": "vertex_normals", "ham_normals_kernel": "vertex_normals",
            "ham_regulariser_kernel": "vertex_normals", "ham_trirec_kernel": "vertex_normals",
            "ham_coverage_meshlet_kernel": "coverage", "ham_scan_kernel": "shade", "ham_shade_kernel": "shade",
            "ham_aa_loss_kernel": "antialias_loss", "ham_pair_bwd_kernel": "pixel_backward",
            "ham_pixel_bwd_kernel": "pixel_backward", "ham_finalize_scalars_kernel": "pixel_backward",
            "ham_normal_grad_kernel": "update_adam", "ham_update_pass2_kernel": "update_adam"}


def main(path, last=3, traffic_out=None):
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
    if traffic_out:
        per_stage = collections.defaultdict(float)
        total = 0.0
        for k, m in agg.items():
            base = k.split('<')[0]
            if base not in STAGE_OF:
                continue
            f = lambda key: sum(m[key][-last:]) / max(1, len(m[key][-last:]))
            b = f('dram__bytes_read.sum') + f('dram__bytes_write.sum')
            per_stage[STAGE_OF[base]] += b
            total += b
        json.dump({"iteration_bytes": total, "kernels": dict(per_stage), "source": path.split('/')[-1],
                   "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --cache-control none "
                           "--clock-control none, steady-state launches of bench.py --no-graphs"}, open(traffic_out, "w"), indent=1)
        print("wrote", traffic_out, "iteration DRAM bytes: %.1f MB" % (total / 1e6))


if __name__ == "__main__":
    main(sys.argv[1], traffic_out=sys.argv[2] if len(sys.argv) > 2 else None)

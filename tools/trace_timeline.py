#!/usr/bin/env python
"""Real schedule of one CUDA-graph replay of the fused iteration, from the %globaltimer stamps of a -DFMHR_TRACE build:

    tools/build_variant.sh trace -DFMHR_TRACE
    FMHR_B200_LIB=variants/libfmhr_trace.so python tools/trace_timeline.py [--workload W] [--views N] [--reps 30] [--out f.json]

Per kernel: start / end offset (us) from the first kernel's first block, averaged over the repetitions, plus the gap to
the previous kernel on the critical chain.  ncu launch lists are serialised and cold-cache; this is the live overlap.
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
NAMES = {0: "vertex_prep", 1: "normals", 2: "trirec", 3: "regulariser", 4: "coverage", 5: "scan", 6: "shade", 7: "aa_loss",
         8: "pair_bwd", 9: "pixel_bwd", 10: "finalize", 11: "normal_grad", 12: "update_adam", 13: "peer_reduce_scatter",
         14: "peer_normal_grad", 15: "peer_reduce_normal_grad", 16: "u8_to_f32", 17: "update_sh", 18: "reg_grad",
         19: "shade_bwd", 20: "rendezvous_a(first post..last seen)", 21: "rendezvous_a first seen",
         22: "rendezvous_b(first post..last seen)", 23: "rendezvous_b first seen"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="interhand_48x512x334")
    ap.add_argument("--views", type=int, default=None)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-graphs", action="store_true")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from fmhr_b200 import _lib, synth
    from fmhr_b200.ham import HamOptimizer
    from fmhr_b200.render import render_views
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    scene = synth.build_scene(a.workload, lambda *x: render_views(*x, device=dev), n_views=a.views, camera_seed=1 + rank)
    c = lambda k, dt=torch.float32: torch.tensor(scene[k], dtype=dt, device=dev)
    opt = HamOptimizer(c("vertices"), c("faces", torch.int32), c("imgs"), c("masks"), c("valid_masks"), c("w2cs"),
                       c("projs"), c("sh_coeffs"), c("albedo"), scene["conf"], use_graphs=not a.no_graphs,
                       exchange=os.environ.get("FMHR_EXCHANGE", "peer") if world > 1 else None)
    n = scene["imgs"].shape[0]
    views = torch.arange(n, dtype=torch.int32, device=dev)
    for _ in range(20):
        opt.step_phase_b(views)
    reps = min(a.reps, 64)
    # free-running loop: [mark][step][mark][step]... with no host synchronisation in between (the ranks of a multi-GPU job
    # stay coupled through the exchange only, as in the benchmark); mark i stores the stamps of step i - 1
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    _lib.check(lib.fmhr_trace_mark(-1, _lib.stream()), "trace_mark")
    for _ in range(10):
        opt.step_phase_b(views)
    for r in range(reps):
        _lib.check(lib.fmhr_trace_mark(r - 1, _lib.stream()), "trace_mark")
        opt.step_phase_b(views)
    _lib.check(lib.fmhr_trace_mark(reps - 1, _lib.stream()), "trace_mark")
    buf = (ctypes.c_ulonglong * (reps * 128))()
    _lib.check(lib.fmhr_trace_read(buf, -reps, 0), "trace_read")
    allrows = np.array(list(buf), dtype=np.uint64).reshape(reps, 64, 2).astype(np.int64)
    rows = [allrows[r] for r in range(reps)]
    rows = np.stack(rows)  # [reps, 64, 2]
    for k in (21, 23):  # single-stamp slots ("first block that has seen every peer")
        rows[:, k, 1] = np.where(rows[:, k, 0] >= 0, rows[:, k, 0], 0)
    used = [k for k in range(64) if (rows[:, k, 1] > 0).all() and (rows[:, k, 0] >= 0).all() and (rows[:, k, 0] != -1).all()]
    period = float(np.median(np.diff(rows[:, used, 0].min(axis=1)))) / 1e3 if reps > 1 else 0.0
    t0 = rows[:, used, 0].min(axis=1, keepdims=True)
    out = {}
    used = [k for k in used if k < 32]
    for j, k in enumerate(used):
        s = (rows[:, k, 0] - t0[:, 0]) / 1e3
        e = (rows[:, k, 1] - t0[:, 0]) / 1e3
        out[NAMES.get(k, str(k))] = {"start_us": float(np.median(s)), "end_us": float(np.median(e)),
                                      "dur_us": float(np.median(e - s))}
        if k < 20 and (rows[:, k + 32, 1] > 0).all():  # kernel scopes: spread of the blocks' entries and exits
            le = (rows[:, k + 32, 1] - t0[:, 0]) / 1e3
            fx = (rows[:, k + 32, 0] - t0[:, 0]) / 1e3
            out[NAMES.get(k, str(k))].update({"last_entry_us": float(np.median(le)), "first_exit_us": float(np.median(fx))})
    total = float(np.median((rows[:, used, 1].max(axis=1) - t0[:, 0]) / 1e3))
    import time
    time.sleep(0.2 * rank)  # keep the ranks' tables apart
    if rank == 0 or world > 1:
        print("rank %d: %s, %d views, graph=%s: first block entry -> last block exit = %.1f us, step period %.1f us "
              "(median of %d free-running steps)" % (rank, a.workload, n, not a.no_graphs, total, period, reps))
        for name, v in sorted(out.items(), key=lambda kv: kv[1]["start_us"]):
            extra = ("   last block entry %8.1f  first block exit %8.1f" % (v["last_entry_us"], v["first_exit_us"])
                     if "last_entry_us" in v else "")
            print("  %-40s start %8.1f  end %8.1f  dur %7.1f%s" % (name, v["start_us"], v["end_us"], v["dur_us"], extra))
    if a.out and rank == 0:
        json.dump({"workload": a.workload, "views": n, "total_us": total, "kernels": out}, open(a.out, "w"), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

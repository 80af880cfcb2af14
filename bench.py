#!/usr/bin/env python
"""bench.py — HAM iterations/sec (fwd + bwd + Adam) on the BASELINE.json workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one phase-B HAM iteration (mesh_sfs_optim.py:253-310) over ALL views of the workload in one batch
(SURVEY.md 8d).  Default workload = BASELINE.json configs[1]: InterHand-shaped 48 views x 512x334, 3x-subdivided
single hand (49,281 verts / 98,432 faces), conf/ih_sfs.conf weights.  Multi-GPU is weak scaling: every rank holds the
workload's view set (its own seeded camera ring), the global batch is N x 48 views, vertex / albedo gradients are
summed with one NCCL all-reduce per iteration and `value` counts 48-view iteration equivalents per second.

Prints ONE JSON line (see the contract in the task description): value (device-resident inputs, CUDA-event timed),
e2e (host buffers, H2D + D2H inside the timed region), roofline (dominant kernel, measured live), cpu_baseline
(the oracle = restated reference loop on the host cores), clocks, gpu_launches.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "HAM iters/sec (fwd+bwd, views x res)"
UNIT = "iters/s"
KERNELS_PER_STEP = 13  # prep, coverage, normals, trirec, regulariser, reg_grad, scan, shade+bwd, antialias_loss, pair_bwd, finalize, normal_grad, adam


def b_alg_bytes(n, H, W, V, F, E):
    """Algorithmic bytes of one iteration, SURVEY.md 8(d) / BASELINE.md section 3."""
    P = n * H * W
    return 68 * P + 204 * V + 12 * F + 8 * (2 * E + V) + 4 * (V + 1) + 164 * n


# Algorithmic HBM bytes per launch of each stage of the fused iteration (DESIGN.md "Kernels"), dense model of SURVEY.md
# 8(d): per pixel P, per vertex V, per view n.  (The kernels only touch covered pixels, so their DRAM traffic is lower.)
def stage_bytes(stage, n, H, W, V, F):
    P = n * H * W
    return {
        "clears": 0,                                       # no clear pass: the scan kernel resets dirty tiles
        "vertex_normals": 36 * V + 12 * F + 24 * V + 160 * F,   # vertices, normals, per-triangle records
        "clip_transform": 0,                               # fused into the coverage kernel
        "coverage": 16 * n * V + 12 * F + 8 * P,           # vertices per view + faces + z-buffer keys
        "shade": 8 * P + 4 * P + 16 * P,                   # read z-buffer + mask, write colour plane
        "antialias_loss": 8 * P + 16 * P + 12 * P + 4 * P + 16 * P,  # zbuf, colour, img, valid_mask, write pixel grads
        "pixel_backward": 8 * P + 16 * P + 16 * P,         # zbuf, colour, pixel grads (vertex atomics hit L2)
        "update_adam": 204 * V + 12 * F,
    }[stage]


def measured_traffic():
    """DRAM bytes per iteration (dram__bytes_read.sum + dram__bytes_write.sum over the iteration's kernels) from the
    committed ncu capture of this command, profiles/r2/traffic.json (written by tools/ncu_launch_table.py)."""
    p = os.path.join(ROOT, "profiles", "r2", "traffic.json")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                clk, mxc = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            mx = mxc
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for nme, val in zip(names, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nme)
        if not sm:
            sm = [float(l.split(",")[0]) for _, l in self.rows[-3:] if l.split(",")[0].strip().replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def quantise_targets(scene):
    """Target images in their native 8-bit form (get_data.py:77-90 reads PNGs and divides by 255): every leg of the
    bench (device-resident, e2e, CPU reference) optimises against img = u8 / 255; the e2e leg uploads the u8 bytes."""
    img_u8 = np.clip(np.rint(np.asarray(scene["imgs"], dtype=np.float64) * 255.0), 0, 255).astype(np.uint8)
    msk_u8 = (np.asarray(scene["masks"]) > 0).astype(np.uint8) * 255
    scene["imgs"] = img_u8.astype(np.float32) / np.float32(255.0)
    scene["masks"] = (msk_u8 > 127).astype(np.float32)
    return img_u8, msk_u8


def cpu_reference_step_rate(scene, steps, warmup, threads):
    """Times the oracle (restated reference loop, PyTorch-CPU + C++ reference rasteriser) on the host cores."""
    import torch
    from oracle import ham as oham
    torch.set_num_threads(threads)
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    st = oham.HamState(scene)
    views = list(range(scene["imgs"].shape[0]))
    for _ in range(warmup):
        oham.phase_b_step(st, views)
    t = time.time()
    for _ in range(steps):
        oham.phase_b_step(st, views)
    return steps / (time.time() - t)


_JSON_FD = None


def minibatch_epochs(opt, n, bsz, n_ep, dev, barrier):
    """SURVEY.md 8(d), secondary metric: the reference's own loop shape (mesh_sfs_optim.py:248-260) - every epoch draws a
    permutation of the views and steps through it in batches of conf `batch` views (conf/ih_sfs.conf:32: 32 -> 32 + 16 for
    48 views), a NEW index tensor per step.  Epochs per second, CUDA events around n_ep epochs."""
    import numpy as np
    import torch
    # one workspace layout for every batch size of the epoch (fmhr_ham_config.n_views_capacity): the short last batch does
    # not force a z-buffer reset before and after it
    opt.set_batch_capacity(n)

    def run_epoch():
        perm = torch.randperm(n, device=dev).to(torch.int32)
        for k in range(0, n, bsz):
            opt.step_phase_b(perm[k:k + bsz])

    for _ in range(3):
        run_epoch()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    m0.record()
    for _ in range(n_ep):
        run_epoch()
    m1.record()
    barrier()
    ms_ep = m0.elapsed_time(m1) / n_ep
    if not all(np.isfinite(opt.losses.cpu().tolist())):
        return {"error": "non-finite optimisation state after the mini-batch leg"}
    return {"value": 1000.0 / ms_ep, "unit": "epochs/s", "batch": bsz, "views": n,
            "steps_per_epoch": (n + bsz - 1) // bsz, "ms_per_epoch": ms_ep, "epochs_timed": n_ep,
            "note": "torch.randperm per epoch, one optimiser step per batch of views (mesh_sfs_optim.py:248-310); "
                    "workspace laid out for the largest batch (set_batch_capacity): no z-buffer reset between batch sizes"}


def emit(obj):
    """The one JSON line of the run, on the process's ORIGINAL stdout."""
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, line)


def run_reference(args, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path = the oracle port (nvdiffrast has no
    CPU build and is not vendored, so rasterize/interpolate/antialias are the plain C++ reference rasteriser)."""
    if rank != 0:
        return
    warnings.filterwarnings("ignore")
    from fmhr_b200 import synth
    from oracle import ham as oham
    from oracle import raster as oraster
    oraster.build()
    threads = os.cpu_count() or 1
    wl = dict(synth.WORKLOADS[args.workload])
    scene = synth.build_scene(wl, oham.render_views)
    quantise_targets(scene)
    steps, warm = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    rate = cpu_reference_step_rate(scene, steps, warm, threads)
    sample = "full workload, %d timed iteration(s) after %d warm-up" % (steps, warm)
    emit({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1000.0 / rate, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # same keys as the CUDA arm's config; the CPU arm times ONE 48-view batch (value is in 48-view iterations per second
        # in both arms: at N > 1 the CUDA arm's weak-scaling value counts N such batches per step)
        "config": {"workload": args.workload, "views_per_gpu": wl["n"], "global_views": wl["n"] * max(1, world), "H": wl["H"],
                   "W": wl["W"], "verts": int(scene["vertices"].shape[0]), "faces": int(scene["faces"].shape[0]),
                   "phase": "B (delta+albedo, conf/ih_sfs.conf weights)", "launch": "CPU, %d threads" % threads},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="interhand_48x512x334")
    ap.add_argument("--views", type=int, default=None,
                    help="keep only the first N views of the workload (e.g. 16 = one rank's shard of stress_128x2048x2048)")
    ap.add_argument("--ncc", action="store_true",
                    help="add the NCC photo-consistency term (BASELINE.json configs[2]): 50,000 surface points x 11x11 patches, "
                         "reference view 0 against the other views (fmhr_b200.ncc_term; single GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--minibatch", type=int, default=32, help="views per step of the reference-faithful epoch leg (conf `batch`)")
    ap.add_argument("--no-minibatch", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel eagerly instead of CUDA-graph replay")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong", "frames"],
                    help="N > 1: weak = every rank renders the workload's whole view set (global batch N x views); strong = "
                         "the workload's views are dealt round-robin to the ranks (global batch fixed, SURVEY.md 8e); frames = "
                         "every rank optimises its OWN frame of a sequence (BASELINE.json configs[3]: independent problems, no "
                         "exchange at all; value = frames' iterations per second summed over the ranks)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "peer-oneshot", "peer-twoshot", "nccl"],
                    help="N > 1: fused NVLink peer-memory exchange inside the update kernel, or one NCCL all-reduce")
    args = ap.parse_args()
    # Libraries (NCCL's version banner) write to fd 1; the contract is ONE JSON line on stdout, so everything but the
    # final line is sent to stderr.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    args.warmup = max(args.warmup, 3)
    if args.ncc:  # the host-batch step and the oracle leg run the iteration without extra terms
        args.no_e2e = args.no_cpu_baseline = args.no_minibatch = True
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    warnings.filterwarnings("ignore")
    import torch
    import torch.distributed as dist
    from fmhr_b200 import _lib, synth
    from fmhr_b200.ham import HamOptimizer, HostStreamingStepper
    from fmhr_b200.render import render_views
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    _lib.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    wl = dict(synth.WORKLOADS[args.workload])
    strong = args.scaling == "strong" and world > 1
    frames = args.scaling == "frames" and world > 1
    if strong:
        # the workload's views (one camera set) dealt round-robin: rank r owns views r, r + world, ... (fmhr_b200.dist.shard_views)
        total_views = args.views or wl["n"]
        mine = len(range(rank, total_views, world))
        scene = synth.build_scene(wl, lambda *a: render_views(*a, device=dev), n_views=mine, view_offset=rank,
                                  view_stride=world, camera_seed=1)
    else:
        total_views = (args.views or wl["n"]) * world
        scene = synth.build_scene(wl, lambda *a: render_views(*a, device=dev), n_views=args.views, camera_seed=1 + rank)
    img_u8, msk_u8 = quantise_targets(scene)
    n, H, W = scene["imgs"].shape[0], scene["H"], scene["W"]
    c = lambda k, dt=torch.float32: torch.tensor(scene[k], dtype=dt, device=dev)
    opt = HamOptimizer(c("vertices"), c("faces", torch.int32), c("imgs"), c("masks"), c("valid_masks"), c("w2cs"),
                       c("projs"), c("sh_coeffs"), c("albedo"), scene["conf"], use_graphs=not args.no_graphs,
                       exchange=args.exchange if (world > 1 and not frames) else None,
                       n_views_global=total_views if (world > 1 and not frames) else None,
                       process_group=False if frames else None)
    ncc_term = None
    if args.ncc:
        if world > 1:
            raise SystemExit("bench.py: --ncc is single-GPU")
        from fmhr_b200.ncc_term import NccTerm
        gray = torch.tensor(np.asarray(scene["imgs"]).mean(-1).astype(np.float32), device=dev)
        ncc_term = NccTerm(opt, gray, ref_view=0, src_views=list(range(1, scene["imgs"].shape[0])), weight=10.0,
                           n_points=50000, half=5, seed=0)
    exchange = "none: every rank optimises an independent frame" if frames else "1 NCCL all-reduce/iter"
    kernels_per_step = KERNELS_PER_STEP + (2 if args.ncc else 0)  # own kernels per iteration (NCCL's all-reduce kernel is not counted)
    if opt.peer is not None:
        # one-shot: the exchange IS the normal-gradient kernel; two-shot: one extra reduce-scatter kernel in front of it
        kernels_per_step += 1 if (opt.peer.mode == 2 or (opt.peer.mode == 0 and world > 2)) else 0
        exchange = "NVLink peer-memory exchange fused into the update kernels (%s)" % (
            "two-shot reduce-scatter" if opt.peer.mode == 2 or (opt.peer.mode == 0 and world > 2) else "one-shot gather")
    V, F = opt.V, opt.T
    E = opt.topo.n_dir_edges // 2
    views = torch.arange(n, dtype=torch.int32, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident throughput
    for _ in range(args.warmup):
        opt.step_phase_b(views)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # nvidia-smi needs ~0.25 s to deliver its first sample; the GPU keeps stepping meanwhile (untimed), so that the timed
    # region does not start on a device that has just idled back to its base clocks
    extra = 0
    if world > 1:  # lock-stepped exchange: every rank must make the same number of steps
        extra = 1000
        for _ in range(extra):
            opt.step_phase_b(views)
    else:
        t_spin = time.time()
        while time.time() - t_spin < 0.3:
            for _ in range(20):
                opt.step_phase_b(views)
            torch.cuda.synchronize()
            extra += 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        opt.step_phase_b(views)
    e1.record()
    barrier()
    t1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_per_step = ms_total / args.steps
    # weak: iteration equivalents of the workload's view set per second over all ranks; strong: iterations of the one job
    value = (1.0 if strong else world) * 1000.0 / ms_per_step
    losses = opt.losses.cpu().tolist()
    if not all(np.isfinite(losses)) or losses[6] <= 0:
        raise SystemExit("bench.py: the optimisation state is not finite (losses %s) - refusing to report a number" % losses)

    # ---------------------------------------------------------------- end to end (host buffers)
    e2e = None
    if not args.no_e2e:
        # The step's HOST inputs (what the reference's loader produces from disk): the 8-bit image batch, the 8-bit
        # masks and the cameras, in pinned memory, uploaded inside the timed region; the loss record is read back every
        # step.  valid_masks stay resident: the reference derives them on the device (mesh_sfs_optim.py:146-163).
        pinf = lambda k: torch.tensor(scene[k], dtype=torch.float32).contiguous().pin_memory()
        h_imgs8 = torch.tensor(img_u8).contiguous().pin_memory()
        h_masks8 = torch.tensor(msk_u8).contiguous().pin_memory()
        h_w2cs, h_projs = pinf("w2cs"), pinf("projs")
        k_e2e = max(3, min(args.steps, 100))
        if world == 1 or opt.peer is not None or frames:
            stepper = HostStreamingStepper(opt, n)
            stepper.set_resident_valid_masks(opt.valid_masks)
            h2d, d2h = stepper.h2d_bytes_u8, stepper.d2h_bytes

            # double-buffered: batch i+1 is submitted (H2D on the copy stream) before step i is launched, so the
            # transfer of the next step's inputs overlaps this step's kernels; every step still uploads its own batch
            # from pinned host memory inside the timed region and reads its loss record back.
            pending = []

            def e2e_begin():
                pending.append(stepper.submit_u8(h_imgs8, h_masks8))

            def e2e_step(last=False):
                if not last:
                    pending.append(stepper.submit_u8(h_imgs8, h_masks8))
                stepper.step_submitted_u8(pending.pop(0), h_w2cs, h_projs, views)
                torch.cuda.current_stream().synchronize()   # the loss record is now readable on the host
                return stepper.losses_host
        else:
            d_imgs8, d_masks8 = torch.empty_like(h_imgs8, device=dev), torch.empty_like(h_masks8, device=dev)
            d_w2cs, d_projs = torch.empty_like(h_w2cs, device=dev), torch.empty_like(h_projs, device=dev)
            h2d, d2h = h_imgs8.numel() + h_masks8.numel() + 4 * (h_w2cs.numel() + h_projs.numel()), 32

            def e2e_begin():
                pass

            def e2e_step(last=False):
                d_imgs8.copy_(h_imgs8, non_blocking=True)
                d_masks8.copy_(h_masks8, non_blocking=True)
                d_w2cs.copy_(h_w2cs, non_blocking=True)
                d_projs.copy_(h_projs, non_blocking=True)
                opt.imgs.copy_(d_imgs8.to(torch.float32) / 255.0)
                opt.masks.copy_((d_masks8 > 127).to(torch.float32))
                opt.w2cs.copy_(d_w2cs)
                opt.projs.copy_(d_projs)
                return opt.step_phase_b(views).cpu()

        def time_leg(begin, step):
            begin()
            for i in range(3):
                step(last=i == 2)
            barrier()
            e0.record()
            begin()
            for i in range(k_e2e):
                step(last=i == k_e2e - 1)
            e1.record()
            barrier()
            ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
            return (1.0 if strong else world) * 1000.0 * k_e2e / float(ms2.item())

        full = {"value": time_leg(e2e_begin, e2e_step), "unit": UNIT, "steps": k_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "inputs": "the WHOLE 8-bit image batch + 8-bit masks + cameras from pinned host memory every step "
                          "(%s), loss record read back every step" % (
                              "fmhr_ham_host_u8_submit + fmhr_ham_step_host_u8_submitted, next batch in flight during "
                              "the step" if (world == 1 or opt.peer is not None or frames) else "torch copies + HamOptimizer.step_phase_b")}
        e2e = full
        if world == 1 or opt.peer is not None or frames:
            # Headline leg: same pipelined step, but of every view only the bounding box of its segmentation travels
            # (fmhr_ham_host_u8_submit_boxes).  Outside it the mask is 0, no pixel is valid (mesh_sfs_optim.py:276-281) and
            # no image byte is read; the boxes are loader metadata, computed once from the masks outside the timed region.
            boxes = HostStreamingStepper.mask_boxes(h_masks8)

            cams = (h_w2cs, h_projs)

            def b_begin():
                pending.append(stepper.submit_u8(h_imgs8, h_masks8, boxes, cameras=cams))

            def b_step(last=False):
                if not last:
                    pending.append(stepper.submit_u8(h_imgs8, h_masks8, boxes, cameras=cams))
                stepper.step_submitted_u8(pending.pop(0), None, None, views)
                torch.cuda.current_stream().synchronize()
                return stepper.losses_host

            # Same step, but the loss record goes through a two-slot pinned ring and is read ONE STEP LATE (after the
            # next step has been launched): every step still uploads its inputs and every record is still read on the
            # host inside the timed region, but the host no longer drains the device between steps.
            inflight = []

            def a_step(last=False):
                if not last:
                    pending.append(stepper.submit_u8(h_imgs8, h_masks8, boxes, cameras=cams))
                t = pending.pop(0)
                stepper.step_submitted_u8(t, None, None, views, async_record=True)
                rec = stepper.read_record(inflight.pop(0)[1]) if inflight else None
                inflight.append(t)
                if last:
                    rec = stepper.read_record(inflight.pop(0)[1])
                    if not bool(torch.isfinite(rec).all()):
                        raise SystemExit("bench.py: non-finite loss record in the end-to-end leg")
                return rec

            sync_leg = time_leg(b_begin, b_step)
            inputs = ("per view the bounding box of its segmentation out of the 8-bit image batch + 8-bit masks + the cameras, "
                      "pulled from pinned host memory every step and converted on arrival (fmhr_ham_host_u8_submit_boxes_direct; "
                      "pixels outside the boxes have mask 0 and are never read), next batch in flight during the step")
            e2e = {"value": time_leg(b_begin, a_step), "unit": UNIT, "steps": k_e2e,
                   "h2d_bytes_per_step": stepper.last_submit_bytes,
                   "d2h_bytes_per_step": d2h,
                   "inputs": inputs + "; the 32-byte loss record of every step is copied to a pinned ring and read on the "
                                      "host one step late (no device drain between steps)",
                   "sync_every_step": {"value": sync_leg, "unit": UNIT,
                                       "inputs": inputs + "; the host waits for every step's loss record before it launches "
                                                          "the next step (the reference's per-iteration .item())"},
                   "full_upload": full}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- roofline (rank 0)
    # the unit of the path is one whole iteration (one CUDA-graph launch of 13 kernels) and its
    # algorithmic bytes B_alg are the single figure for roofline.achieved; the dominant kernel is reported beside it.
    peak, peak_src = measured_peak_gbs()
    balg = b_alg_bytes(n, H, W, V, F, E) + (ncc_term.algorithmic_bytes() if ncc_term is not None else 0)
    iter_ach = balg / (ms_per_step * 1e-3) / 1e9
    traffic = measured_traffic()
    roof = {"bound": "hbm", "kernel": "fused HAM iteration (%d kernels, one graph launch)" % kernels_per_step,
            "achieved": iter_ach, "peak": peak, "unit": "GB/s", "frac": iter_ach / peak,
            "traffic": traffic["iteration_bytes"] if traffic else None, "algorithmic_bytes": balg,
            "iter_ms": ms_per_step, "peak_source": peak_src,
            "note": "SURVEY.md 8(d): B_alg / t_iter; B_alg is the dense model (68 B per pixel of every view), the kernels "
                    "only visit covered pixels so measured DRAM traffic is far below it - the iteration is bound by "
                    "instruction issue (coverage) and gather latency (pixel passes), not by HBM"}
    if world == 1:
        stages = opt.stage_times(views, repeats=10)
        top = max(stages, key=lambda k: stages[k])
        byts = stage_bytes(top, n, H, W, V, F)
        ach = byts / (stages[top] * 1e-3) / 1e9
        roof["dominant_kernel"] = {"kernel": top, "kernel_ms": stages[top], "algorithmic_bytes": byts, "achieved": ach,
                                   "frac": ach / peak,
                                   "traffic": (traffic or {}).get("kernels", {}).get(top),
                                   "share_of_step": stages[top] / sum(stages.values())}
        roof["stage_ms"] = stages
        # second half of BASELINE.json's metric: "rasterize GB/s vs HBM peak" (B_rast = 16nV + 12F + 16P
        # over the visibility pass alone; the fused path's coverage kernel also does the clip transform)
        b_rast = 16 * n * V + 12 * F + 16 * n * H * W
        roof["rasterize"] = {"algorithmic_bytes": b_rast, "kernel_ms": stages["coverage"],
                             "achieved": b_rast / (stages["coverage"] * 1e-3) / 1e9,
                             "frac": b_rast / (stages["coverage"] * 1e-3) / 1e9 / peak,
                             "mtri_per_s": n * F / (stages["coverage"] * 1e-3) / 1e6}

    cpu, parity = None, None
    if world == 1 and not args.no_cpu_baseline:
        # The oracle leg doubles as the parity check at the FULL benchmark shape: its warm-up iteration and one
        # iteration of the CUDA path start from the same state (the scene's initial state, quantised targets) and their
        # losses, n_valid and gradients are compared (oracle/compare.py); the second oracle iteration is the timed one.
        from oracle import compare as ocompare
        from oracle import ham as oham
        from oracle import raster as oraster
        oraster.build()
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        os.environ.setdefault("OMP_NUM_THREADS", str(threads))
        parity, st, _ = ocompare.ham_step_parity(scene, device=dev)
        t_cpu = time.time()
        oham.phase_b_step(st, list(range(n)))
        rate = 1.0 / (time.time() - t_cpu)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "full workload (%d views), 1 timed iteration after 1 warm-up (the warm-up is the parity iteration)" % n}

    # ---------------------------------------------------------------- reference-faithful mini-batches (secondary)
    # last leg of the run, and guarded: nothing it does can touch the numbers above
    epochs = None
    if world == 1 and not args.no_minibatch and n > args.minibatch:
        try:
            epochs = minibatch_epochs(opt, n, args.minibatch, max(10, args.steps // 4), dev, barrier)
        except Exception as ex:  # noqa: BLE001 - a secondary metric must not take the headline line down
            epochs = {"error": "%s: %s" % (type(ex).__name__, ex)}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "steps_per_s": 1000.0 / ms_per_step, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "views_per_gpu": n, "global_views": total_views, "H": H, "W": W, "verts": V,
                   "faces": F, "phase": "B (delta+albedo, conf/ih_sfs.conf weights)",
                   "launch": "eager" if args.no_graphs else "cuda-graph replay",
                   "warmup_extra_steps": extra,
                   "ncc": ("50,000 surface points x 121-pixel patches, reference view 0 vs %d source views, weight 10" % (n - 1))
                   if ncc_term is not None else None,
                   "l2": "per-iteration working set %.0f MB > 126 MB L2 (no flush needed)" % (
                       (8 + 32 + 20) * n * H * W / 1e6),
                   "parallelism": ("frames x%d (one independent frame of the sequence per rank), %s" % (world, exchange)) if frames
                   else "views x%d (%s), %s" % (world, "strong: the workload's views dealt round-robin" if strong else
                                                "weak: every rank renders its own view set", exchange)
                   if world > 1 else "single GPU"},
        "gpu_launches": kernels_per_step * args.steps,
        "clocks": clocks,
        "e2e": e2e,
        "epochs_minibatch": epochs,
        "roofline": roof,
        "cpu_baseline": cpu,
        "parity": parity,
        "losses_last": {k: v for k, v in zip(["sfs", "lap", "albedo", "mask", "edge", "delta", "n_valid", "total"], losses)},
    }
    emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Where the end-to-end step goes: host-side cost of the per-step calls against the device time (bench workload).

    python tools/diag_e2e_host.py [--steps 200]
Prints per step: wall time of submit_u8 / step_submitted_u8 calls alone (no reads, device free-running), the device
period of the same loop, and the device period with the conversion kernel's share (stage marks off)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fmhr_b200 import synth
from fmhr_b200.ham import HamOptimizer, HostStreamingStepper
from fmhr_b200.render import render_views

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--workload", default="interhand_48x512x334")
a = ap.parse_args()
dev = torch.device("cuda", 0)
scene = synth.build_scene(a.workload, lambda *x: render_views(*x, device=dev))
c = lambda k, dt=torch.float32: torch.tensor(scene[k], dtype=dt, device=dev)
opt = HamOptimizer(c("vertices"), c("faces", torch.int32), c("imgs"), c("masks"), c("valid_masks"), c("w2cs"),
                   c("projs"), c("sh_coeffs"), c("albedo"), scene["conf"], use_graphs=True)
n = scene["imgs"].shape[0]
views = torch.arange(n, dtype=torch.int32, device=dev)
img_u8 = np.clip(np.rint(np.asarray(scene["imgs"], dtype=np.float64) * 255.0), 0, 255).astype(np.uint8)
msk_u8 = (np.asarray(scene["masks"]) > 0).astype(np.uint8) * 255
pin = lambda x, dt: torch.tensor(x, dtype=dt).contiguous().pin_memory()
h_img, h_msk = pin(img_u8, torch.uint8), pin(msk_u8, torch.uint8)
h_w2c, h_proj = pin(scene["w2cs"], torch.float32), pin(scene["projs"], torch.float32)
st = HostStreamingStepper(opt, n)
st.set_resident_valid_masks(opt.valid_masks)
boxes = HostStreamingStepper.mask_boxes(msk_u8)


def loop(k, read):
    t_sub = t_step = 0.0
    pend = [st.submit_u8(h_img, h_msk, boxes, cameras=(h_w2c, h_proj))]
    infl = []
    for i in range(k):
        t0 = time.perf_counter()
        if i + 1 < k:
            pend.append(st.submit_u8(h_img, h_msk, boxes, cameras=(h_w2c, h_proj)))
        t1 = time.perf_counter()
        t = pend.pop(0)
        st.step_submitted_u8(t, None, None, views, async_record=True)
        t2 = time.perf_counter()
        t_sub += t1 - t0
        t_step += t2 - t1
        if read:
            if infl:
                st.read_record(infl.pop(0)[1])
            infl.append(t)
    return t_sub / k, t_step / k


loop(10, True)
torch.cuda.synchronize()
for read in (False, True):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    e0.record()
    ts, tp = loop(a.steps, read)
    e1.record()
    w1 = time.perf_counter()   # host done issuing (device may lag)
    torch.cuda.synchronize()
    w2 = time.perf_counter()
    print("read_one_step_late=%s: host submit %.1f us + step %.1f us per step; host loop %.1f us/step; device period %.1f us/step "
          "(wall incl. drain %.1f us/step)" % (read, ts * 1e6, tp * 1e6, (w1 - w0) / a.steps * 1e6,
                                               e0.elapsed_time(e1) / a.steps * 1e3, (w2 - w0) / a.steps * 1e6))
# resident graph replay for reference
for _ in range(20):
    opt.step_phase_b(views)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    opt.step_phase_b(views)
e1.record()
torch.cuda.synchronize()
print("resident graph replay: %.1f us/step" % (e0.elapsed_time(e1) / a.steps * 1e3))

#!/usr/bin/env python
"""Device timeline of the end-to-end (host-batch) step, from torch.profiler / CUPTI: every kernel and memcpy of a few
consecutive steps with stream, start offset and duration.  python tools/diag_e2e_trace.py [--out gpurun_out/e2e_trace.txt]"""
import argparse
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fmhr_b200 import synth
from fmhr_b200.ham import HamOptimizer, HostStreamingStepper
from fmhr_b200.render import render_views

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/e2e_trace.txt")
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


def loop(k):
    pend = [st.submit_u8(h_img, h_msk, boxes, cameras=(h_w2c, h_proj))]
    infl = []
    for i in range(k):
        if i + 1 < k:
            pend.append(st.submit_u8(h_img, h_msk, boxes, cameras=(h_w2c, h_proj)))
        t = pend.pop(0)
        st.step_submitted_u8(t, None, None, views, async_record=True)
        if infl:
            st.read_record(infl.pop(0)[1])
        infl.append(t)
    torch.cuda.synchronize()


loop(20)
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    loop(8)
tmp = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(tmp)
ev = [e for e in json.load(open(tmp))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
lines = []
for e in ev:
    name = e["name"].replace("fmhr::", "").split("(")[0][:60]
    lines.append("%10.1f %8.1f  stream %-4s %s" % (e["ts"] - t0, e["dur"], e.get("args", {}).get("stream", "?"), name))
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
open(a.out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[len(lines) // 2: len(lines) // 2 + 60]))

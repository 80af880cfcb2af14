"""Scratch diagnostic: why do stage times change after the e2e loop?"""
import sys, os, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from fmhr_b200 import synth
from fmhr_b200.ham import HamOptimizer, HostStreamingStepper
from fmhr_b200.render import render_views
dev = torch.device("cuda", 0)
scene = synth.build_scene("interhand_48x512x334", lambda *a: render_views(*a, device=dev))
c = lambda k, dt=torch.float32: torch.tensor(scene[k], dtype=dt, device=dev)
opt = HamOptimizer(c("vertices"), c("faces", torch.int32), c("imgs"), c("masks"), c("valid_masks"), c("w2cs"),
                   c("projs"), c("sh_coeffs"), c("albedo"), scene["conf"])
n = scene["imgs"].shape[0]
views = torch.arange(n, dtype=torch.int32, device=dev)
def timed(k):
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): opt.step_phase_b(views)
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/k
def report(tag):
    st = opt.stage_times(views, repeats=5)
    print(tag, "ms/step %.4f" % timed(100), "stages sum %.4f" % sum(st.values()), {k: round(v,4) for k,v in st.items()},
          "losses", [round(x,4) for x in opt.losses.cpu().tolist()], "delta nan", bool(torch.isnan(opt.delta).any()),
          "delta max", float(opt.delta.abs().max()))
report("A initial")
report("B again")
pin = lambda k: torch.tensor(scene[k], dtype=torch.float32).pin_memory()
h = [pin(k) for k in ("imgs","masks","valid_masks","w2cs","projs")]
stepper = HostStreamingStepper(opt, n)
for _ in range(5):
    stepper.step_phase_b(*h, views); torch.cuda.synchronize()
print("e2e losses", stepper.losses_host.tolist())
report("C after e2e")
report("D again")

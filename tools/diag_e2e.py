import sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from fmhr_b200 import synth
from fmhr_b200.ham import HamOptimizer, HostStreamingStepper
from fmhr_b200.render import render_views
dev = torch.device("cuda", 0)
scene = synth.build_scene("small", lambda *a: render_views(*a, device=dev))
c = lambda k, dt=torch.float32: torch.tensor(scene[k], dtype=dt, device=dev)
opt = HamOptimizer(c("vertices"), c("faces", torch.int32), c("imgs"), c("masks"), c("valid_masks"), c("w2cs"),
                   c("projs"), c("sh_coeffs"), c("albedo"), scene["conf"])
n = scene["imgs"].shape[0]
views = torch.arange(n, dtype=torch.int32, device=dev)
print("normal", opt.step_phase_b(views).cpu().tolist())
pin = lambda k: torch.tensor(scene[k], dtype=torch.float32).contiguous().pin_memory()
h = [pin(k) for k in ("imgs","masks","valid_masks","w2cs","projs")]
stepper = HostStreamingStepper(opt, n)
stepper.step_phase_b(*h, views); torch.cuda.synchronize()
print("e2e", stepper.losses_host.tolist())
for nm, d, o in (("imgs", stepper.d_imgs, opt.imgs), ("masks", stepper.d_masks, opt.masks), ("valid", stepper.d_valid, opt.valid_masks),
                 ("w2cs", stepper.d_w2cs, opt.w2cs), ("projs", stepper.d_projs, opt.projs)):
    print(nm, torch.equal(d, o), float((d - o).abs().max()))
print("delta nan", bool(torch.isnan(opt.delta).any()), "packed tail", opt.packed[-4:].tolist())

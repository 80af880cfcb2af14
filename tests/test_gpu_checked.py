"""Index assertions in place of compute-sanitizer (which this GPU pool refuses to run: tools/sanitize.sh, profiles/r2/
sanitizer.md).  A -DFMHR_CHECKED build of the library (variants/libfmhr_checked.so, tools/build_variant.sh) asserts every
write into a per-warp shared buffer / bounded work list of the coverage, scan and antialias kernels; this test drives the
fused iteration, the stand-alone rasteriser and the initialisation through it in a SUBPROCESS (the library path is fixed
at import) on the micropolygon and the coarse scene and expects a clean mask.  Skipped when the variant has not been built."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "variants", "libfmhr_checked.so")

SCRIPT = r"""
import sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, %r)
import numpy as np, torch
from fmhr_b200 import _lib, synth
from fmhr_b200.ham import HamOptimizer
from oracle import ham as oham
import nvdiffrast.torch as dr
lib = _lib.load()
assert lib.fmhr_debug_checks() == 0, "not a checked build / stale flags"
dev = torch.device("cuda")
for name in ("small", "coarse"):
    scene = synth.build_scene(name, oham.render_views)
    c = lambda k, dt=torch.float32: torch.tensor(scene[k], dtype=dt, device=dev)
    opt = HamOptimizer(c("vertices"), c("faces", torch.int32), c("imgs"), c("masks"), c("valid_masks"), c("w2cs"),
                       c("projs"), c("sh_coeffs"), c("albedo"), scene["conf"])
    views = list(range(scene["imgs"].shape[0]))
    opt.initialise(torch.tensor(scene["imgs"]).mean(-1).cuda())
    for _ in range(2):
        opt.step_phase_a(views)
    for _ in range(3):
        rec = opt.step_phase_b(views)
    assert bool(torch.isfinite(rec).all())
    pos = torch.randn(2, 300, 4, device=dev); pos[..., 3] = pos[..., 3].abs() + 0.1
    tri = torch.randint(0, 300, (500, 3), dtype=torch.int32, device=dev)
    dr.rasterize(dr.RasterizeGLContext(), pos, tri, resolution=(77, 334))
    flags = lib.fmhr_debug_checks()
    print("CHECKED", name, "flags", flags)
    assert flags == 0, "index assertion sites fired: mask 0x%%x" %% flags
print("CHECKED ok")
""" % ROOT


@pytest.mark.skipif(not os.path.exists(LIB), reason="variants/libfmhr_checked.so not built (tools/build_variant.sh checked -DFMHR_CHECKED)")
def test_checked_build_reports_no_out_of_bounds_index():
    env = dict(os.environ, FMHR_B200_LIB=LIB)
    r = subprocess.run([sys.executable, "-c", SCRIPT], env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "CHECKED ok" in r.stdout

#!/usr/bin/env python
"""Cross-check of the drop-in operators against UPSTREAM nvdiffrast, for a box that has it (this project's containers do not:
nvdiffrast is an un-vendored, un-pinned dependency of the reference, requirements.txt:14, and its rasteriser needs a GL or
CUDA context).  Never part of the test suite; it exists so that the day such a box is available the unpinned part of the
oracle can be pinned in one command:

    python tools/crosscheck_nvdiffrast.py [--workload small] [--cuda-context]

What it reports, per operator, on the synthetic scenes of fmhr_b200.synth (same tensors into both implementations):

* rasterize: fraction of pixels whose triangle id differs (expected: isolated silhouette / shared-edge pixels only - upstream's
  coverage is the GPU's fixed-function rule (GL) or its own CUDA rasteriser's, this project's is the documented rule of
  DESIGN.md section 2), max |u|, |v|, |z/w| difference on the pixels where the ids agree, rast_db likewise;
* interpolate (A = 7 and 30) and antialias (C = 1 and 3) forward on UPSTREAM's rast (so that coverage differences do not
  leak into them), max abs difference;
* backward of each: gradient w.r.t. pos / attr / color for a fixed random cotangent, max difference relative to the largest
  entry.

Upstream is imported from site-packages with this repository's `nvdiffrast/` shim package kept off sys.path."""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def import_upstream():
    """The real `nvdiffrast.torch`: import it with the repo root (which holds the shim package of the same name) removed
    from sys.path, then restore the path."""
    saved = list(sys.path)
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
    for k in [k for k in sys.modules if k == "nvdiffrast" or k.startswith("nvdiffrast.")]:
        del sys.modules[k]
    try:
        up = importlib.import_module("nvdiffrast.torch")
    except ImportError:
        raise SystemExit("crosscheck: upstream nvdiffrast is not installed on this box (nothing to compare against)")
    finally:
        sys.path[:] = saved
    if os.path.abspath(os.path.dirname(os.path.dirname(up.__file__))) == ROOT:
        raise SystemExit("crosscheck: `nvdiffrast.torch` resolved to this repository's shim - upstream nvdiffrast is not installed")
    for k in [k for k in sys.modules if k == "nvdiffrast" or k.startswith("nvdiffrast.")]:
        sys.modules["upstream_" + k] = sys.modules.pop(k)
    return up


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="small")
    ap.add_argument("--cuda-context", action="store_true", help="upstream RasterizeCudaContext instead of RasterizeGLContext")
    a = ap.parse_args()
    up = import_upstream()
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    from fmhr_b200 import dr as ours
    from fmhr_b200 import synth
    dev = torch.device("cuda")
    wl = synth.WORKLOADS[a.workload]
    v, f = synth.hand_mesh(wl["subdiv"], wl["hands"], seed=0)
    w2c, proj = synth.make_cameras(wl["n"], wl["H"], wl["W"], v.mean(0).astype(np.float64),
                                   extent=float(v[:, 1].max() - v[:, 1].min()))
    H, W, n = wl["H"], wl["W"], wl["n"]
    vh = torch.cat([torch.tensor(v), torch.ones(v.shape[0], 1)], 1)[None].expand(n, -1, -1)
    pos = torch.einsum('ijk,ikl->ijl', torch.einsum('ijk,ikl->ijl', vh, torch.tensor(w2c)), torch.tensor(proj)).contiguous().to(dev)
    tri = torch.tensor(f, dtype=torch.int32, device=dev)
    g = torch.Generator().manual_seed(0)
    ctx_up = up.RasterizeCudaContext() if a.cuda_context else up.RasterizeGLContext()
    ctx_our = ours.RasterizeGLContext()
    # ---- rasterize
    pu = pos.clone().requires_grad_(True)
    po = pos.clone().requires_grad_(True)
    ru, dbu = up.rasterize(ctx_up, pu, tri, resolution=(H, W))
    ro, dbo = ours.rasterize(ctx_our, po, tri, resolution=(H, W))
    same = ru[..., 3] == ro[..., 3]
    cov = (ru[..., 3] > 0) | (ro[..., 3] > 0)
    print("rasterize: ids differ on %.3e of the frame (%.3e of the covered pixels); covered %d / %d" % (
        float((~same).float().mean()), float((~same & cov).float().sum() / cov.float().sum().clamp_min(1)),
        int((ru[..., 3] > 0).sum()), int((ro[..., 3] > 0).sum())))
    m = same & (ru[..., 3] > 0)
    for k, name in enumerate(("u", "v", "z/w")):
        print("  max |%s| difference where ids agree: %.3e" % (name, float((ru[..., k] - ro[..., k])[m].abs().max())))
    print("  rast_db max difference where ids agree: %.3e" % float((dbu - dbo)[m].abs().max()))
    ct = torch.randn(ru.shape, generator=g).to(dev)
    ct[..., 2:] = 0
    (ru * ct * m[..., None]).sum().backward()
    (ro * ct * m[..., None]).sum().backward()
    print("  grad pos (cotangent on u, v of the agreeing pixels): rel %.3e" % rel(po.grad, pu.grad))
    # ---- interpolate / antialias on upstream's rast
    rast = ru.detach()
    for A in (7, 30):
        attr = torch.randn(n, v.shape[0], A, generator=g).to(dev)
        au, ao = attr.clone().requires_grad_(True), attr.clone().requires_grad_(True)
        ou, _ = up.interpolate(au, rast, tri)
        oo, _ = ours.interpolate(ao, rast, tri)
        c2 = torch.randn(ou.shape, generator=g).to(dev)
        (ou * c2).sum().backward()
        (oo * c2).sum().backward()
        print("interpolate A=%d: fwd max abs %.3e, grad attr rel %.3e" % (A, float((ou - oo).abs().max()), rel(ao.grad, au.grad)))
    for C in (1, 3):
        col = torch.rand(n, H, W, C, generator=g).to(dev) * (rast[..., 3:4] > 0)
        cu, co = col.clone().requires_grad_(True), col.clone().requires_grad_(True)
        pu2, po2 = pos.clone().requires_grad_(True), pos.clone().requires_grad_(True)
        ou = up.antialias(cu, rast, pu2, tri)
        oo = ours.antialias(co, rast, po2, tri)
        c2 = torch.randn(ou.shape, generator=g).to(dev)
        (ou * c2).sum().backward()
        (oo * c2).sum().backward()
        print("antialias C=%d: fwd max abs %.3e (pixels changed by upstream %d, by ours %d), grad color rel %.3e, grad pos rel %.3e" % (
            C, float((ou - oo).abs().max()), int(((ou - col).abs().sum(-1) > 0).sum()), int(((oo - col).abs().sum(-1) > 0).sum()),
            rel(co.grad, cu.grad), rel(po2.grad, pu2.grad)))


if __name__ == "__main__":
    main()

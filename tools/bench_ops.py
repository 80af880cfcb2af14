"""Per-operator throughput of the drop-in surface (nvdiffrast.torch shim + models/utils names) at BASELINE config 2
shapes (48 views x 512x334, sub3 mesh), the way mesh_sfs_optim.py:253-310 calls them: CUDA-event timed, 20 repetitions
after 3 warm-ups, every working set larger than the 126 MB L2.  Algorithmic bytes per call are the tensors the op must
read / write once (SURVEY.md 8d's per-op terms); the fraction is against the measured HBM copy peak.

    python tools/bench_ops.py [--out profiles/r1_ops.json]
"""
import argparse
import json
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch

from fmhr_b200 import dr, synth, utils
from fmhr_b200.render import render_views


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="interhand_48x512x334")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak = 6557.1
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    scene = synth.build_scene(args.workload, lambda *a: render_views(*a, device=dev))
    c = lambda k, dt=torch.float32: torch.tensor(scene[k], dtype=dt, device=dev)
    verts, faces = c("vertices"), c("faces", torch.int32)
    w2cs, projs = c("w2cs"), c("projs")
    n, H, W = scene["imgs"].shape[0], scene["H"], scene["W"]
    V, F = verts.shape[0], faces.shape[0]
    P = n * H * W
    vh = torch.cat([verts, torch.ones_like(verts[:, :1])], 1)
    pos = ((vh[None] @ w2cs) @ projs).contiguous()  # mesh_sfs_optim.py:262-264
    glctx = dr.RasterizeGLContext()
    rows = []

    def row(name, ms, nbytes, note=""):
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append({"op": name, "ms": ms, "algorithmic_bytes": nbytes, "GB/s": gbs, "frac_of_hbm_peak": gbs / peak,
                     "note": note})
        print("%-34s %8.3f ms  %8.1f MB  %7.0f GB/s  %5.1f %%  %s" % (name, ms, nbytes / 1e6, gbs, 100 * gbs / peak, note))

    # ---- rasterize (mesh_sfs_optim.py:267), default grad_db=True as the reference calls it
    ms = timed(lambda: dr.rasterize(glctx, pos, faces, resolution=(H, W)))
    row("rasterize fwd (+rast_db)", ms, 16 * n * V + 12 * F + 32 * P, "%.0f Mtri/s" % (n * F / ms / 1e3))
    ms = timed(lambda: dr.rasterize(glctx, pos, faces, resolution=(H, W), grad_db=False))
    row("rasterize fwd (grad_db=False)", ms, 16 * n * V + 12 * F + 16 * P, "%.0f Mtri/s" % (n * F / ms / 1e3))
    g1 = dr.RasterizeGLContext()
    g1.use_meshlets = False
    ms = timed(lambda: dr.rasterize(g1, pos, faces, resolution=(H, W)))
    row("rasterize fwd (+rast_db), v1 kernel", ms, 16 * n * V + 12 * F + 32 * P,
        "one thread per triangle + clear + dense resolve (fmhr_rasterize_fwd)")
    rast, _ = dr.rasterize(glctx, pos, faces, resolution=(H, W))
    r1, _ = dr.rasterize(g1, pos, faces, resolution=(H, W))
    assert torch.equal(rast, r1), "meshlet and v1 rasterisers must agree bit for bit"
    covered = int((rast[..., 3] > 0).sum())

    # ---- interpolate, A = 7 (normals | albedo | 1, :268) and A = 30 (train_mlp.py:180-184)
    for A in (7, 30):
        attr = torch.rand(n, V, A, device=dev)
        ms = timed(lambda: dr.interpolate(attr, rast, faces))
        row("interpolate fwd A=%d" % A, ms, 16 * P + 4 * A * P + 4 * A * n * V + 12 * F)
        attr_g = attr.clone().requires_grad_(True)
        rast_g = rast.clone().requires_grad_(True)
        out, _ = dr.interpolate(attr_g, rast_g, faces)
        dy = torch.rand_like(out)
        ms = timed(lambda: torch.autograd.grad(out, (attr_g, rast_g), dy, retain_graph=True))
        row("interpolate bwd A=%d" % A, ms, 16 * P + 4 * A * P + 8 * A * n * V + 16 * P + 12 * F,
            "grads to attr and rast")
        del attr, attr_g, rast_g, out, dy

    # ---- antialias C = 3 (image, :287) and C = 1 (mask, :274)
    for C in (3, 1):
        color = torch.rand(n, H, W, C, device=dev) * (rast[..., 3:] > 0)
        ms = timed(lambda: dr.antialias(color, rast, pos, faces))
        row("antialias fwd C=%d" % C, ms, 16 * P + 8 * C * P + 16 * n * V)
        color_g = color.clone().requires_grad_(True)
        pos_g = pos.clone().requires_grad_(True)
        out = dr.antialias(color_g, rast, pos_g, faces)
        dy = torch.rand_like(out)
        ms = timed(lambda: torch.autograd.grad(out, (color_g, pos_g), dy, retain_graph=True))
        row("antialias bwd C=%d" % C, ms, 16 * P + 12 * C * P + 32 * n * V, "grads to color and pos")
        del color, color_g, pos_g, out, dy

    # ---- rasterize bwd (grad of u, v to pos)
    pos_g = pos.clone().requires_grad_(True)
    r2, _ = dr.rasterize(glctx, pos_g, faces, resolution=(H, W))
    dy = torch.rand_like(r2)
    ms = timed(lambda: torch.autograd.grad(r2, pos_g, dy, retain_graph=True))
    row("rasterize bwd", ms, 32 * P + 32 * n * V + 12 * F)
    del pos_g, r2, dy

    # ---- models/utils names
    # the reference passes an EXPANDED view of one mesh, vertsw[:, :, :3] (mesh_sfs_optim.py:262,265): computed once
    fl = faces.long()
    v1 = verts.clone().requires_grad_(True)
    vexp = torch.cat([v1, torch.ones_like(v1[:, :1])], 1).unsqueeze(0).expand(n, -1, -1)[:, :, :3]
    ms = timed(lambda: utils.get_normals(vexp, fl))
    row("get_normals fwd (expanded [n,V,3])", ms, 24 * V + 12 * F)
    nrm = utils.get_normals(vexp, fl)
    dy = torch.rand(n, V, 3, device=dev)
    ms = timed(lambda: torch.autograd.grad(nrm, v1, dy, retain_graph=True))
    row("get_normals bwd (expanded [n,V,3])", ms, 12 * n * V + 36 * V + 12 * F, "incl. torch's sum over the n copies")
    x = verts.clone().requires_grad_(True)
    ms = timed(lambda: utils.laplacian_smoothing(x, fl, method="uniform"))
    row("laplacian_smoothing fwd", ms, 24 * V + 8 * 3 * F, "CSR cached on the mesh")
    nv = covered
    coeff = torch.rand(nv, 9, device=dev)
    nn = torch.nn.functional.normalize(torch.rand(nv, 3, device=dev), dim=1)
    ms = timed(lambda: utils.get_radiance(coeff, nn, 3))
    row("get_radiance fwd [N_valid]", ms, nv * (36 + 12 + 4))

    summary = {"workload": args.workload, "n": n, "H": H, "W": W, "V": V, "F": F, "covered_pixels": covered,
               "hbm_peak_gbs": peak, "timing": "CUDA events, 20 reps after 3 warm-ups", "ops": rows}
    if args.out:
        with open(args.out, "w") as f:
            json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()

"""ORACLE (test infrastructure) — measured parity of the product's fused HAM iteration against the oracle.

``ham_step_parity`` starts the CUDA path (fmhr_b200.ham.HamOptimizer, debug gradients on) and the restated reference
loop (oracle.ham, mesh_sfs_optim.py:253-310 / :198-237) from the SAME state, runs ONE iteration on each and returns the
measured errors; callers (tests/, bench.py's cpu_baseline leg) assert or report them.  Tolerances of BASELINE.json's
north_star: ids / coverage / n_valid exact, losses and images 1e-5 relative, gradients 1e-4 relative to the largest
gradient entry.
"""
import time

import numpy as np
import torch

from . import ham as oham
from . import raster as orc

LOSS_TERMS = ("sfs", "lap", "albedo", "mask", "edge", "delta")
TOL_LOSS = 1e-5    # relative, per loss term                     (north_star: "1e-5 relative on images" / losses)
TOL_IMAGE = 1e-5   # relative to the largest image value         (north_star)
TOL_GRAD = 1e-4    # relative to the largest gradient entry      (north_star: "1e-4 on gradients")


def rel_to_max(a, b):
    """max |a - b| / max |b|"""
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def make_optimizer(scene, device, **kw):
    from fmhr_b200.ham import HamOptimizer
    c = lambda k, dt=torch.float32: torch.tensor(np.asarray(scene[k]), dtype=dt, device=device)
    return HamOptimizer(c("vertices"), c("faces", torch.int32), c("imgs"), c("masks"), c("valid_masks"), c("w2cs"),
                        c("projs"), c("sh_coeffs"), c("albedo"), scene["conf"], **kw)


def ham_step_parity(scene, views=None, device="cuda", phase="b", planes=False, state=None):
    """One iteration from identical state on both sides.  Returns (report dict, oracle HamState, oracle seconds).

    report: per-term relative loss errors, n_valid of both sides, gradient errors relative to the largest entry of the
    oracle's gradient (delta / albedo, SH in phase A) and - with planes=True - the forward planes: triangle ids
    bit-exact against the oracle rasteriser on the product's own clip positions, id mismatches against the oracle's
    own positions, antialiased image / coverage errors on the pixels both sides assign to the same triangle."""
    dev = torch.device(device)
    n = np.asarray(scene["imgs"]).shape[0]
    views = list(range(n)) if views is None else list(views)
    opt = make_optimizer(scene, dev, debug=True)
    st = state or oham.HamState(scene)
    rep = {"views": len(views), "H": int(scene["H"]), "W": int(scene["W"]), "verts": int(opt.V), "faces": int(opt.T),
           "phase": phase}
    keep = {}
    if planes and phase == "b":
        ex = opt.export(views)
        oham.phase_b_forward(st, views, keep=keep)
        pos = ex["pos"].cpu()
        rep["pos_rel"] = rel_to_max(pos, keep["proj_verts"])
        ref_rast, _, _ = orc.rasterize_fwd(pos, st.faces, (st.H, st.W), want_db=False)
        ours = ex["rast"].cpu()
        rep["rast_bit_exact_on_own_positions"] = bool(torch.equal(ours, ref_rast))
        same = ours[..., 3] == keep["rast_out"][..., 3]
        rep["id_mismatch_frac_vs_oracle_positions"] = float(1.0 - same.float().mean())
        # pixels both sides assign to the same triangle (1-ulp differences between the einsum-built and the fused clip
        # positions can flip isolated silhouette pixels): what remains is interpolate / normalise / SH / antialias
        img_scale = float(keep["tmp_img"].abs().max().clamp_min(1e-30))
        rep["image_rel_same_pixels"] = float((ex["image"].cpu() - keep["tmp_img"]).abs()[same].max()) / img_scale
        rep["coverage_abs_same_pixels"] = float((ex["pred_mask"].cpu() - keep["pred_mask"]).abs()[same].max())
        del ex, ours, ref_rast, pos
        keep = {}
    t0 = time.time()
    ref = oham.phase_b_step(st, views, keep=keep) if phase == "b" else oham.phase_a_step(st, views, keep=keep)
    oracle_s = time.time() - t0
    rec = (opt.step_phase_b(views) if phase == "b" else opt.step_phase_a(views)).cpu().tolist()
    names = LOSS_TERMS if phase == "b" else ("sfs",)
    idx = {name: k for k, name in enumerate(LOSS_TERMS)}
    rep["loss_rel"] = {name: abs(rec[idx[name]] - ref[name]) / max(abs(ref[name]), 1e-30) for name in names}
    rep["losses"] = {name: rec[idx[name]] for name in names}
    rep["losses_oracle"] = {name: ref[name] for name in names}
    rep["n_valid"] = int(rec[6])
    rep["n_valid_oracle"] = int(ref["n_valid"])
    g = opt.dbg_grad.cpu()
    if phase == "b":
        rep["grad_delta_rel"] = rel_to_max(g[:, :3], keep["grad_delta"])
    rep["grad_albedo_rel"] = rel_to_max(g[:, 3:], keep["grad_albedo"][0])
    if phase == "a":
        rep["grad_sh_rel"] = rel_to_max(opt.dbg_grad_sh.cpu(), keep["grad_sh"])
    rep["max_loss_rel"] = max(rep["loss_rel"].values())
    # pass / fail per term with the fp32 resolution of the loss record as absolute floor (a term that is ~0 by
    # cancellation, e.g. the mask loss of the initial mesh against its own valid_masks, has no meaningful relative error)
    loss_ok = all(abs(rep["losses"][k] - rep["losses_oracle"][k]) <= TOL_LOSS * abs(rep["losses_oracle"][k]) + 1e-7
                  for k in names)
    rep["tolerances"] = {"loss_rel": TOL_LOSS, "image_rel": TOL_IMAGE, "grad_rel_to_max": TOL_GRAD, "n_valid": "exact"}
    ok = rep["n_valid"] == rep["n_valid_oracle"] and loss_ok
    for k in ("grad_delta_rel", "grad_albedo_rel", "grad_sh_rel"):
        if k in rep:
            ok = ok and rep[k] <= TOL_GRAD
    if "image_rel_same_pixels" in rep:
        ok = ok and rep["rast_bit_exact_on_own_positions"] and rep["image_rel_same_pixels"] <= TOL_IMAGE
    rep["within_tolerance"] = bool(ok)
    return rep, st, oracle_s

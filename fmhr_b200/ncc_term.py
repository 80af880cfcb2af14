"""NCC photo-consistency term of a phase-B HAM iteration (BASELINE.json configs[2]: "16 views x 1024x1024 with NCC loss").

The reference ships the arithmetic (models/ncc_utils.py:4-35, `NCC`) but no caller, so the wiring is this
project's:

* `num_sample` points (conf/demo_sfs.conf:15: 50,000) are drawn once, area-weighted, on the initial mesh and stay attached
  to their faces through barycentric coordinates;
* every iteration each point is projected into the reference view and the source views, a (2 half + 1)^2 patch (11 x 11 =
  121 pixels) of the gray image is sampled bilinearly around each projection, and
      loss = weight * mean over (source view, point) of (1 - NCC(ref patch, src patch, src mask));
* the gradient flows through the source sampling positions to the vertices (the reference patch is a constant of the step)
  and joins the iteration through fmhr_ham_add_delta_grad, between the render and the update half.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr


def sample_surface_points(vertices, faces, n_points, seed=0):
    """Area-weighted points on the mesh: (face ids int32 [Np], barycentrics float32 [Np,2])."""
    v, f = np.asarray(vertices, dtype=np.float64), np.asarray(faces, dtype=np.int64)
    area = 0.5 * np.linalg.norm(np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]]), axis=1)
    rng = np.random.default_rng(seed)
    face = rng.choice(f.shape[0], size=n_points, p=area / area.sum())
    r1, r2 = np.sqrt(rng.uniform(size=n_points)), rng.uniform(size=n_points)
    bary = np.stack([1.0 - r1, r1 * (1.0 - r2)], axis=1)
    return face.astype(np.int32), bary.astype(np.float32)


class NccTerm:
    """One instance per optimiser.  grayimgs [num,H,W] (CUDA); ref_view / src_views index the optimiser's view arrays."""

    def __init__(self, opt, grayimgs, ref_view, src_views, weight, n_points=50000, half=5, seed=0, fused=True):
        self.lib = _lib.load()
        dev = opt.device
        self.weight, self.half, self.npx = float(weight), int(half), (2 * int(half) + 1) ** 2
        face, bary = sample_surface_points(opt.vertices_tmp.cpu().numpy(), opt.faces.cpu().numpy(), n_points, seed)
        self.pt_face, self.pt_bary = torch.from_numpy(face).to(dev), torch.from_numpy(bary).to(dev)
        self.view_idx = torch.tensor([int(ref_view)] + [int(s) for s in src_views], dtype=torch.int32, device=dev)
        self.nv1, self.np_ = self.view_idx.numel(), n_points
        self.gray = grayimgs.detach().to(device=dev, dtype=torch.float32).contiguous()
        if self.gray.shape != opt.masks.shape:
            raise RuntimeError("fmhr_b200.NccTerm: grayimgs must be [num,H,W]")
        f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        self.patches = self.patch_mask = self.grad_patches = self.grad_ncc = None   # unfused chain only (_alloc_patches)
        self.fused = fused
        self.ncc = f32(self.nv1 - 1, n_points)
        self.vertices = f32(opt.V, 3)
        self.grad_delta = f32(opt.V, 3)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        opt.extra_terms.append(self)

    # bytes the term moves per iteration in the model of SURVEY.md 8(d): source patches written + read (forward, backward),
    # the reference patch, the ncc values
    def algorithmic_bytes(self):
        nv = self.nv1 - 1
        return 8 * nv * self.np_ * self.npx + 4 * self.np_ * self.npx + 4 * nv * self.np_

    def accumulate(self, opt, cfg, buf, sp, fused=None):
        """Enqueue the term on stream `sp` (between fmhr_ham_step_render and fmhr_ham_step_update; graph-capturable).
        fused (default: whenever the patch has at most 128 samples): ONE kernel samples, evaluates and back-propagates
        (fmhr_ncc_term_fused); otherwise the four-call chain that materialises patches / masks / patch gradients (also
        what the parity tests inspect: self.patches)."""
        lib = self.lib
        V, T, H, W = opt.V, opt.T, opt.H, opt.W
        torch.add(opt.vertices_tmp, opt.delta, out=self.vertices)
        args = (ptr(self.vertices), ptr(opt.faces), ptr(self.pt_face), ptr(self.pt_bary), ptr(opt.w2cs), ptr(opt.projs),
                ptr(self.view_idx), self.nv1, ptr(self.gray))
        nv = self.nv1 - 1
        if fused is None:
            fused = self.fused and self.npx <= 128
        self.grad_delta.zero_()
        if fused:
            check(lib.fmhr_ncc_term_fused(*args, ptr(opt.masks), V, self.np_, H, W, self.half,
                                          -self.weight / float(nv * self.np_), ptr(self.ncc), ptr(self.grad_delta), sp),
                  "ncc_term_fused")
        else:
            self._alloc_patches()
            check(lib.fmhr_ncc_sample_fwd(*args, ptr(opt.masks), V, self.np_, H, W, self.half, ptr(self.patches),
                                          ptr(self.patch_mask), sp), "ncc_sample_fwd")
            ref, src, msk = self.patches[0:1], self.patches[1:], self.patch_mask[1:]
            check(lib.fmhr_ncc_fwd(ptr(ref), ptr(src), ptr(msk), nv, self.np_, self.npx, ptr(self.ncc), sp), "ncc_fwd")
            check(lib.fmhr_ncc_bwd(ptr(ref), ptr(src), ptr(msk), ptr(self.grad_ncc), nv, self.np_, self.npx,
                                   ptr(self.grad_patches[1:]), sp), "ncc_bwd")
            check(lib.fmhr_ncc_sample_bwd(*args, V, self.np_, H, W, self.half, ptr(self.grad_patches), ptr(self.grad_delta),
                                          sp), "ncc_sample_bwd")
        check(lib.fmhr_ham_add_delta_grad(ctypes.byref(cfg), ctypes.byref(buf), ptr(self.grad_delta), sp),
              "ham_add_delta_grad")
        torch.mul(1.0 - self.ncc.mean(), self.weight, out=self.loss)

    def _alloc_patches(self):
        if self.patches is None:
            dev = self.gray.device
            f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
            self.patches, self.patch_mask = f32(self.nv1, self.np_, self.npx), f32(self.nv1, self.np_, self.npx)
            self.grad_patches = torch.zeros(self.nv1, self.np_, self.npx, dtype=torch.float32, device=dev)
            self.grad_ncc = torch.full((self.nv1 - 1, self.np_), -self.weight / float((self.nv1 - 1) * self.np_),
                                       dtype=torch.float32, device=dev)

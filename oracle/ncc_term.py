"""ORACLE (test infrastructure) — the NCC photo-consistency term in plain PyTorch (CPU, autograd).

The arithmetic of NCC itself is the reference's (oracle.refmath.NCC restates models/ncc_utils.py:4-35 and is pinned by
the golden fixtures); the wiring around it (surface points, patch sampling, which patches carry gradient) has no
counterpart in the reference and is specified in DESIGN.md - this file is its executable statement, the
CUDA path (fmhr_b200/csrc/ncc_loop.cu) is checked against it.
"""
import torch

from .refmath import NCC


def _bilinear(img, su, sv):
    """img [H,W]; su, sv [...] continuous coordinates (integers = pixel centres).  Returns (value, all-four-taps-inside,
    nearest x, nearest y); taps outside the frame read 0.  Differentiable w.r.t. su / sv."""
    H, W = img.shape
    x0f, y0f = torch.floor(su.detach()), torch.floor(sv.detach())
    fx, fy = su - x0f, sv - y0f
    x0, y0 = x0f.long(), y0f.long()

    def tap(x, y):
        ok = (x >= 0) & (x < W) & (y >= 0) & (y < H)
        return torch.where(ok, img[y.clamp(0, H - 1), x.clamp(0, W - 1)], torch.zeros((), dtype=img.dtype)), ok

    (i00, a), (i10, b), (i01, c), (i11, d) = tap(x0, y0), tap(x0 + 1, y0), tap(x0, y0 + 1), tap(x0 + 1, y0 + 1)
    val = (1 - fy) * ((1 - fx) * i00 + fx * i10) + fy * ((1 - fx) * i01 + fx * i11)
    xn = x0 + (fx.detach() >= 0.5).long()
    yn = y0 + (fy.detach() >= 0.5).long()
    return val, a & b & c & d, xn, yn


def ncc_term(vertices, faces, pt_face, pt_bary, w2cs, projs, view_idx, gray, masks, weight, half):
    """Returns (loss, ncc [Nv,Np], patches [Nv+1,Np,Npx], patch_mask).  vertices may require grad."""
    H, W = gray.shape[1:]
    f = faces.long()[pt_face.long()]
    b0, b1 = pt_bary[:, 0:1], pt_bary[:, 1:2]
    X = b0 * vertices[f[:, 0]] + b1 * vertices[f[:, 1]] + (1 - b0 - b1) * vertices[f[:, 2]]
    Xh = torch.cat([X, torch.ones_like(X[:, :1])], 1)
    d = torch.arange(-half, half + 1, dtype=torch.float32)
    dy, dx = torch.meshgrid(d, d, indexing="ij")
    dx, dy = dx.reshape(1, -1), dy.reshape(1, -1)
    patches, pmasks = [], []
    for s, view in enumerate(view_idx.tolist()):
        clip = Xh @ (w2cs[view] @ projs[view])
        u = (clip[:, 0] / clip[:, 3] * 0.5 + 0.5) * W - 0.5
        v = (clip[:, 1] / clip[:, 3] * 0.5 + 0.5) * H - 0.5
        su, sv = u[:, None] + dx, v[:, None] + dy
        val, inside, xn, yn = _bilinear(gray[view], su, sv)
        front = (clip[:, 3] > 0)[:, None]
        mk = inside & front & (masks[view][yn.clamp(0, H - 1), xn.clamp(0, W - 1)] > 0.5)
        patches.append(torch.where(front, val, torch.zeros(())))
        pmasks.append(mk.float())
    patches, pmasks = torch.stack(patches), torch.stack(pmasks)
    ncc = NCC(patches[0:1].detach(), patches[1:], None, pmasks[1:])
    loss = weight * (1.0 - ncc).mean()
    return loss, ncc, patches, pmasks

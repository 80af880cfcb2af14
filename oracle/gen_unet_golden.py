"""ORACLE (test infrastructure) — golden vectors of the reference's second caller of the boundary.

Runs only in the build container (needs /root/reference).  /root/reference/train_unet.py is imported verbatim and its
`unet_forward` (:155-198) is executed UNCHANGED, in both modes (`if_geo=False`: the 30-wide feature stack, `if_geo=True`:
the 10-wide one), inside one stage-2 iteration of neural_render.train (:186-208: image losses of both nets, the reference's
own `laplacian_smoothing`, edge hinge, delta loss), with

  * `nvdiffrast.torch` -> oracle.raster (the CPU restatement; the real dependency is absent),
  * `models.utils`     -> the reference's own file (verbatim),
  * the two networks and the positional encoder -> small seeded stand-ins (`standin_nets`): they are arguments of
    `unet_forward`, dense conv / MLP work outside the hot path, and only have to be the SAME
    differentiable functions on both sides of the comparison.

The one random line of `unet_forward` (`uni_vertices = vertices.clone().uniform_(0, 1)`) draws from torch's CPU generator
here (seeded); the drawn values are stored so that the GPU replay feeds the same numbers.

Outputs go to tests/golden/unet_forward_v1.npz; tests/test_gpu_neural_render.py replays the same lines on the CUDA shim.

    python -m oracle.gen_unet_golden
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

from .gen_golden import OUT, REF, import_reference

SEED = 7
B, H, W = 3, 64, 44          # W % 8 != 0: exercises unet_forward's padding branch (:192-194)


class StandinPE(torch.nn.Module):
    """[..., 3] -> [..., 87] like the reference's PostionalEncoding(min_deg=0, max_deg=1): identity + sin / cos of 42
    seeded directions."""

    def __init__(self, seed=SEED):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.register_buffer("dirs", torch.randn(3, 42, generator=g) * 2.0)

    def forward(self, x):
        p = x @ self.dirs.to(x.device)
        return torch.cat([x, torch.sin(p), torch.cos(p)], -1)


def standin_nets(seed=SEED):
    """(net 284 -> 3, net_g 12 -> 3, pe): 3x3 convolutions with seeded weights in place of UNet(284, 3, 2, 0) / UNet(12, 3, 2, 0)."""
    g = torch.Generator().manual_seed(seed + 1)
    net = torch.nn.Conv2d(284, 3, 3, padding=1)
    net_g = torch.nn.Conv2d(12, 3, 3, padding=1)
    with torch.no_grad():
        for m in (net, net_g):
            m.weight.copy_(torch.randn(m.weight.shape, generator=g) * 0.05)
            m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1 + 0.4)
    return net, net_g, StandinPE(seed)


def make_inputs():
    from fmhr_b200 import synth
    v, f = synth.hand_mesh(0, 1, seed=0)
    w2c, proj = synth.make_cameras(B, H, W, v.mean(0).astype(np.float64), extent=float(v[:, 1].max() - v[:, 1].min()), seed=4)
    rng = np.random.default_rng(SEED)
    V = v.shape[0]
    return dict(vertices_tmp=v.astype(np.float32), vertices=(v + rng.normal(0, 4e-4, v.shape)).astype(np.float32),
                faces=f.astype(np.int32), w2cs=w2c.astype(np.float32), projs=proj.astype(np.float32),
                albedo=rng.uniform(0.3, 0.9, (1, V, 3)).astype(np.float32),
                vertex_feat=rng.normal(0, 1, (V, 20)).astype(np.float32),
                rays=rng.normal(0, 1, (B, H, W, 3)).astype(np.float32),
                imgs=rng.uniform(0, 1, (B, H, W, 3)).astype(np.float32),
                gt_masks=(rng.uniform(0, 1, (B, H, W)) > 0.3).astype(np.float32))


def stage2_iteration(forward, laplacian_smoothing, nets, inp, dev="cpu"):
    """neural_render.py:176-208 for one batch holding every view (the optimiser step itself is torch.optim.Adam on
    both sides and is left out): returns (loss terms, outputs, gradients)."""
    net, net_g, pe = nets
    t = lambda k, dt=torch.float32: torch.as_tensor(inp[k], dtype=dt, device=dev)
    vertices_tmp, faces = t("vertices_tmp"), t("faces", torch.int32)
    vertices = t("vertices").clone().requires_grad_(True)
    albedo = t("albedo").clone().requires_grad_(True)
    vertex_feat = t("vertex_feat").clone().requires_grad_(True)
    w2c, proj, img, gt_mask, ray = t("w2cs"), t("projs"), t("imgs"), t("gt_masks"), t("rays")
    n = w2c.shape[0]
    resolution = (img.shape[1], img.shape[2])
    glctx = None
    a = vertices_tmp[faces[:, 0].long()]
    b = vertices_tmp[faces[:, 1].long()]
    c = vertices_tmp[faces[:, 2].long()]
    edge_length_mean = torch.cat([((a - b) ** 2).sum(1), ((c - b) ** 2).sum(1), ((a - c) ** 2).sum(1)]).mean()
    # ---- neural_render.py:189-208 ----
    render_z, masks = forward(net, pe, glctx, [ray, w2c, proj, vertices.unsqueeze(0).expand(n, -1, -1),
                                                faces, albedo.expand(n, -1, -1), vertex_feat], resolution)
    render_imgs, masks = forward(net_g, pe, glctx, [ray, w2c, proj, vertices.unsqueeze(0).expand(n, -1, -1),
                                                    faces, albedo.expand(n, -1, -1), render_z.detach()], resolution, True)
    valid_index = (masks[:, :, :, 0] > 0) & (gt_mask > 0)
    img_loss = F.l1_loss(render_imgs[valid_index], img[valid_index])
    imgz_loss = F.l1_loss(render_z[valid_index], img[valid_index])
    lap_loss = 100 * laplacian_smoothing(vertices, faces.long(), method="uniform")
    mask_loss = F.l1_loss(masks[:, :, :, 0], gt_mask) * 0
    a = vertices[faces[:, 0].long()]
    b = vertices[faces[:, 1].long()]
    c = vertices[faces[:, 2].long()]
    edge_length = torch.cat([((a - b) ** 2).sum(1), ((c - b) ** 2).sum(1), ((a - c) ** 2).sum(1)])
    edge_loss = torch.clip(edge_length - edge_length_mean, 0, 1).mean() * 150000
    delta_loss = ((vertices_tmp - vertices) ** 2).sum(1).mean() * 50000
    loss = img_loss + imgz_loss + lap_loss + mask_loss + edge_loss + delta_loss
    for p_ in list(net.parameters()) + list(net_g.parameters()):
        p_.grad = None
    loss.backward()
    terms = dict(img=float(img_loss), imgz=float(imgz_loss), lap=float(lap_loss), edge=float(edge_loss), delta=float(delta_loss))
    outs = dict(render_z=render_z.detach().cpu(), render_imgs=render_imgs.detach().cpu(), masks=masks.detach().cpu())
    grads = dict(vertices=vertices.grad.cpu(), albedo=albedo.grad.cpu(), vertex_feat=vertex_feat.grad.cpu(),
                 net_w=net.weight.grad.detach().cpu().clone(), net_g_w=net_g.weight.grad.detach().cpu().clone())
    return terms, outs, grads


def import_train_unet(dr_module):
    """/root/reference/train_unet.py, verbatim, with stand-ins for the modules this container lacks."""
    import_reference()
    nvd = types.ModuleType("nvdiffrast")
    nvd.torch = dr_module
    saved = {k: sys.modules.get(k) for k in ("nvdiffrast", "nvdiffrast.torch", "pyhocon", "skimage.metrics", "train_unet")}
    sys.modules["nvdiffrast"], sys.modules["nvdiffrast.torch"] = nvd, dr_module
    ph = types.ModuleType("pyhocon")
    ph.ConfigFactory = None
    sys.modules["pyhocon"] = ph
    sm = types.ModuleType("skimage.metrics")
    sm.structural_similarity = None
    sys.modules["skimage.metrics"] = sm
    sys.modules.pop("train_unet", None)
    try:
        import train_unet
    finally:
        for k, m in saved.items():
            if k == "train_unet":
                continue
            if m is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = m
    return train_unet


def run_reference(inp):
    """The reference's own unet_forward + laplacian_smoothing; also records what the random line drew and the clip positions."""
    from . import raster as oraster
    tu = import_train_unet(oraster)
    import models.utils as ru
    nets = standin_nets()  # (before the hook: nn.Conv2d's own initialisation draws uniform numbers too)
    drawn, pos = [], []
    orig_uniform, orig_rast = torch.Tensor.uniform_, oraster.rasterize

    def uniform_(self, *a, **k):
        r = orig_uniform(self, *a, **k)
        drawn.append(self.detach().clone())
        return r

    def rasterize(glctx, p, *a, **k):
        pos.append(p.detach().clone())
        return orig_rast(glctx, p, *a, **k)

    torch.Tensor.uniform_, oraster.rasterize = uniform_, rasterize
    try:
        torch.manual_seed(SEED)
        terms, outs, grads = stage2_iteration(tu.unet_forward, ru.laplacian_smoothing, nets, inp)
    finally:
        torch.Tensor.uniform_, oraster.rasterize = orig_uniform, orig_rast
    assert len(drawn) == 2 and len(pos) == 2 and torch.equal(pos[0], pos[1])
    return terms, outs, grads, drawn, pos[0]


def main():
    inp = make_inputs()
    terms, outs, grads, drawn, pos = run_reference(inp)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "unet_forward_v1.npz")
    np.savez_compressed(path, **{"in_" + k: v for k, v in inp.items()},
                        uni_vertices_0=drawn[0].numpy(), uni_vertices_1=drawn[1].numpy(), proj_verts=pos.numpy(),
                        **{"loss_" + k: np.float64(v) for k, v in terms.items()},
                        **{"out_" + k: v.numpy() for k, v in outs.items()},
                        **{"grad_" + k: v.numpy() for k, v in grads.items()})
    print(path, os.path.getsize(path), "bytes;", terms, "covered pixels", int((outs["masks"] > 0).sum()))


if __name__ == "__main__":
    main()

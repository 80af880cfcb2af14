"""Forward renderer on the CUDA operators (host-side helper, follows mesh_sfs_optim.py:138-148 + :282-287).

Used to synthesise benchmark targets and the initial `valid_masks` (antialiased coverage of the initial mesh,
mesh_sfs_optim.py:146,158,163) on the device; the CPU tests use oracle.ham.render_views with the same signature.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import dr, utils


def render_views(vertices, faces, albedo, sh, w2cs, projs, H, W, device="cuda", chunk=16):
    """-> numpy (img[n,H,W,3], coverage[n,H,W], aa_coverage[n,H,W])."""
    dev = torch.device(device)
    t = lambda a, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)
    vertices, faces, albedo, sh = t(vertices), t(faces, torch.int32), t(albedo), t(sh)
    w2cs, projs = t(w2cs), t(projs)
    glctx = dr.RasterizeGLContext()
    imgs, covs, aas = [], [], []
    with torch.no_grad():
        normals1 = utils.get_normals(vertices[None], faces)
        vertsw1 = torch.cat([vertices, torch.ones_like(vertices[:, :1])], 1)[None]
        for k in range(0, w2cs.shape[0], chunk):
            w2c, proj, shc = w2cs[k:k + chunk], projs[k:k + chunk], sh[k:k + chunk]
            n = w2c.shape[0]
            vertsw = vertsw1.expand(n, -1, -1)
            proj_verts = torch.einsum('ijk,ikl->ijl', torch.einsum('ijk,ikl->ijl', vertsw, w2c), proj)
            rast_out, _ = dr.rasterize(glctx, proj_verts, faces, resolution=(H, W), grad_db=False)
            feat = torch.cat([normals1.expand(n, -1, -1), albedo[None].expand(n, -1, -1),
                              torch.ones_like(vertsw[:, :, :1])], dim=2)
            feat, _ = dr.interpolate(feat, rast_out, faces)
            pred_normals = F.normalize(feat[..., :3], p=2, dim=3)
            aa_mask = dr.antialias(feat[..., 6:7].contiguous(), rast_out, proj_verts, faces).squeeze(-1)
            valid_idx = torch.where(rast_out[..., 3] > 0)
            radiance = utils.get_radiance(shc[valid_idx[0]], pred_normals[valid_idx], 3).unsqueeze(-1)
            img = torch.zeros(n, H, W, 3, device=dev)
            img[valid_idx] = radiance * feat[..., 3:6][valid_idx]
            img = dr.antialias(img, rast_out, proj_verts, faces)
            imgs.append(img.cpu())
            covs.append((rast_out[..., 3] > 0).float().cpu())
            aas.append(aa_mask.cpu())
    return torch.cat(imgs).numpy(), torch.cat(covs).numpy(), torch.cat(aas).numpy()

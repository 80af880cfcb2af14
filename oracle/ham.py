"""ORACLE (test infrastructure) — PyTorch-CPU restatement of the reference's HAM loops.

``phase_b_step`` follows mesh_sfs_optim.py:253-310 line for line and ``phase_a_step`` follows
mesh_sfs_optim.py:198-237, with ``oracle.raster`` standing in for ``nvdiffrast.torch`` and
``oracle.refmath`` for ``models.utils``.  ``render_views`` follows the init render at
mesh_sfs_optim.py:138-148 plus the shading of :282-287 and is used to synthesise targets.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import raster as dr
from .refmath import get_normals, get_radiance, laplacian_smoothing


# "spec": the forward VALUES of the clip positions follow the fixed evaluation order the GPU path and this checker
# share (oracle.raster.clip_positions; DESIGN.md section 2, rule 0), while autograd still flows through the reference's
# two einsums.  "einsum": the reference lines alone - ATen's summation order, 1 ulp away from the GPU's, which flips
# isolated silhouette pixels (coverage is discontinuous in the positions) and is therefore only used to REPORT that effect.
POSITIONS = "spec"


def _clip_positions(vertices, w2c, proj):
    """mesh_sfs_optim.py:262-264."""
    n = w2c.shape[0]
    vertsw = torch.cat([vertices, torch.ones_like(vertices[:, 0:1])], axis=1).unsqueeze(0).expand(n, -1, -1)
    rot_verts = torch.einsum('ijk,ikl->ijl', vertsw, w2c)
    proj_verts = torch.einsum('ijk,ikl->ijl', rot_verts, proj)
    if POSITIONS == "spec":
        from . import raster as _oraster  # (tests may rebind `dr` to the CUDA shim: the position rule stays the checker's)
        spec = _oraster.clip_positions(vertices.detach().cpu(), w2c.cpu(), proj.cpu()).to(proj_verts.device)
        # a + (b - a) == b exactly for floats 1 ulp apart (b - a is exact, the sum is representable)
        proj_verts = proj_verts + (spec - proj_verts.detach())
    return vertsw, proj_verts


def render_views(vertices, faces, albedo, sh, w2cs, projs, H, W):
    """Returns numpy (img[n,H,W,3], coverage[n,H,W], aa_coverage[n,H,W])."""
    with torch.no_grad():
        vertices = torch.as_tensor(vertices, dtype=torch.float32)
        faces = torch.as_tensor(faces, dtype=torch.int32)
        albedo = torch.as_tensor(albedo, dtype=torch.float32)
        sh = torch.as_tensor(sh, dtype=torch.float32)
        w2c = torch.as_tensor(w2cs, dtype=torch.float32)
        proj = torch.as_tensor(projs, dtype=torch.float32)
        n = w2c.shape[0]
        glctx = dr.RasterizeGLContext()
        vertsw, proj_verts = _clip_positions(vertices, w2c, proj)
        normals = get_normals(vertsw[:, :, :3], faces.long())
        rast_out, _ = dr.rasterize(glctx, proj_verts, faces, resolution=(H, W))
        feat = torch.cat([normals, albedo[None].expand(n, -1, -1), torch.ones_like(vertsw[:, :, :1])], dim=2)
        feat, _ = dr.interpolate(feat, rast_out, faces)
        pred_normals = F.normalize(feat[..., :3].contiguous(), p=2, dim=3)
        rast_albedo = feat[..., 3:6].contiguous()
        pred_mask = feat[..., 6:7].contiguous()
        aa_mask = dr.antialias(pred_mask, rast_out, proj_verts, faces).squeeze(-1)
        valid_idx = torch.where(rast_out[..., 3] > 0)
        radiance = get_radiance(sh[valid_idx[0]], pred_normals[valid_idx], 3).unsqueeze(-1)
        img = torch.zeros(n, H, W, 3)
        img[valid_idx] = radiance * rast_albedo[valid_idx]
        img = dr.antialias(img, rast_out, proj_verts, faces)
        cov = (rast_out[..., 3] > 0).float()
    return img.numpy(), cov.numpy(), aa_mask.numpy()


class HamState:
    """Optimisation state exactly as mesh_sfs_optim.py:176-193,242-244 sets it up."""

    def __init__(self, scene):
        t = lambda a, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a), dtype=dt)
        self.H, self.W = scene["H"], scene["W"]
        self.conf = dict(scene["conf"])
        self.faces = t(scene["faces"], torch.int32)
        self.vertices_tmp = t(scene["vertices"])
        self.imgs, self.masks, self.valid_masks = t(scene["imgs"]), t(scene["masks"]), t(scene["valid_masks"])
        self.w2cs, self.projs = t(scene["w2cs"]), t(scene["projs"])
        self.albedo = t(scene["albedo"]).unsqueeze(0).clone().requires_grad_(True)
        self.sh_coeffs = t(scene["sh_coeffs"]).clone().requires_grad_(True)
        self.delta = torch.zeros_like(self.vertices_tmp).requires_grad_(True)
        v, f = self.vertices_tmp, self.faces.long()
        a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
        # mesh_sfs_optim.py:184-188
        self.edge_length_mean = torch.cat([((a - b) ** 2).sum(1), ((c - b) ** 2).sum(1), ((a - c) ** 2).sum(1)]).mean()
        self.glctx = dr.RasterizeGLContext()
        self.opt_a = None
        self.opt_b = None

    def optimizer_a(self):
        if self.opt_a is None:  # mesh_sfs_optim.py:193
            c = self.conf
            self.opt_a = torch.optim.Adam([{'params': self.albedo, 'lr': c["albedo_lr"]},
                                           {'params': self.sh_coeffs, 'lr': c["sh_lr"]}])
        return self.opt_a

    def optimizer_b(self):
        if self.opt_b is None:  # mesh_sfs_optim.py:242-244
            c = self.conf
            self.sh_coeffs.requires_grad_(False)
            self.opt_b = torch.optim.Adam([{'params': self.delta, 'lr': c["lr"]},
                                           {'params': self.albedo, 'lr': c["albedo_lr"]},
                                           {'params': self.sh_coeffs, 'lr': c["sh_lr"]}])
        return self.opt_b


def phase_b_forward(st, view_idx, albedo_weight=None, keep=None):
    """Forward of one phase-B iteration (mesh_sfs_optim.py:253-306).  Returns (loss, dict of loss terms)."""
    c = st.conf
    faces = st.faces
    resolution = (st.H, st.W)
    albedo_weight = c["albedo_weight"] if albedo_weight is None else albedo_weight
    perm = torch.as_tensor(view_idx, dtype=torch.long)
    vertices = st.vertices_tmp + st.delta
    n = perm.numel()
    w2c, proj = st.w2cs[perm], st.projs[perm]
    img, mask, valid_mask = st.imgs[perm], st.masks[perm], st.valid_masks[perm]
    sh_coeff = st.sh_coeffs[perm]

    vertsw, proj_verts = _clip_positions(vertices, w2c, proj)
    normals = get_normals(vertsw[:, :, :3], faces.long())

    rast_out, _ = dr.rasterize(st.glctx, proj_verts, faces, resolution=resolution)
    feat = torch.cat([normals, st.albedo.expand(n, -1, -1), torch.ones_like(vertsw[:, :, :1])], dim=2)
    feat, _ = dr.interpolate(feat, rast_out, faces)
    pred_normals = feat[:, :, :, :3].contiguous()
    rast_albedo = feat[:, :, :, 3:6].contiguous()
    pred_mask = feat[:, :, :, 6:7].contiguous()
    pred_normals = F.normalize(pred_normals, p=2, dim=3)
    pred_mask = dr.antialias(pred_mask, rast_out, proj_verts, faces).squeeze(-1)

    valid_idx = torch.where((mask > 0) & (rast_out[:, :, :, 3] > 0))
    valid_normals = pred_normals[valid_idx]
    valid_shcoeff = sh_coeff[valid_idx[0]]
    valid_albedo = rast_albedo[valid_idx]
    valid_img = img[valid_idx]
    radiance = get_radiance(valid_shcoeff, valid_normals, 3).unsqueeze(-1)
    pred_img = radiance * valid_albedo

    tmp_img = torch.zeros_like(img)
    tmp_img[valid_idx] = pred_img
    tmp_img = dr.antialias(tmp_img, rast_out, proj_verts, faces)

    sfs_loss = c["sfs_weight"] * F.l1_loss(tmp_img[valid_idx], valid_img)
    lap_loss = c["lap_weight"] * laplacian_smoothing(vertices, faces.long(), method="uniform")
    albedo_loss = albedo_weight * laplacian_smoothing(st.albedo.squeeze(0), faces.long(), method="uniform")
    mask_loss = c["mask_weight"] * F.mse_loss(pred_mask, valid_mask)
    a = vertices[faces[:, 0].long()]
    b = vertices[faces[:, 1].long()]
    cc = vertices[faces[:, 2].long()]
    edge_length = torch.cat([((a - b) ** 2).sum(1), ((cc - b) ** 2).sum(1), ((a - cc) ** 2).sum(1)])
    edge_loss = torch.clip(edge_length - st.edge_length_mean, 0, 1).mean() * c["edge_weight"]
    delta_loss = (st.delta ** 2).sum(1).mean() * c["delta_weight"]
    loss = sfs_loss + lap_loss + albedo_loss + mask_loss + delta_loss + edge_loss
    terms = dict(sfs=sfs_loss, lap=lap_loss, albedo=albedo_loss, mask=mask_loss, edge=edge_loss, delta=delta_loss,
                 n_valid=valid_idx[0].numel())
    if keep is not None:
        keep.update(rast_out=rast_out.detach(), proj_verts=proj_verts.detach(), tmp_img=tmp_img.detach(),
                    pred_mask=pred_mask.detach(), normals=normals.detach(), vertices=vertices.detach())
    return loss, terms


def phase_b_step(st, view_idx, albedo_weight=None, keep=None):
    """One full phase-B iteration: forward, backward, Adam (mesh_sfs_optim.py:253-310)."""
    opt = st.optimizer_b()
    loss, terms = phase_b_forward(st, view_idx, albedo_weight, keep)
    opt.zero_grad()
    loss.backward()
    if keep is not None:
        keep.update(grad_delta=st.delta.grad.detach().clone(), grad_albedo=st.albedo.grad.detach().clone())
    opt.step()
    return {k: (float(v) if torch.is_tensor(v) else v) for k, v in terms.items()}


def phase_a_step(st, view_idx, keep=None):
    """One phase-A (albedo + SH warm-up) iteration, mesh_sfs_optim.py:198-237."""
    c = st.conf
    opt = st.optimizer_a()
    faces = st.faces
    perm = torch.as_tensor(view_idx, dtype=torch.long)
    vertices = (st.vertices_tmp + st.delta).detach()
    n = perm.numel()
    w2c, proj = st.w2cs[perm], st.projs[perm]
    img, mask = st.imgs[perm], st.masks[perm]
    sh_coeff = st.sh_coeffs[perm]
    vertsw, proj_verts = _clip_positions(vertices, w2c, proj)
    normals = get_normals(vertsw[:, :, :3], faces.long())
    rast_out, _ = dr.rasterize(st.glctx, proj_verts, faces, resolution=(st.H, st.W))
    feat = torch.cat([normals, st.albedo.expand(n, -1, -1)], dim=2)
    feat, _ = dr.interpolate(feat, rast_out, faces)
    pred_normals = feat[:, :, :, :3].contiguous()
    rast_albedo = feat[:, :, :, 3:6].contiguous()
    pred_normals = dr.antialias(pred_normals, rast_out, proj_verts, faces)
    pred_normals = F.normalize(pred_normals, p=2, dim=3)
    rast_albedo = dr.antialias(rast_albedo, rast_out, proj_verts, faces)
    valid_idx = torch.where((mask > 0) & (rast_out[:, :, :, 3] > 0))
    valid_normals = pred_normals[valid_idx]
    valid_shcoeff = sh_coeff[valid_idx[0]]
    valid_albedo = rast_albedo[valid_idx]
    valid_img = img[valid_idx]
    radiance = get_radiance(valid_shcoeff, valid_normals, 3).unsqueeze(-1)
    pred_img = radiance * valid_albedo
    sfs_loss = c["sfs_weight"] * F.l1_loss(pred_img, valid_img)
    albedo_loss = c["albedo_weight"] * laplacian_smoothing(st.albedo.squeeze(0), faces.long(), method="uniform")
    loss = sfs_loss
    opt.zero_grad()
    loss.backward()
    if keep is not None:
        keep.update(grad_albedo=st.albedo.grad.detach().clone(), grad_sh=st.sh_coeffs.grad.detach().clone(),
                    pred_img=pred_img.detach(), valid_idx=valid_idx)
    opt.step()
    return dict(sfs=float(sfs_loss), albedo=float(albedo_loss), n_valid=valid_idx[0].numel())


def ham_init(vertices, faces, imgs, grayimgs, masks, w2cs, projs, H, W, degree=3):
    """HAM initialisation, mesh_sfs_optim.py:124-177 line for line (oracle.raster for nvdiffrast, numpy lstsq as in the
    reference): per view the antialiased coverage of the initial mesh (`valid_masks`, :146,163) and the least-squares SH
    lighting of the antialiased, re-normalised normals onto the gray image (:152-153); then the global SH fit over all
    views (:165-166) and the mean albedo img / radiance over the valid pixels (:173-174).
    Returns dict(valid_masks [num,H,W], sh_coeffs [num,9], sh_coeff [9], albedo_mean [3], n_valid [num])."""
    from .refmath import get_matrix
    vertices = torch.as_tensor(vertices, dtype=torch.float32)
    faces = torch.as_tensor(faces, dtype=torch.int32)
    imgs, grayimgs, masks = (torch.as_tensor(a, dtype=torch.float32) for a in (imgs, grayimgs, masks))
    w2cs, projs = torch.as_tensor(w2cs, dtype=torch.float32), torch.as_tensor(projs, dtype=torch.float32)
    glctx = dr.RasterizeGLContext()
    num = imgs.shape[0]
    with torch.no_grad():
        valid_normals, valid_grayimgs, valid_masks, valid_imgs, sh_coeffs, counts = [], [], [], [], [], []
        for k in range(num):
            w2c, proj = w2cs[k:k + 1], projs[k:k + 1]
            mask, img, grayimg = masks[k:k + 1], imgs[k:k + 1], grayimgs[k:k + 1]
            vertsw, proj_verts = _clip_positions(vertices, w2c, proj)
            normals = get_normals(vertsw[:, :, :3], faces.long())
            rast_out, _ = dr.rasterize(glctx, proj_verts, faces, resolution=(H, W))
            feat, _ = dr.interpolate(torch.cat([normals, torch.ones_like(vertsw[:, :, :1])], 2), rast_out, faces)
            pred_normals = feat[:, :, :, :3].contiguous()
            pred_mask = feat[:, :, :, 3:4].contiguous()
            pred_mask = dr.antialias(pred_mask, rast_out, proj_verts, faces).squeeze(-1)
            pred_normals = dr.antialias(pred_normals, rast_out, proj_verts, faces)
            pred_normals = F.normalize(pred_normals, p=2, dim=3)
            valid_idx = (mask > 0) & (rast_out[:, :, :, 3] > 0)
            valid_normal = pred_normals[valid_idx].numpy()
            valid_grayimg = grayimg[valid_idx].numpy()
            matrix = get_matrix(torch.from_numpy(valid_normal), degree).numpy()
            sh_coeff = np.linalg.lstsq(matrix, valid_grayimg, rcond=None)[0]
            valid_normals.append(valid_normal)
            valid_imgs.append(img[valid_idx])
            valid_grayimgs.append(valid_grayimg)
            valid_masks.append(pred_mask)
            sh_coeffs.append(torch.from_numpy(sh_coeff.astype(np.float32)).unsqueeze(0))
            counts.append(int(valid_idx.sum()))
        valid_normals = np.concatenate(valid_normals, axis=0)
        valid_imgs = torch.cat(valid_imgs, 0)
        valid_grayimgs = np.concatenate(valid_grayimgs, axis=0)
        valid_masks = torch.cat(valid_masks, 0)
        matrix = get_matrix(torch.from_numpy(valid_normals), degree).numpy()
        sh_coeff = torch.from_numpy(np.linalg.lstsq(matrix, valid_grayimgs, rcond=None)[0].astype(np.float32))
        radiance = get_radiance(sh_coeff, torch.from_numpy(valid_normals), degree).unsqueeze(-1)
        albedo_mean = (valid_imgs / radiance).mean(0)
    return dict(valid_masks=valid_masks, sh_coeffs=torch.cat(sh_coeffs, 0), sh_coeff=sh_coeff, albedo_mean=albedo_mean,
                n_valid=counts)

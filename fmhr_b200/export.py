"""Result files of the HAM stage in the reference's formats (mesh_sfs_optim.py:19-28, 321-343), so that the downstream
scripts (neural_render.py:90-96 reads `<scan>.pt` and `<scan>.obj`) can consume a run of the fused optimiser unchanged.

    <scan>.pt     torch.save({'sh_coeff': [num,9], 'albedo': [1,V,3]})                     (:323)
    <scan>.obj    refined mesh, vertex and face order preserved (trimesh export, :328-329)
    ori_<scan>.obj  the subdivided input mesh before refinement (:114-118)
    <scan>_c.obj  vertices with per-vertex colour clamp(0.5 * albedo, 0, 1) in RGB (albedo is BGR), faces with
                  FLIPPED winding f0 f2 f1 (save_obj_mesh_with_color, :19-28, :336-337)
    rerender/mesh_%02d.png  the last epoch's antialiased renders, one per view (:339-345) - optional

Host-side code (numpy / torch CPU): nothing here is on the hot path.
"""
import os

import numpy as np
import torch


def save_obj_mesh(mesh_path, verts, faces):
    """Plain OBJ, 1-based faces in the given winding (what trimesh's exporter writes for an un-processed mesh)."""
    verts, faces = np.asarray(verts), np.asarray(faces)
    with open(mesh_path, "w") as f:
        for v in verts:
            f.write("v %.8f %.8f %.8f\n" % (v[0], v[1], v[2]))
        for t in faces:
            f.write("f %d %d %d\n" % (t[0] + 1, t[1] + 1, t[2] + 1))


def save_obj_mesh_with_color(mesh_path, verts, faces, colors):
    """mesh_sfs_optim.py:19-28: `v x y z r g b` with %.4f, faces written as (f0, f2, f1) + 1."""
    verts, faces, colors = np.asarray(verts), np.asarray(faces), np.asarray(colors)
    with open(mesh_path, "w") as f:
        for v, c in zip(verts, colors):
            f.write("v %.4f %.4f %.4f %.4f %.4f %.4f\n" % (v[0], v[1], v[2], c[0], c[1], c[2]))
        for t in faces:
            f.write("f %d %d %d\n" % (t[0] + 1, t[2] + 1, t[1] + 1))


def save_ham_results(out_dir, scan_id, vertices, faces, albedo, sh_coeffs, rendered=None, perm_last=None,
                     ori_vertices=None):
    """Writes the reference's result set.  vertices [V,3], faces [F,3], albedo [V,3] or [1,V,3] (BGR), sh_coeffs [num,9];
    rendered [k,H,W,3] in [0,1] (BGR, as cv2 expects) with perm_last giving each image's view index; ori_vertices: the
    subdivided input mesh, written as ori_<scan>.obj (mesh_sfs_optim.py:114-118)."""
    os.makedirs(out_dir, exist_ok=True)
    if ori_vertices is not None:
        save_obj_mesh(os.path.join(out_dir, "ori_%d.obj" % scan_id), torch.as_tensor(ori_vertices).detach().float().cpu().numpy(),
                      torch.as_tensor(faces).detach().cpu().numpy())
    v = torch.as_tensor(vertices).detach().float().cpu()
    f = torch.as_tensor(faces).detach().cpu()
    a = torch.as_tensor(albedo).detach().float().cpu().reshape(1, -1, 3)
    sh = torch.as_tensor(sh_coeffs).detach().float().cpu()
    torch.save({"sh_coeff": sh, "albedo": a}, os.path.join(out_dir, "%d.pt" % scan_id))
    save_obj_mesh(os.path.join(out_dir, "%d.obj" % scan_id), v.numpy(), f.numpy())
    color = torch.clamp(0.5 * a, 0, 1).numpy()[0][:, 2::-1]  # BGR -> RGB
    save_obj_mesh_with_color(os.path.join(out_dir, "%d_c.obj" % scan_id), v.numpy(), f.numpy(), color)
    if rendered is not None:
        import cv2
        os.makedirs(os.path.join(out_dir, "rerender"), exist_ok=True)
        imgs = torch.as_tensor(rendered).detach().float().cpu().numpy()
        for i, idx in enumerate(perm_last if perm_last is not None else range(imgs.shape[0])):
            cv2.imwrite(os.path.join(out_dir, "rerender", "mesh_%02d.png" % int(idx)), (imgs[i] * 255).astype(np.int32))
    return os.path.join(out_dir, "%d.pt" % scan_id)


def load_obj(mesh_path):
    """Minimal OBJ reader for the files above: returns (verts [V,3], colors [V,3] or None, faces [F,3] 0-based)."""
    vs, cs, fs = [], [], []
    for line in open(mesh_path):
        p = line.split()
        if not p:
            continue
        if p[0] == "v":
            vs.append([float(x) for x in p[1:4]])
            if len(p) >= 7:
                cs.append([float(x) for x in p[4:7]])
        elif p[0] == "f":
            fs.append([int(x.split("/")[0]) - 1 for x in p[1:4]])
    return np.asarray(vs), (np.asarray(cs) if cs else None), np.asarray(fs, dtype=np.int64)

"""ORACLE (test infrastructure) — PyTorch-CPU restatement of the reference's shading / geometry math.

Each function cites the reference lines it follows.  PINNED: tests/test_oracle_golden.py checks
every function here against fixtures produced by the reference's own code (oracle/gen_golden.py
imports /root/reference/models/utils.py and models/ncc_utils.py verbatim in the build container).
"""
import torch
import torch.nn.functional as F


def get_normals(vertices, faces):
    """Area-weighted vertex normals.  models/utils.py:508-548.

    vertices [B,V,3], faces [F,3] int64.  Accumulation order (corner 1, corner 2, corner 0) and the
    normalisation eps (1e-6) follow the reference.
    """
    vn = torch.zeros_like(vertices)
    vf = vertices[:, faces]  # B F 3 3
    p0, p1, p2 = vf[:, :, 0], vf[:, :, 1], vf[:, :, 2]
    vn.index_add_(1, faces[:, 1], torch.cross(p2 - p1, p0 - p1, dim=2))
    vn.index_add_(1, faces[:, 2], torch.cross(p0 - p2, p1 - p2, dim=2))
    vn.index_add_(1, faces[:, 0], torch.cross(p1 - p0, p2 - p0, dim=2))
    return F.normalize(vn, p=2, dim=2, eps=1e-6)


def get_matrix(normal, degree=3):
    """SH polynomial basis rows [1, y, z, x, xy, yz, 2z^2-x^2-y^2, zx, x^2-y^2].  models/utils.py:188-206."""
    x, y, z = normal[:, 0], normal[:, 1], normal[:, 2]
    cols = [torch.ones_like(x)]
    if degree > 1:
        cols += [y, z, x]
    if degree > 2:
        cols += [x * y, y * z, 2 * z * z - x * x - y * y, z * x, x * x - y * y]
    return torch.stack(cols, dim=1)


def get_radiance(coeff, normal, degree=3):
    """radiance = coeff . basis(normal), accumulated term by term.  models/utils.py:208-226."""
    x, y, z = normal[:, 0], normal[:, 1], normal[:, 2]
    r = coeff[..., 0]
    if degree > 1:
        r = r + coeff[..., 1] * y
        r = r + coeff[..., 2] * z
        r = r + coeff[..., 3] * x
    if degree > 2:
        r = r + coeff[..., 4] * x * y
        r = r + coeff[..., 5] * y * z
        r = r + coeff[..., 6] * (2 * z * z - x * x - y * y)
        r = r + coeff[..., 7] * z * x
        r = r + coeff[..., 8] * (x * x - y * y)
    return r


def get_edges(verts, faces):
    """Unique undirected edges [E,2] (sorted pairs, ascending hash V*a+b).  models/utils.py:551-571."""
    V = verts.shape[0]
    v0, v1, v2 = faces[:, 0], faces[:, 1], faces[:, 2]
    e = torch.cat([torch.stack([v1, v2], 1), torch.stack([v2, v0], 1), torch.stack([v0, v1], 1)], 0)
    e, _ = e.sort(dim=1)
    h = torch.unique(V * e[:, 0] + e[:, 1])
    return torch.stack([h // V, h % V], dim=1)


def compute_laplacian(verts, faces):
    """Uniform graph Laplacian L = D^-1 A - I as sparse COO.  models/utils.py:661-693."""
    V = verts.shape[0]
    edges = get_edges(verts, faces)
    e0, e1 = edges[:, 0], edges[:, 1]
    idx = torch.cat([torch.stack([e0, e1], 0), torch.stack([e1, e0], 0)], dim=1)
    A = torch.sparse_coo_tensor(idx, torch.ones(idx.shape[1], dtype=torch.float32), (V, V))
    deg = torch.sparse.sum(A, dim=1).to_dense()
    d0 = deg[e0]
    d0 = torch.where(d0 > 0.0, 1.0 / d0, d0)
    d1 = deg[e1]
    d1 = torch.where(d1 > 0.0, 1.0 / d1, d1)
    L = torch.sparse_coo_tensor(idx, torch.cat([d0, d1]), (V, V))
    di = torch.arange(V)
    L = L - torch.sparse_coo_tensor(torch.stack([di, di], 0), torch.ones(V, dtype=torch.float32), (V, V))
    return L


def laplacian_smoothing(verts, faces, method="uniform"):
    """sum_i ||(L x)_i||_2 / V.  models/utils.py:696-722 (uniform branch only; every call site in
    mesh_sfs_optim.py:231,292,293 passes method="uniform")."""
    assert method == "uniform"
    with torch.no_grad():
        L = compute_laplacian(verts, faces)
    loss = L.mm(verts).norm(dim=1) * (1.0 / verts.shape[0])
    return loss.sum()


def NCC(ref, src, ref_valid_mask, src_valid_mask):
    """Masked normalised cross-correlation.  models/ncc_utils.py:4-35.

    ref [1,Np,Npx], src / src_valid_mask [Nv,Np,Npx] -> [Nv,Np].  ref_valid_mask is ignored, the
    valid count of an empty patch becomes 1 and a zero variance becomes 1, exactly as the reference.
    """
    nv = src.shape[0]
    cnt = src_valid_mask.sum(dim=2, keepdim=True)
    cnt = torch.where(cnt == 0, torch.ones_like(cnt), cnt)
    r = ref.expand(nv, -1, -1)
    r_mean = (r * src_valid_mask).sum(dim=2, keepdim=True) / cnt
    r_var = ((r - r_mean) * src_valid_mask).square().sum(dim=2, keepdim=True) / cnt
    r_var = r_var + (r_var == 0).to(r_var.dtype)
    s_mean = (src * src_valid_mask).sum(dim=2, keepdim=True) / cnt
    s_var = ((src - s_mean) * src_valid_mask).square().sum(dim=2, keepdim=True) / cnt
    s_var = s_var + (s_var == 0).to(s_var.dtype)
    cov = ((r - r_mean) * (src - s_mean) * src_valid_mask).sum(dim=2, keepdim=True) / cnt
    return (cov / (torch.sqrt(r_var) * torch.sqrt(s_var))).squeeze()

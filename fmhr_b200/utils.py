"""CUDA counterparts of the `models/utils.py` / `models/ncc_utils.py` functions the HAM loop calls
(same names, argument meaning and return values as the reference), each one C-ABI call per direction:

    get_normals(vertices, faces)               models/utils.py:508-548   (mesh_sfs_optim.py:141,210,265)
    get_matrix(normal, degree)                 models/utils.py:188-206   (mesh_sfs_optim.py:153,165)
    get_radiance(coeff, normal, degree)        models/utils.py:208-226   (mesh_sfs_optim.py:173,227,282)
    laplacian_smoothing(verts, faces, method)  models/utils.py:696-722   (mesh_sfs_optim.py:231,292,293)
    NCC(ref, src, ref_valid_mask, src_valid_mask)  models/ncc_utils.py:4-35

The sparse Laplacian / face adjacency are built once per face tensor (fmhr_mesh_topology_build) instead of on
every call.
"""
import weakref

import torch

from . import _lib
from ._lib import check, ptr, stream
from .dr import Topology

_TOPO = {}  # (shape, n_verts, device) -> list of (weakref(tensor)|None, version, tri_int32, Topology)


def topology_for(faces, n_verts):
    """Topology of `faces` ([F,3] int32 or int64).  The same tensor object (unchanged version) is trusted; a new
    object is compared element-wise with the cached triangle list (one device sync) before reuse."""
    key = (tuple(faces.shape), int(n_verts), str(faces.device))
    entries = _TOPO.setdefault(key, [])
    for ref, ver, tri32, topo in entries:
        obj = ref() if ref is not None else None
        if obj is faces and ver == faces._version:
            return topo
    tri_new = faces.to(torch.int32).contiguous()
    for i, (ref, ver, tri32, topo) in enumerate(entries):
        if torch.equal(tri32, tri_new):
            entries[i] = (weakref.ref(faces), faces._version, tri32, topo)
            return topo
    topo = Topology(tri_new, n_verts)
    entries.append((weakref.ref(faces), faces._version, tri_new, topo))
    del entries[:-4]
    return topo


class _VertexNormals(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, topo):
        lib = _lib.load()
        verts = verts.contiguous()
        V = verts.shape[0]
        normals = torch.empty_like(verts)
        raw = torch.empty_like(verts)
        with torch.cuda.device(verts.device):
            check(lib.fmhr_vertex_normals_fwd(ptr(verts), ptr(topo.tri), ptr(topo.v2f_ptr), ptr(topo.v2f_idx), V, topo.T,
                                              ptr(normals), ptr(raw), stream()), "vertex_normals_fwd")
        ctx.save_for_backward(verts, raw)
        ctx.topo = topo
        return normals

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        verts, raw = ctx.saved_tensors
        topo = ctx.topo
        dy = dy.contiguous()
        scratch = torch.empty_like(verts)
        grad = torch.empty_like(verts)
        with torch.cuda.device(verts.device):
            check(lib.fmhr_vertex_normals_bwd(ptr(verts), ptr(topo.tri), ptr(topo.v2f_ptr), ptr(topo.v2f_idx), ptr(raw),
                                              ptr(dy), verts.shape[0], topo.T, ptr(scratch), ptr(grad), stream()),
                  "vertex_normals_bwd")
        return grad, None


def get_normals(vertices, faces):
    """vertices [B,V,3] float32 CUDA, faces [F,3] -> normals [B,V,3].  A batch made by `.expand` of one mesh (what the
    reference passes, mesh_sfs_optim.py:262,265) is computed ONCE and expanded; distinct meshes loop over B."""
    _lib.require_cuda(vertices, faces)
    if vertices.dim() != 3 or vertices.shape[2] != 3:
        raise RuntimeError("fmhr_b200.get_normals: vertices must be [B,V,3]")
    B, V, _ = vertices.shape
    topo = topology_for(faces, V)
    if B == 1 or vertices.stride(0) == 0:
        return _VertexNormals.apply(vertices[0], topo).unsqueeze(0).expand(B, -1, -1)
    return torch.stack([_VertexNormals.apply(vertices[b], topo) for b in range(B)], 0)


class _Laplacian(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, topo):
        lib = _lib.load()
        x = x.contiguous()
        V, C = x.shape
        yhat = torch.empty_like(x)
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.fmhr_laplacian_fwd(ptr(x), ptr(topo.v2v_ptr), ptr(topo.v2v_idx), V, C, ptr(yhat), ptr(loss),
                                         stream()), "laplacian_fwd")
        ctx.save_for_backward(yhat)
        ctx.topo = topo
        return loss

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        (yhat,) = ctx.saved_tensors
        topo = ctx.topo
        V, C = yhat.shape
        grad = torch.empty_like(yhat)
        with torch.cuda.device(yhat.device):
            check(lib.fmhr_laplacian_bwd(ptr(yhat), ptr(topo.v2v_ptr), ptr(topo.v2v_idx), V, C, 1.0, ptr(grad), stream()),
                  "laplacian_bwd")
        return grad * dy, None


def laplacian_smoothing(verts, faces, method="uniform"):
    """sum_i ||(L verts)_i|| / V with the uniform graph Laplacian (the only method the reference's call sites use)."""
    if method != "uniform":
        raise RuntimeError("fmhr_b200.laplacian_smoothing: only method='uniform' is implemented "
                           "(mesh_sfs_optim.py:231,292,293 never select another)")
    _lib.require_cuda(verts, faces)
    if verts.dim() != 2 or not 1 <= verts.shape[1] <= 4:
        raise RuntimeError("fmhr_b200.laplacian_smoothing: verts must be [V,C], C<=4")
    return _Laplacian.apply(verts, topology_for(faces, verts.shape[0]))


def get_matrix(normal, degree=3):
    """SH basis rows [1, y, z, x, xy, yz, 2z^2-x^2-y^2, zx, x^2-y^2] (elementwise; plain tensor ops)."""
    x, y, z = normal[:, 0], normal[:, 1], normal[:, 2]
    cols = [torch.ones_like(x)]
    if degree > 1:
        cols += [y, z, x]
    if degree > 2:
        cols += [x * y, y * z, 2 * z * z - x * x - y * y, z * x, x * x - y * y]
    return torch.stack(cols, dim=1)


class _Radiance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coeff, normal):
        lib = _lib.load()
        coeff = coeff.contiguous()
        normal = normal.contiguous()
        n = normal.shape[0]
        rows = 1 if coeff.dim() == 1 else coeff.shape[0]
        out = torch.empty(n, dtype=torch.float32, device=normal.device)
        with torch.cuda.device(normal.device):
            check(lib.fmhr_sh_radiance_fwd(ptr(coeff), rows, ptr(normal), n, ptr(out), stream()), "sh_radiance_fwd")
        ctx.save_for_backward(coeff, normal)
        return out

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        coeff, normal = ctx.saved_tensors
        n = normal.shape[0]
        rows = 1 if coeff.dim() == 1 else coeff.shape[0]
        gc = torch.empty_like(coeff)
        gn = torch.empty_like(normal)
        dy = dy.contiguous()
        with torch.cuda.device(normal.device):
            check(lib.fmhr_sh_radiance_bwd(ptr(coeff), rows, ptr(normal), ptr(dy), n, ptr(gc), ptr(gn), stream()),
                  "sh_radiance_bwd")
        return gc, gn


def get_radiance(coeff, normal, degree=3):
    """coeff [9] or [n,9], normal [n,3] -> radiance [n]."""
    if degree != 3:
        raise RuntimeError("fmhr_b200.get_radiance: only degree 3 (9 coefficients) is implemented (conf/*.conf: degree = 3)")
    _lib.require_cuda(coeff, normal)
    if coeff.dim() == 2 and coeff.shape[0] != normal.shape[0]:
        raise RuntimeError("fmhr_b200.get_radiance: coeff rows must match normal rows")
    return _Radiance.apply(coeff, normal)


class _NCC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ref, src, mask):
        lib = _lib.load()
        ref, src, mask = ref.contiguous(), src.contiguous(), mask.contiguous()
        Nv, Np, Npx = src.shape
        out = torch.empty(Nv, Np, dtype=torch.float32, device=src.device)
        with torch.cuda.device(src.device):
            check(lib.fmhr_ncc_fwd(ptr(ref), ptr(src), ptr(mask), Nv, Np, Npx, ptr(out), stream()), "ncc_fwd")
        ctx.save_for_backward(ref, src, mask)
        return out

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        ref, src, mask = ctx.saved_tensors
        Nv, Np, Npx = src.shape
        grad = torch.empty_like(src)
        dy = dy.contiguous()
        with torch.cuda.device(src.device):
            check(lib.fmhr_ncc_bwd(ptr(ref), ptr(src), ptr(mask), ptr(dy), Nv, Np, Npx, ptr(grad), stream()), "ncc_bwd")
        return None, grad, None


def NCC(ref, src, ref_valid_mask, src_valid_mask):
    """ref [1,Np,Npx], src / src_valid_mask [Nv,Np,Npx] -> ncc [Nv,Np]; differentiable w.r.t. src (ref_valid_mask is
    ignored as in the reference, models/ncc_utils.py:4-35)."""
    _lib.require_cuda(ref, src, src_valid_mask)
    if src.dim() != 3 or ref.shape[-2:] != src.shape[-2:]:
        raise RuntimeError("fmhr_b200.NCC: ref must be [1,Np,Npx] and src [Nv,Np,Npx]")
    return _NCC.apply(ref, src, src_valid_mask).squeeze()

"""ORACLE (test infrastructure) — ctypes bindings + CPU autograd Functions over raster_oracle.cpp.

Mirrors the nvdiffrast.torch call surface the reference uses (mesh_sfs_optim.py:120,142-147,
212-219,267-287): RasterizeGLContext(), rasterize(), interpolate(), antialias() on CPU tensors.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle_raster.so")
    src = os.path.join(_HERE, "raster_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_raster.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


def rasterize_fwd(pos, tri, resolution, want_db=True, want_keys=False):
    pos = _f32(pos)
    tri = tri.detach().to(torch.int32).contiguous()
    N, V, _ = pos.shape
    T = tri.shape[0]
    H, W = resolution
    rast = torch.empty(N, H, W, 4, dtype=torch.float32)
    db = torch.empty(N, H, W, 4, dtype=torch.float32) if want_db else None
    keys = torch.empty(N, H, W, dtype=torch.int64) if want_keys else None
    rc = lib().orc_rasterize_fwd(_p(pos), _p(tri), N, V, T, H, W, _p(rast), _p(db), _p(keys))
    assert rc == 0
    return rast, db, keys


def clip_positions(vertices, w2cs, projs):
    """[n,V,4] clip positions in the fixed evaluation order the GPU path and the checker share (orc_clip_positions)."""
    v, w, p = _f32(vertices), _f32(w2cs), _f32(projs)
    n, V = w.shape[0], v.shape[0]
    out = torch.empty(n, V, 4, dtype=torch.float32)
    rc = lib().orc_clip_positions(_p(v), V, _p(w), _p(p), n, _p(out))
    assert rc == 0
    return out


def rasterize_bwd(pos, tri, rast, dy):
    pos = _f32(pos)
    tri = tri.detach().to(torch.int32).contiguous()
    N, V, _ = pos.shape
    _, H, W, _ = rast.shape
    g = torch.zeros_like(pos)
    lib().orc_rasterize_bwd(_p(pos), _p(tri), _p(_f32(rast)), _p(_f32(dy)), N, V, tri.shape[0], H, W, _p(g))
    return g


def interpolate_fwd(attr, rast, tri):
    attr = _f32(attr)
    rast = _f32(rast)
    tri = tri.detach().to(torch.int32).contiguous()
    NA, V, A = attr.shape
    N, H, W, _ = rast.shape
    out = torch.empty(N, H, W, A, dtype=torch.float32)
    lib().orc_interpolate_fwd(_p(attr), _p(rast), _p(tri), N, NA, V, tri.shape[0], H, W, A, _p(out))
    return out


def interpolate_bwd(attr, rast, tri, dy):
    attr = _f32(attr)
    rast = _f32(rast)
    tri = tri.detach().to(torch.int32).contiguous()
    NA, V, A = attr.shape
    N, H, W, _ = rast.shape
    ga = torch.zeros_like(attr)
    gr = torch.empty_like(rast)
    lib().orc_interpolate_bwd(_p(attr), _p(rast), _p(tri), _p(_f32(dy)), N, NA, V, tri.shape[0], H, W, A, _p(ga), _p(gr))
    return ga, gr


def antialias_topology(tri):
    tri = tri.detach().to(torch.int32).contiguous()
    opp = torch.empty_like(tri)
    lib().orc_antialias_topology(_p(tri), tri.shape[0], _p(opp))
    return opp


def antialias_fwd(color, rast, pos, tri, opp=None, want_items=False):
    color = _f32(color)
    rast = _f32(rast)
    pos = _f32(pos)
    tri = tri.detach().to(torch.int32).contiguous()
    if opp is None:
        opp = antialias_topology(tri)
    N, H, W, C = color.shape
    V = pos.shape[1]
    out = torch.empty_like(color)
    max_items = 2 * N * H * W if want_items else 0
    items = torch.zeros(max(max_items, 1), 8, dtype=torch.int32) if want_items else None
    cnt = ctypes.c_int(0)
    lib().orc_antialias_fwd(_p(color), _p(rast), _p(pos), _p(tri), _p(opp), N, H, W, C, V, tri.shape[0], _p(out),
                            _p(items), max_items, ctypes.byref(cnt))
    if want_items:
        return out, items[: cnt.value].clone()
    return out


def antialias_bwd(color, rast, pos, tri, dy, opp=None):
    color = _f32(color)
    rast = _f32(rast)
    pos = _f32(pos)
    tri = tri.detach().to(torch.int32).contiguous()
    if opp is None:
        opp = antialias_topology(tri)
    N, H, W, C = color.shape
    V = pos.shape[1]
    gc = torch.empty_like(color)
    gp = torch.zeros_like(pos)
    lib().orc_antialias_bwd(_p(color), _p(rast), _p(pos), _p(tri), _p(opp), _p(_f32(dy)), N, H, W, C, V, tri.shape[0],
                            _p(gc), _p(gp))
    return gc, gp


# ----------------------------------------------------------------------------------------------
# nvdiffrast.torch-shaped surface on CPU tensors (what the oracle HAM step calls)
# ----------------------------------------------------------------------------------------------
class RasterizeGLContext:
    def __init__(self, *a, **k):
        pass


RasterizeCudaContext = RasterizeGLContext


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, tri, resolution):
        rast, db, _ = rasterize_fwd(pos, tri, resolution, want_db=True)
        ctx.save_for_backward(pos, tri, rast)
        return rast, db

    @staticmethod
    def backward(ctx, dy, ddb):
        pos, tri, rast = ctx.saved_tensors
        return rasterize_bwd(pos, tri, rast, dy), None, None


def rasterize(glctx, pos, tri, resolution, ranges=None, grad_db=True):
    assert ranges is None and pos.dim() == 3
    return _Rasterize.apply(pos, tri, tuple(resolution))


class _Interpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attr, rast, tri):
        out = interpolate_fwd(attr, rast, tri)
        ctx.save_for_backward(attr, rast, tri)
        return out

    @staticmethod
    def backward(ctx, dy):
        attr, rast, tri = ctx.saved_tensors
        ga, gr = interpolate_bwd(attr, rast, tri, dy)
        return ga, gr, None


def interpolate(attr, rast, tri, rast_db=None, diff_attrs=None):
    assert rast_db is None and diff_attrs is None
    out = _Interpolate.apply(attr, rast, tri)
    return out, torch.empty(*out.shape[:-1], 0)


class _Antialias(torch.autograd.Function):
    @staticmethod
    def forward(ctx, color, rast, pos, tri):
        opp = antialias_topology(tri)
        out = antialias_fwd(color, rast, pos, tri, opp)
        ctx.save_for_backward(color, rast, pos, tri, opp)
        return out

    @staticmethod
    def backward(ctx, dy):
        color, rast, pos, tri, opp = ctx.saved_tensors
        gc, gp = antialias_bwd(color, rast, pos, tri, dy, opp)
        return gc, None, gp, None


def antialias(color, rast, pos, tri, topology_hash=None, pos_gradient_boost=1.0):
    assert pos_gradient_boost == 1.0
    return _Antialias.apply(color, rast, pos, tri)

"""`nvdiffrast.torch` call surface on top of libfmhr_b200.so.

The reference imports `nvdiffrast.torch as dr` (mesh_sfs_optim.py:14) and uses exactly four names:
`dr.RasterizeGLContext()` (:120), `dr.rasterize(glctx, pos, tri, resolution=...)` (:142,212,267),
`dr.interpolate(attr, rast, tri)` (:143,214,269) and `dr.antialias(color, rast, pos, tri)`
(:146-147,217-219,274,287).  Each op is a `torch.autograd.Function` over CUDA tensors whose forward and
backward are one C-ABI call; the repo-root package `nvdiffrast/` re-exports this module so the reference
scripts run unchanged.  Keyword signatures follow upstream.
"""
import collections
import os

import torch

from . import _lib
from ._lib import check, ptr, stream


# ------------------------------------------------------------------------------------------------
# per-mesh topology cache (replaces upstream's per-call topology hash)
# ------------------------------------------------------------------------------------------------
class Topology:
    """opp / vertex->face CSR / vertex->vertex CSR of one triangle tensor (fmhr_mesh_topology_build)."""

    def __init__(self, tri, n_verts):
        lib = _lib.load()
        tri = tri.contiguous()
        T = tri.shape[0]
        V = int(n_verts)
        dev = tri.device
        mx = int(tri.max().item()) if T > 0 else -1
        mn = int(tri.min().item()) if T > 0 else 0
        if mn < 0 or mx >= V:
            raise RuntimeError("fmhr_b200: triangle tensor references vertex %d outside [0,%d)" % (mx if mx >= V else mn, V))
        self.V, self.T = V, T
        self.opp = torch.empty(T, 3, dtype=torch.int32, device=dev)
        self.v2f_ptr = torch.empty(V + 1, dtype=torch.int32, device=dev)
        self.v2f_idx = torch.empty(3 * T, dtype=torch.int32, device=dev)
        self.v2v_ptr = torch.empty(V + 1, dtype=torch.int32, device=dev)
        v2v_idx = torch.empty(6 * T, dtype=torch.int32, device=dev)
        ws_bytes = lib.fmhr_mesh_topology_workspace_bytes(V, T)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        n_dir = _lib.c_i(0)
        with torch.cuda.device(dev):
            check(lib.fmhr_mesh_topology_build(ptr(tri), V, T, ptr(self.opp), ptr(self.v2f_ptr), ptr(self.v2f_idx),
                                               ptr(self.v2v_ptr), ptr(v2v_idx), _lib.ctypes.byref(n_dir), ptr(ws),
                                               ws_bytes, stream()), "mesh_topology_build")
        self.n_dir_edges = n_dir.value
        self.v2v_idx = v2v_idx[: self.n_dir_edges].clone()
        self.tri = tri
        self.v2f_nbr = torch.empty(3 * T, 2, dtype=torch.int32, device=dev)
        self.inv_deg = torch.empty(V, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.fmhr_mesh_topology_derive(ptr(tri), ptr(self.v2f_idx), ptr(self.v2v_ptr), V, T, ptr(self.v2f_nbr),
                                                ptr(self.inv_deg), stream()), "mesh_topology_derive")


_TOPO_CACHE = collections.OrderedDict()


def get_topology(tri, n_verts):
    key = (tri.data_ptr(), tri._version, tuple(tri.shape), int(n_verts), str(tri.device))
    topo = _TOPO_CACHE.get(key)
    if topo is None:
        topo = Topology(tri, n_verts)
        _TOPO_CACHE[key] = topo
        while len(_TOPO_CACHE) > 8:
            _TOPO_CACHE.popitem(last=False)
    else:
        _TOPO_CACHE.move_to_end(key)
    return topo


class Meshlets:
    """Meshlets of one triangle tensor for the stand-alone rasteriser (fmhr_rasterize_fwd_meshlets): faces ordered along a
    Morton curve of their centroids in the NDC of the first view, cut into groups of <= 1024 triangles / 1024 distinct
    vertices (fmhr_meshlets_build_host, once per mesh).  `ok` is False for meshes the builder cannot take (indices outside
    [0,V), no triangles): rasterize then uses the one-thread-per-triangle kernel, which skips such triangles."""
    TRIS = 1024

    def __init__(self, tri, n_verts, pos0):
        import ctypes

        import numpy as np
        lib = _lib.load()
        self.ok = False
        self.tri = tri  # keeps the storage alive: the cache key (data_ptr, version) cannot be recycled by the allocator
        T, V = tri.shape[0], int(n_verts)
        if T == 0:
            return
        tri_h = np.ascontiguousarray(tri.detach().cpu().numpy(), dtype=np.int32)
        if tri_h.min() < 0 or tri_h.max() >= V:
            return
        w = pos0[:, 3:4]
        ndc = torch.nan_to_num(pos0[:, :3] / torch.where(w.abs() > 1e-12, w, torch.ones_like(w)), 0.0, 0.0, 0.0)
        verts = np.ascontiguousarray(ndc.clamp(-4.0, 4.0).detach().cpu().numpy(), dtype=np.float32)
        hp = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)
        nm, nr, mv = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        call = lambda a, b, c: lib.fmhr_meshlets_build_host(hp(tri_h), hp(verts), V, T, self.TRIS, ctypes.byref(nm),
                                                            ctypes.byref(nr), ctypes.byref(mv), hp(a), hp(b), hp(c))
        if call(None, None, None) != 0 or nm.value <= 0:
            return
        vptr = np.zeros(nm.value + 1, dtype=np.int32)
        vrefs = np.zeros(max(nr.value, 1), dtype=np.int32)
        tri2 = np.zeros((nm.value * self.TRIS, 2), dtype=np.uint32)
        if call(vptr, vrefs, tri2) != 0:
            return
        dev = tri.device
        self.vptr = torch.from_numpy(vptr).to(dev)
        self.verts = torch.from_numpy(vrefs).to(dev)
        self.tri2 = torch.from_numpy(tri2.view(np.int32)).to(dev)
        self.n, self.max_verts = nm.value, mv.value
        self.ok = True


_MESHLET_CACHE = collections.OrderedDict()


def get_meshlets(tri, n_verts, pos):
    key = (tri.data_ptr(), tri._version, tuple(tri.shape), int(n_verts), str(tri.device))
    ml = _MESHLET_CACHE.get(key)
    if ml is None:
        ml = Meshlets(tri, n_verts, pos[0].detach())
        _MESHLET_CACHE[key] = ml
        while len(_MESHLET_CACHE) > 8:
            _MESHLET_CACHE.popitem(last=False)
    else:
        _MESHLET_CACHE.move_to_end(key)
    return ml


# ------------------------------------------------------------------------------------------------
# contexts
# ------------------------------------------------------------------------------------------------
class RasterizeCudaContext:
    """Opaque rasteriser context.  Owns the z-buffer scratch so steady-state calls do not allocate."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("fmhr_b200: no CUDA device - the rasteriser has no CPU path")
        _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._ws = None
        self._ws_clean = False   # every byte 0xFF: the meshlet rasteriser then needs no clear pass and leaves it so
        self._tile_bits = None
        self.use_meshlets = os.environ.get("FMHR_RASTER", "meshlets") != "v1"

    def workspace(self, nbytes, device, clean=False):
        """The z-buffer scratch.  clean=True: make sure it is all-0xFF (filled once at allocation; the meshlet rasteriser
        restores that state itself); clean=False: the caller clears / dirties it."""
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != device:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._ws_clean = False
        if clean and not self._ws_clean:
            self._ws.fill_(255)
        self._ws_clean = clean
        return self._ws

    def tile_bits(self, words, device):
        if self._tile_bits is None or self._tile_bits.numel() < words or self._tile_bits.device != device:
            self._tile_bits = torch.empty(max(words, 1), dtype=torch.int32, device=device)
        return self._tile_bits


class RasterizeGLContext(RasterizeCudaContext):
    """Name kept for the reference (`dr.RasterizeGLContext()`, mesh_sfs_optim.py:120); there is no OpenGL here."""

    def __init__(self, output_db=True, mode='automatic', device=None):
        super().__init__(device)


def _f32c(t, name):
    if not t.is_cuda:
        raise RuntimeError("fmhr_b200.%s: expected a CUDA tensor" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("fmhr_b200.%s: expected float32, got %s" % (name, t.dtype))
    return t.contiguous()


def _tri_c(tri):
    if not tri.is_cuda or tri.dtype != torch.int32 or tri.dim() != 2 or tri.shape[1] != 3:
        raise RuntimeError("fmhr_b200: tri must be a CUDA int32 tensor of shape [T,3]")
    return tri.contiguous()


# ------------------------------------------------------------------------------------------------
# rasterize
# ------------------------------------------------------------------------------------------------
class _RasterizeFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, glctx, pos, tri, resolution, grad_db):
        lib = _lib.load()
        pos = _f32c(pos, "rasterize(pos)")
        tri = _tri_c(tri)
        if pos.dim() != 3 or pos.shape[2] != 4:
            raise RuntimeError("fmhr_b200.rasterize: pos must be [N,V,4] (instanced mode); range mode is not supported")
        N, V, _ = pos.shape
        H, W = int(resolution[0]), int(resolution[1])
        rast = torch.empty(N, H, W, 4, dtype=torch.float32, device=pos.device)
        rast_db = torch.empty(N, H, W, 4, dtype=torch.float32, device=pos.device) if grad_db else None
        nbytes = lib.fmhr_rasterize_workspace_bytes(N, H, W)
        ml = get_meshlets(tri, V, pos) if (glctx.use_meshlets and N <= 65535) else None
        with torch.cuda.device(pos.device):
            if ml is not None and ml.ok:
                ws = glctx.workspace(nbytes, pos.device, clean=True)
                bits = glctx.tile_bits(lib.fmhr_rasterize_tile_words(N, H, W), pos.device)
                glctx._ws_clean = False  # until the call below has been issued successfully
                check(lib.fmhr_rasterize_fwd_meshlets(ptr(pos), ptr(tri), N, V, tri.shape[0], H, W, ptr(rast), ptr(rast_db),
                                                      ptr(ml.vptr), ptr(ml.verts), ptr(ml.tri2), ml.n, ml.TRIS, ml.max_verts,
                                                      ptr(ws), ws.numel(), 1, ptr(bits), stream()), "rasterize_fwd_meshlets")
                glctx._ws_clean = True
            else:
                ws = glctx.workspace(nbytes, pos.device)
                check(lib.fmhr_rasterize_fwd(ptr(pos), ptr(tri), N, V, tri.shape[0], H, W, ptr(rast), ptr(rast_db), ptr(ws),
                                             ws.numel(), stream()), "rasterize_fwd")
        ctx.save_for_backward(pos, tri, rast)
        if rast_db is None:
            rast_db = torch.zeros(N, H, W, 0, dtype=torch.float32, device=pos.device)
        ctx.mark_non_differentiable(rast_db) if not grad_db else None
        return rast, rast_db

    @staticmethod
    def backward(ctx, dy, ddb):
        lib = _lib.load()
        pos, tri, rast = ctx.saved_tensors
        N, V, _ = pos.shape
        _, H, W, _ = rast.shape
        grad_pos = torch.empty_like(pos)
        dy = dy.contiguous()
        with torch.cuda.device(pos.device):
            check(lib.fmhr_rasterize_bwd(ptr(pos), ptr(tri), ptr(rast), ptr(dy), N, V, tri.shape[0], H, W,
                                         ptr(grad_pos), stream()), "rasterize_bwd")
        return None, grad_pos, None, None, None


def rasterize(glctx, pos, tri, resolution, ranges=None, grad_db=True):
    """-> (rast[N,H,W,4] = (u, v, z/w, triangle_id+1), rast_db[N,H,W,4]).

    Limitation (rule 1): a triangle with ANY vertex at w <= 0, outside -w <= z <= w, or outside the
    +-16384 px guard band is dropped as a whole - upstream nvdiffrast clips such triangles against the view volume.  The
    reference's cameras never produce one (projection z' = -0.1, w' = z_cam with the hand ~1.9 units away,
    get_data.py:66-73), but a caller whose geometry crosses the near plane gets holes, not clipped triangles
    (pinned by tests/test_gpu_ops.py::test_rasterize_drops_triangles_crossing_the_near_plane)."""
    if not isinstance(glctx, RasterizeCudaContext):
        raise RuntimeError("fmhr_b200.rasterize: glctx must be a RasterizeGLContext / RasterizeCudaContext")
    if ranges is not None:
        raise RuntimeError("fmhr_b200.rasterize: range mode (`ranges`) is not supported; the reference never uses it")
    return _RasterizeFunc.apply(glctx, pos, tri, tuple(resolution), bool(grad_db))


# ------------------------------------------------------------------------------------------------
# interpolate
# ------------------------------------------------------------------------------------------------
class _InterpolateFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attr, rast, tri):
        lib = _lib.load()
        attr = _f32c(attr, "interpolate(attr)")
        rast = _f32c(rast, "interpolate(rast)")
        tri = _tri_c(tri)
        if attr.dim() != 3:
            raise RuntimeError("fmhr_b200.interpolate: attr must be [N,V,A] or [1,V,A] (instanced mode)")
        NA, V, A = attr.shape
        N, H, W, _ = rast.shape
        if NA != N and NA != 1:
            raise RuntimeError("fmhr_b200.interpolate: attr batch %d does not match rast batch %d" % (NA, N))
        out = torch.empty(N, H, W, A, dtype=torch.float32, device=attr.device)
        with torch.cuda.device(attr.device):
            check(lib.fmhr_interpolate_fwd(ptr(attr), ptr(rast), ptr(tri), N, NA, V, tri.shape[0], H, W, A, ptr(out),
                                           stream()), "interpolate_fwd")
        ctx.save_for_backward(attr, rast, tri)
        return out

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        attr, rast, tri = ctx.saved_tensors
        NA, V, A = attr.shape
        N, H, W, _ = rast.shape
        grad_attr = torch.empty_like(attr)
        grad_rast = torch.empty_like(rast)
        dy = dy.contiguous()
        with torch.cuda.device(attr.device):
            check(lib.fmhr_interpolate_bwd(ptr(attr), ptr(rast), ptr(tri), ptr(dy), N, NA, V, tri.shape[0], H, W, A,
                                           ptr(grad_attr), ptr(grad_rast), stream()), "interpolate_bwd")
        return grad_attr, grad_rast, None


def interpolate(attr, rast, tri, rast_db=None, diff_attrs=None):
    """-> (out[N,H,W,A], out_da[N,H,W,0]).  Pixel-differential attributes (`rast_db`/`diff_attrs`) are not
    supported; the reference never requests them."""
    if diff_attrs is not None:
        raise RuntimeError("fmhr_b200.interpolate: diff_attrs is not supported; the reference never uses it")
    out = _InterpolateFunc.apply(attr, rast, tri)
    return out, torch.empty(*out.shape[:-1], 0, dtype=out.dtype, device=out.device)


# ------------------------------------------------------------------------------------------------
# antialias
# ------------------------------------------------------------------------------------------------
class _AntialiasFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, color, rast, pos, tri, topo, boost):
        lib = _lib.load()
        color = _f32c(color, "antialias(color)")
        rast = _f32c(rast, "antialias(rast)")
        pos = _f32c(pos, "antialias(pos)")
        tri = _tri_c(tri)
        if pos.dim() != 3:
            raise RuntimeError("fmhr_b200.antialias: pos must be [N,V,4] (instanced mode)")
        N, H, W, C = color.shape
        V = pos.shape[1]
        out = torch.empty_like(color)
        with torch.cuda.device(color.device):
            check(lib.fmhr_antialias_fwd(ptr(color), ptr(rast), ptr(pos), ptr(tri), ptr(topo.opp), N, H, W, C, V,
                                         tri.shape[0], ptr(out), stream()), "antialias_fwd")
        ctx.save_for_backward(color, rast, pos, tri)
        ctx.topo = topo
        ctx.boost = boost
        return out

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        color, rast, pos, tri = ctx.saved_tensors
        N, H, W, C = color.shape
        V = pos.shape[1]
        grad_color = torch.empty_like(color)
        need_pos = ctx.needs_input_grad[2]
        grad_pos = torch.empty_like(pos) if need_pos else None
        dy = dy.contiguous()
        with torch.cuda.device(color.device):
            check(lib.fmhr_antialias_bwd(ptr(color), ptr(rast), ptr(pos), ptr(tri), ptr(ctx.topo.opp), ptr(dy), N, H, W,
                                         C, V, tri.shape[0], ptr(grad_color), ptr(grad_pos), stream()), "antialias_bwd")
        if need_pos and ctx.boost != 1.0:
            grad_pos = grad_pos * ctx.boost
        return grad_color, None, grad_pos, None, None, None


def get_antialias_topology_hash(tri, n_verts):
    """Counterpart of upstream's antialias_construct_topology_hash(tri)."""
    return get_topology(_tri_c(tri), n_verts)


def antialias(color, rast, pos, tri, topology_hash=None, pos_gradient_boost=1.0):
    """-> color with silhouette edges blended.  The per-mesh topology is cached on the `tri` tensor
    (data_ptr, version), so repeated calls on a static mesh do not rebuild it."""
    topo = topology_hash if isinstance(topology_hash, Topology) else get_topology(_tri_c(tri), pos.shape[-2])
    return _AntialiasFunc.apply(color, rast, pos, tri, topo, float(pos_gradient_boost))

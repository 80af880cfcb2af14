"""GPU parity tests of the nvdiffrast-shaped operators (through the C ABI) against the CPU oracle.

Bars (BASELINE.json north_star): triangle id / coverage bit-exact; barycentrics, images <= 1e-5 relative;
gradients <= 1e-4 relative (atomics make the summation order non-deterministic).
"""
import numpy as np
import pytest
import torch

from fmhr_b200 import synth
from oracle import raster as orc

pytestmark = pytest.mark.gpu


def _scene(workload="small", seed=0):
    v, f = synth.hand_mesh(synth.WORKLOADS[workload]["subdiv"], 1, seed=seed)
    wl = synth.WORKLOADS[workload]
    w2c, proj = synth.make_cameras(wl["n"], wl["H"], wl["W"], v.mean(0).astype(np.float64),
                                   extent=float(v[:, 1].max() - v[:, 1].min()))
    vt = torch.tensor(v)
    vh = torch.cat([vt, torch.ones_like(vt[:, :1])], 1)[None].expand(wl["n"], -1, -1)
    pos = torch.einsum('ijk,ikl->ijl', torch.einsum('ijk,ikl->ijl', vh, torch.tensor(w2c)), torch.tensor(proj)).contiguous()
    return pos, torch.tensor(f), wl["H"], wl["W"]


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


@pytest.fixture(scope="module")
def dr():
    from fmhr_b200 import dr as _dr
    return _dr


@pytest.mark.parametrize("workload", ["tiny", "small"])
def test_rasterize_bit_exact(dr, workload):
    pos, tri, H, W = _scene(workload)
    ref, ref_db, _ = orc.rasterize_fwd(pos, tri, (H, W))
    ctx = dr.RasterizeGLContext()
    rast, db = dr.rasterize(ctx, pos.cuda(), tri.cuda(), resolution=(H, W))
    rast, db = rast.cpu(), db.cpu()
    assert torch.equal(rast[..., 3], ref[..., 3]), "triangle id / coverage must be bit-exact"
    assert torch.equal(rast[..., 2], ref[..., 2]), "depth must be bit-exact (it decides the winner)"
    assert torch.equal(rast[..., :2], ref[..., :2]), "barycentrics use the same un-contracted fp32 ops"
    assert (ref[..., 3] > 0).float().mean() > 0.02
    assert torch.allclose(db, ref_db, rtol=1e-3, atol=1e-4)


def test_rasterize_meshlet_and_v1_paths_agree(dr):
    """dr.rasterize runs on the meshlet coverage kernel with a self-cleaning z-buffer (fmhr_rasterize_fwd_meshlets);
    the first-generation kernel (fmhr_rasterize_fwd) stays as the fallback for meshes the meshlet builder rejects.
    Both must give the oracle's result bit for bit - also on REPEATED calls with changing sizes on one context (the
    z-buffer is only cleared once, every call has to leave it clean) and after a v1 call dirtied the shared scratch."""
    ctx = dr.RasterizeGLContext()
    v1 = dr.RasterizeGLContext()
    v1.use_meshlets = False
    assert ctx.use_meshlets
    for workload in ("small", "tiny", "small", "coarse"):
        pos, tri, H, W = _scene(workload)
        ref, _, _ = orc.rasterize_fwd(pos, tri, (H, W), want_db=False)
        for rep in range(2):
            a, _ = dr.rasterize(ctx, pos.cuda(), tri.cuda(), resolution=(H, W))
            assert torch.equal(a.cpu(), ref), (workload, rep)
        b, _ = dr.rasterize(v1, pos.cuda(), tri.cuda(), resolution=(H, W), grad_db=False)
        assert torch.equal(b.cpu(), ref)
        assert ctx._ws_clean and bool((ctx._ws == 255).all()), "the meshlet path must leave the z-buffer clean"
    # a context that alternates between the two kernels
    pos, tri, H, W = _scene("tiny")
    ref, _, _ = orc.rasterize_fwd(pos, tri, (H, W), want_db=False)
    for use in (False, True, False, True, True):
        ctx.use_meshlets = use
        a, _ = dr.rasterize(ctx, pos.cuda(), tri.cuda(), resolution=(H, W))
        assert torch.equal(a.cpu(), ref), use


def test_rasterize_odd_sizes_and_culling(dr):
    # W=334 is not a multiple of 8; some triangles behind the near plane / off-screen / degenerate
    g = torch.Generator().manual_seed(3)
    V, T, N, H, W = 300, 500, 3, 77, 334
    pos = torch.randn(N, V, 4, generator=g)
    pos[..., 3] = pos[..., 3].abs() * 2 + 0.05
    pos[..., 2] = pos[..., 2] * 0.5
    pos[0, :20, 3] = -1.0  # behind the camera
    pos[1, 20:30, 0] = 1e9  # outside the guard band
    tri = torch.randint(0, V, (T, 3), generator=g, dtype=torch.int32)
    tri[:10, 1] = tri[:10, 0]  # degenerate
    ref, _, _ = orc.rasterize_fwd(pos, tri, (H, W), want_db=False)
    rast, _ = dr.rasterize(dr.RasterizeGLContext(), pos.cuda(), tri.cuda(), resolution=(H, W))
    assert torch.equal(rast.cpu(), ref)
    assert (ref[..., 3] > 0).any()


def test_rasterize_known_answers(dr):
    # two triangles forming the pixel-aligned quad [2,6]x[1,5] of an 8x8 image: every pixel centre inside is covered
    # exactly once (watertight shared edge), nearer triangle wins, equal depth -> lower index wins.
    H = W = 8

    def ndc(x, y):
        return [2.0 * x / W - 1.0, 2.0 * y / H - 1.0]

    quad = [ndc(2, 1), ndc(6, 1), ndc(6, 5), ndc(2, 5)]
    pos = torch.tensor([[q + [0.0, 1.0] for q in quad] + [q + [-0.5, 1.0] for q in quad]], dtype=torch.float32)
    tri = torch.tensor([[0, 1, 2], [0, 2, 3], [4, 5, 6], [0, 1, 2]], dtype=torch.int32)
    ctx = dr.RasterizeGLContext()
    r, _ = dr.rasterize(ctx, pos.cuda(), tri[:2].cuda(), resolution=(H, W))
    ids = r[0, :, :, 3].cpu()
    assert (ids[1:5, 2:6] > 0).all() and int((ids > 0).sum()) == 16
    r1, _ = dr.rasterize(ctx, pos.cuda(), tri[:1].cuda(), resolution=(H, W))
    r2, _ = dr.rasterize(ctx, pos.cuda(), tri[1:2].cuda(), resolution=(H, W))
    assert int(((r1[..., 3] > 0) & (r2[..., 3] > 0)).sum()) == 0, "shared edge covered twice"
    assert int(((r1[..., 3] > 0) | (r2[..., 3] > 0)).sum()) == 16
    # triangle 2 (z=-0.5) is nearer than triangle 0 (z=0) where they overlap; triangle 3 duplicates 0 and loses the tie
    r3, _ = dr.rasterize(ctx, pos.cuda(), tri.cuda(), resolution=(H, W))
    ids3 = r3[0, :, :, 3].cpu()
    cov0 = r1[0, :, :, 3].cpu() > 0
    assert (ids3[cov0] == 3).all()
    r4, _ = dr.rasterize(ctx, pos.cuda(), tri[[0, 3]].cuda(), resolution=(H, W))
    assert (r4[0, :, :, 3].cpu()[cov0] == 1).all()
    for rr in (r, r3):
        ref, _, _ = orc.rasterize_fwd(pos, tri[:2] if rr is r else tri, (H, W), want_db=False)
        assert torch.equal(rr.cpu(), ref)


def test_rasterize_drops_triangles_crossing_the_near_plane(dr):
    """Documented limitation of the shim (dr.rasterize docstring, DESIGN.md section 2 rule 1): a triangle with one vertex
    behind the near plane (z < -w) or the eye (w <= 0) is dropped as a whole, where upstream nvdiffrast would clip it.
    The same triangle pulled fully inside the volume is drawn; the oracle follows the same rule."""
    H = W = 16
    inside = [[-0.5, -0.5, 0.0, 1.0], [0.5, -0.5, 0.0, 1.0], [0.0, 0.5, 0.0, 1.0]]
    for bad in ([0.0, 0.5, -2.0, 1.0], [0.0, 0.5, 0.0, -0.5]):   # z < -w (crosses the near plane); w <= 0 (behind the eye)
        pos = torch.tensor([inside[:2] + [bad]], dtype=torch.float32)
        tri = torch.tensor([[0, 1, 2]], dtype=torch.int32)
        rast, _ = dr.rasterize(dr.RasterizeGLContext(), pos.cuda(), tri.cuda(), resolution=(H, W))
        assert int((rast[..., 3] > 0).sum()) == 0
        ref, _, _ = orc.rasterize_fwd(pos, tri, (H, W), want_db=False)
        assert torch.equal(rast.cpu(), ref)
    rast, _ = dr.rasterize(dr.RasterizeGLContext(), torch.tensor([inside]).cuda(), tri.cuda(), resolution=(H, W))
    assert int((rast[..., 3] > 0).sum()) > 20


def test_rasterize_backward(dr):
    pos, tri, H, W = _scene("small")
    rast_ref, _, _ = orc.rasterize_fwd(pos, tri, (H, W))
    g = torch.Generator().manual_seed(0)
    dy = torch.randn(rast_ref.shape, generator=g)
    ref = orc.rasterize_bwd(pos, tri, rast_ref, dy)
    p = pos.cuda().requires_grad_(True)
    rast, _ = dr.rasterize(dr.RasterizeGLContext(), p, tri.cuda(), resolution=(H, W))
    rast.backward(dy.cuda())
    assert _rel(p.grad.cpu(), ref) < 1e-4


@pytest.mark.parametrize("A,bcast", [(1, False), (3, False), (4, False), (5, True), (6, True), (7, False), (7, True),
                                     (10, False), (13, False), (30, True)])
def test_interpolate(dr, A, bcast):
    pos, tri, H, W = _scene("tiny")
    N, V, _ = pos.shape
    rast, _, _ = orc.rasterize_fwd(pos, tri, (H, W))
    g = torch.Generator().manual_seed(A)
    attr = torch.randn(1 if bcast else N, V, A, generator=g)
    dy = torch.randn(N, H, W, A, generator=g)
    ref = orc.interpolate_fwd(attr, rast, tri)
    ref_ga, ref_gr = orc.interpolate_bwd(attr, rast, tri, dy)
    a = attr.cuda().requires_grad_(True)
    r = rast.cuda().requires_grad_(True)
    out, out_da = dr.interpolate(a, r, tri.cuda())
    assert out_da.shape == (N, H, W, 0)
    assert torch.allclose(out.cpu(), ref, rtol=1e-5, atol=1e-6)
    out.backward(dy.cuda())
    assert _rel(a.grad.cpu(), ref_ga) < 1e-4
    assert _rel(r.grad.cpu(), ref_gr) < 1e-4
    assert torch.equal(r.grad[..., 2:].cpu(), torch.zeros(N, H, W, 2))


@pytest.mark.parametrize("A", [1, 3, 7, 30])
def test_interpolate_ragged_tail(dr, A):
    """The forward writes one float4 of the flat [N*H*W*A] output per thread: a plane whose element count is not a
    multiple of four (odd crop of one view) exercises the scalar tail, and quads that straddle pixel boundaries."""
    pos, tri, H, W = _scene("tiny")
    rast, _, _ = orc.rasterize_fwd(pos, tri, (H, W))
    covered = (rast[0, :, :, 3] > 0).nonzero()
    y0, x0 = [int(v) for v in covered[len(covered) // 2]]
    y0, x0 = max(0, min(y0 - 2, H - 5)), max(0, min(x0 - 3, W - 7))
    crop = rast[:1, y0:y0 + 5, x0:x0 + 7].contiguous()  # 35 pixels, some covered
    assert (crop[..., 3] > 0).any() and (35 * A) % 4 != 0
    attr = torch.randn(1, pos.shape[1], A, generator=torch.Generator().manual_seed(100 + A))
    ref = orc.interpolate_fwd(attr, crop, tri)
    out, _ = dr.interpolate(attr.cuda(), crop.cuda(), tri.cuda())
    assert torch.allclose(out.cpu(), ref, rtol=1e-5, atol=1e-6)
    empty = torch.zeros(1, 3, 3, 4)
    out0, _ = dr.interpolate(attr.cuda(), empty.cuda(), tri.cuda())
    assert torch.equal(out0.cpu(), torch.zeros(1, 3, 3, A))


def test_topology_matches_oracle(dr):
    _, tri, _, _ = _scene("small")
    V = int(tri.max()) + 1
    topo = dr.get_antialias_topology_hash(tri.cuda(), V)
    assert torch.equal(topo.opp.cpu(), orc.antialias_topology(tri))
    # CSR sanity: every (vertex, face, corner) incidence once; neighbour lists sorted, symmetric, no self loops
    ptr, idx = topo.v2f_ptr.cpu().long(), topo.v2f_idx.cpu().long()
    assert ptr[0] == 0 and ptr[-1] == 3 * tri.shape[0]
    vv = torch.repeat_interleave(torch.arange(V), ptr[1:] - ptr[:-1])
    assert torch.equal(tri.long()[idx >> 2, idx & 3], vv)
    p2, i2 = topo.v2v_ptr.cpu().long(), topo.v2v_idx.cpu().long()
    rows = torch.repeat_interleave(torch.arange(V), p2[1:] - p2[:-1])
    assert (rows != i2).all()
    e = set(zip(rows.tolist(), i2.tolist()))
    assert all((b, a) in e for a, b in e)
    from fmhr_b200.synth import unique_edges
    ue, _ = unique_edges(tri.numpy().astype(np.int64), V)
    assert len(e) == 2 * ue.shape[0]


@pytest.mark.parametrize("C", [1, 3])
def test_antialias(dr, C):
    pos, tri, H, W = _scene("coarse")
    N = pos.shape[0]
    rast, _, _ = orc.rasterize_fwd(pos, tri, (H, W))
    g = torch.Generator().manual_seed(C)
    color = torch.rand(N, H, W, C, generator=g) * (rast[..., 3:] > 0)
    if C == 1:
        color = (rast[..., 3:] > 0).float()
    dy = torch.randn(N, H, W, C, generator=g)
    ref, items = orc.antialias_fwd(color, rast, pos, tri, want_items=True)
    assert items.shape[0] > 100, "the scene must exercise silhouette pairs"
    ref_gc, ref_gp = orc.antialias_bwd(color, rast, pos, tri, dy)
    c = color.cuda().requires_grad_(True)
    p = pos.cuda().requires_grad_(True)
    out = dr.antialias(c, rast.cuda(), p, tri.cuda())
    assert torch.allclose(out.cpu(), ref, rtol=1e-5, atol=1e-6)
    assert (ref != color).any()
    out.backward(dy.cuda())
    assert _rel(c.grad.cpu(), ref_gc) < 1e-5
    assert _rel(p.grad.cpu(), ref_gp) < 1e-4


def test_errors_are_loud(dr):
    ctx = dr.RasterizeGLContext()
    pos = torch.zeros(1, 4, 4, device="cuda")
    tri = torch.zeros(1, 3, dtype=torch.int32, device="cuda")
    with pytest.raises(RuntimeError):
        dr.rasterize(ctx, pos.cpu(), tri, resolution=(8, 8))
    with pytest.raises(RuntimeError):
        dr.rasterize(ctx, pos, tri.long(), resolution=(8, 8))
    with pytest.raises(RuntimeError):
        dr.rasterize(ctx, pos[0], tri, resolution=(8, 8), ranges=torch.zeros(1, 2, dtype=torch.int32))
    # empty triangle list renders an empty image
    r, _ = dr.rasterize(ctx, pos, tri[:0], resolution=(8, 8))
    assert float(r.abs().sum()) == 0.0


def test_reference_import_path():
    """`import nvdiffrast.torch as dr` (mesh_sfs_optim.py:14) resolves to this package."""
    import nvdiffrast.torch as ndr
    from fmhr_b200 import dr as fdr
    assert ndr.rasterize is fdr.rasterize and ndr.antialias is fdr.antialias
    assert ndr.RasterizeGLContext is fdr.RasterizeGLContext


def test_mlp_forward_geometry_stage_on_the_shim(dr):
    """the rendering half of train_mlp.mlp_forward (:165-185) - per-view camera-space normals (get_normals
    on a batch of DISTINCT meshes), rasterize, ONE interpolate of the 30-wide feature stack [1 | normals | albedo | uniform
    | vertex_feat.expand], mask = first channel - line for line on the CUDA shim and on the CPU oracle: same coverage,
    features <= 1e-5, and the gradient that trains the per-vertex features (summed over the expanded batch) <= 1e-4."""
    from fmhr_b200 import utils as futils
    from oracle import refmath
    wl = synth.WORKLOADS["tiny"]
    v, f = synth.hand_mesh(wl["subdiv"], 1, seed=0)
    w2c, proj = synth.make_cameras(wl["n"], wl["H"], wl["W"], v.mean(0).astype(np.float64),
                                   extent=float(v[:, 1].max() - v[:, 1].min()))
    B, H, W, V = wl["n"], wl["H"], wl["W"], v.shape[0]
    g = torch.Generator().manual_seed(5)
    vertices = torch.tensor(v)[None].expand(B, -1, -1).contiguous()
    faces = torch.tensor(f)
    albedo = torch.rand(B, V, 3, generator=g)
    uni_vertices = torch.rand(B, V, 3, generator=g)      # vertices.clone().uniform_(0, 1) in the reference
    vertex_feat0 = torch.randn(V, 20, generator=g)
    wts = torch.randn(B, H, W, 30, generator=g)
    # the two einsums run once on the CPU so that both sides rasterise bit-identical clip positions
    vertsw = torch.cat([vertices, torch.ones_like(vertices[:, :, 0:1])], 2)
    rot_verts = torch.einsum('ijk,ikl->ijl', vertsw, torch.tensor(w2c))
    proj_verts = torch.einsum('ijk,ikl->ijl', rot_verts, torch.tensor(proj)).contiguous()

    def stage(drmod, get_normals, dev):
        to = lambda t: t.to(dev)
        vertex_feat = to(vertex_feat0).clone().requires_grad_(True)
        normals = get_normals(to(rot_verts)[:, :, :3], to(faces).long())
        rast_out, _ = drmod.rasterize(drmod.RasterizeGLContext(), to(proj_verts), to(faces), resolution=(H, W))
        feat = torch.cat([torch.ones_like(to(vertsw)[:, :, :1]), normals, to(albedo), to(uni_vertices),
                          vertex_feat.unsqueeze(0).expand(B, -1, -1)], 2)
        feat, _ = drmod.interpolate(feat, rast_out, to(faces))
        masks = feat[:, :, :, :1].contiguous()
        (feat * to(wts)).sum().backward()
        return rast_out.detach().cpu(), feat.detach().cpu(), masks.detach().cpu(), vertex_feat.grad.cpu()

    r_ref, f_ref, m_ref, g_ref = stage(orc, refmath.get_normals, "cpu")
    r_gpu, f_gpu, m_gpu, g_gpu = stage(dr, futils.get_normals, "cuda")
    assert f_gpu.shape == (B, H, W, 30)
    assert torch.equal(r_gpu[..., 3], r_ref[..., 3]) and (r_ref[..., 3] > 0).any()
    assert torch.equal(m_gpu[..., 0] > 0, r_ref[..., 3] > 0), "the reference selects pixels with masks[..., 0] > 0"
    assert torch.allclose(f_gpu, f_ref, rtol=1e-5, atol=2e-6)
    assert _rel(g_gpu, g_ref) < 1e-4


def test_interhand_rendered_mask_step_on_the_shim(dr):
    """get_interhand_data's mask step (get_data.py:246-254) - the segmentation of an InterHand capture is
    the rendered coverage of the fitted MANO mesh: rasterize, interpolate a ones attribute (A = 1), squeeze - line for line
    on the CUDA shim and on the CPU oracle; the two einsums run once so that both sides rasterise the same positions."""
    wl = synth.WORKLOADS["small"]
    v, f = synth.hand_mesh(wl["subdiv"], 1, seed=0)
    num, res = wl["n"], (wl["W"], wl["H"])   # the loader's `res` is (w, h)
    w2c, proj = synth.make_cameras(num, wl["H"], wl["W"], v.mean(0).astype(np.float64),
                                   extent=float(v[:, 1].max() - v[:, 1].min()))
    vertices, faces = torch.tensor(v)[None], torch.tensor(f)
    w2cs, projs = torch.tensor(w2c), torch.tensor(proj)
    vertsw = torch.cat([vertices, torch.ones_like(vertices[:, :, 0:1])], axis=2).expand(num, -1, -1)
    rot_verts = torch.einsum('ijk,ikl->ijl', vertsw, w2cs)
    proj_verts = torch.einsum('ijk,ikl->ijl', rot_verts, projs)

    def masks_of(drmod, dev):
        glctx = drmod.RasterizeGLContext()
        rast_out, _ = drmod.rasterize(glctx, proj_verts.to(dev), faces.to(dev), resolution=(res[1], res[0]))
        feat = torch.ones_like(vertsw[:, :, :1]).to(dev)   # an expanded (stride-0) view, as in the loader
        feat, _ = drmod.interpolate(feat, rast_out, faces.to(dev))
        return feat[:, :, :, :1].contiguous().squeeze(-1).cpu()

    ref, ours = masks_of(orc, "cpu"), masks_of(dr, "cuda")
    assert ours.shape == (num, wl["H"], wl["W"]) and (ref > 0).float().mean() > 0.02
    assert torch.equal(ours > 0, ref > 0)
    assert float((ours - ref).abs().max()) <= 2e-6   # u + v + (1 - u - v) of a ones attribute

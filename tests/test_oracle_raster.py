"""CPU checks of the C++ oracle (rasterize / interpolate / antialias): known answers and hand-written backward
passes against torch autograd of an fp64 restatement of the forward formulas (discrete decisions frozen)."""
import numpy as np
import torch

from fmhr_b200 import synth
from oracle import raster as orc


def _scene(workload="tiny"):
    wl = synth.WORKLOADS[workload]
    v, f = synth.hand_mesh(wl["subdiv"], 1, seed=0)
    w2c, proj = synth.make_cameras(wl["n"], wl["H"], wl["W"], v.mean(0).astype(np.float64),
                                   extent=float(v[:, 1].max() - v[:, 1].min()))
    vt = torch.tensor(v)
    vh = torch.cat([vt, torch.ones_like(vt[:, :1])], 1)[None].expand(wl["n"], -1, -1)
    pos = torch.einsum('ijk,ikl->ijl', torch.einsum('ijk,ikl->ijl', vh, torch.tensor(w2c)), torch.tensor(proj)).contiguous()
    return pos, torch.tensor(f), wl["H"], wl["W"]


def test_known_coverage_and_depth():
    H = W = 8
    ndc = lambda x, y: [2.0 * x / W - 1.0, 2.0 * y / H - 1.0]
    quad = [ndc(2, 1), ndc(6, 1), ndc(6, 5), ndc(2, 5)]
    pos = torch.tensor([[q + [0.0, 1.0] for q in quad] + [q + [-0.5, 1.0] for q in quad]], dtype=torch.float32)
    tri = torch.tensor([[0, 1, 2], [0, 2, 3], [4, 5, 6], [0, 1, 2]], dtype=torch.int32)
    r, _, keys = orc.rasterize_fwd(pos, tri[:2], (H, W), want_keys=True)
    ids = r[0, :, :, 3]
    assert (ids[1:5, 2:6] > 0).all() and int((ids > 0).sum()) == 16
    assert int((keys != -1).sum()) == 16
    r1, _, _ = orc.rasterize_fwd(pos, tri[:1], (H, W))
    r2, _, _ = orc.rasterize_fwd(pos, tri[1:2], (H, W))
    assert int(((r1[..., 3] > 0) & (r2[..., 3] > 0)).sum()) == 0
    r3, _, _ = orc.rasterize_fwd(pos, tri, (H, W))
    cov0 = r1[0, :, :, 3] > 0
    assert (r3[0, :, :, 3][cov0] == 3).all()          # nearer triangle wins
    r4, _, _ = orc.rasterize_fwd(pos, tri[[0, 3]], (H, W))
    assert (r4[0, :, :, 3][cov0] == 1).all()          # equal depth -> lower index
    # barycentrics: weights sum to one and reproduce the pixel centre
    u, v = r1[0, ..., 0][cov0], r1[0, ..., 1][cov0]
    p = pos[0, :3, :2]
    rec = u[:, None] * p[0] + v[:, None] * p[1] + (1 - u - v)[:, None] * p[2]
    ys, xs = torch.where(cov0)
    cen = torch.stack([(2 * xs + 1) / W - 1, (2 * ys + 1) / H - 1], 1)
    assert torch.allclose(rec, cen, atol=1e-6)
    # a triangle with a vertex behind the near plane (w <= 0) or outside |z| <= w is rejected
    bad = pos.clone()
    bad[0, 0, 3] = -1.0
    rb, _, _ = orc.rasterize_fwd(bad, tri[:1], (H, W))
    assert float(rb.abs().sum()) == 0.0


def test_rasterize_bwd_vs_autograd():
    pos, tri, H, W = _scene()
    rast, db, _ = orc.rasterize_fwd(pos, tri, (H, W))
    g = torch.Generator().manual_seed(0)
    dy = torch.randn(rast.shape, generator=g)
    got = orc.rasterize_bwd(pos, tri, rast, dy)
    p = pos.double().requires_grad_(True)
    n_i, y_i, x_i = torch.where(rast[..., 3] > 0)
    t = rast[n_i, y_i, x_i, 3].long() - 1
    fx = (2.0 * x_i.double() + 1.0) / W - 1.0
    fy = (2.0 * y_i.double() + 1.0) / H - 1.0
    P = [p[n_i, tri[t, k].long()] for k in range(3)]
    q = [(P[k][:, 0] - fx * P[k][:, 3], P[k][:, 1] - fy * P[k][:, 3]) for k in range(3)]
    cr = lambda a, b: a[0] * b[1] - a[1] * b[0]
    a0, a1, a2 = cr(q[1], q[2]), cr(q[2], q[0]), cr(q[0], q[1])
    at = a0 + a1 + a2
    iw = 1.0 / (at + 1e-6 * torch.sign(at.detach()))
    u, v = a0 * iw, a1 * iw
    # forward consistency (away from the clamp) and rast_db against finite differences of the same formula
    inside = (u > 1e-3) & (v > 1e-3) & (u + v < 1 - 1e-3)
    (u * dy[n_i, y_i, x_i, 0].double() + v * dy[n_i, y_i, x_i, 1].double()).sum().backward()
    ref = p.grad.float()
    assert float((got - ref).abs().max() / ref.abs().max()) < 1e-3  # fp32 oracle vs fp64 autograd
    assert int(inside.sum()) > 100


def test_interpolate_bwd_vs_autograd():
    pos, tri, H, W = _scene()
    N, V, _ = pos.shape
    rast, _, _ = orc.rasterize_fwd(pos, tri, (H, W))
    g = torch.Generator().manual_seed(1)
    for NA in (1, N):
        attr = torch.randn(NA, V, 5, generator=g)
        dy = torch.randn(N, H, W, 5, generator=g)
        ga, gr = orc.interpolate_bwd(attr, rast, tri, dy)
        a = attr.double().requires_grad_(True)
        r = rast.double().requires_grad_(True)
        n_i, y_i, x_i = torch.where(rast[..., 3] > 0)
        t = rast[n_i, y_i, x_i, 3].long() - 1
        u, v = r[n_i, y_i, x_i, 0], r[n_i, y_i, x_i, 1]
        ab = a[0] if NA == 1 else None
        A = [(ab[tri[t, k].long()] if NA == 1 else a[n_i, tri[t, k].long()]) for k in range(3)]
        out = u[:, None] * A[0] + v[:, None] * A[1] + (1 - u - v)[:, None] * A[2]
        assert torch.allclose(out.float(), orc.interpolate_fwd(attr, rast, tri)[n_i, y_i, x_i], rtol=1e-5, atol=1e-6)
        (out * dy[n_i, y_i, x_i].double()).sum().backward()
        assert torch.allclose(ga, a.grad.float(), rtol=1e-4, atol=1e-5)
        assert torch.allclose(gr, r.grad.float(), rtol=1e-4, atol=1e-5)


def test_antialias_semantics_and_bwd_vs_autograd():
    pos, tri, H, W = _scene("coarse")
    N = pos.shape[0]
    rast, _, _ = orc.rasterize_fwd(pos, tri, (H, W))
    g = torch.Generator().manual_seed(2)
    color = torch.rand(N, H, W, 3, generator=g) * (rast[..., 3:] > 0)
    out, items = orc.antialias_fwd(color, rast, pos, tri, want_items=True)
    assert items.shape[0] > 100
    pix0, pix1, t, di, d, from1 = [items[:, k].long() for k in range(6)]
    alpha = items[:, 6].contiguous().view(torch.float32)
    clamped = items[:, 7].bool()
    assert float(alpha.abs().max()) <= 0.5
    # every blended pair straddles a silhouette: in this closed-ish mesh most pairs have one empty pixel
    flat_id = rast[..., 3].reshape(-1)
    assert (flat_id[pix0] != flat_id[pix1]).all()
    # the interior of the coverage mask is untouched, the blend only moves values between the two pixels of a pair
    touched = torch.zeros(N * H * W, dtype=torch.bool)
    touched[torch.where(alpha > 0, pix0, pix1)] = True
    assert torch.equal(out.reshape(-1, 3)[~touched], color.reshape(-1, 3)[~touched])
    # coverage antialias: values stay in [0,1]
    cov = (rast[..., 3:] > 0).float()
    aa_cov = orc.antialias_fwd(cov, rast, pos, tri)
    assert float(aa_cov.min()) >= 0.0 and float(aa_cov.max()) <= 1.0 and (aa_cov != cov).any()

    # backward vs autograd of alpha(pos) with the discrete choices frozen
    dy = torch.randn(N, H, W, 3, generator=g)
    gc, gp = orc.antialias_bwd(color, rast, pos, tri, dy)
    p = pos.double().requires_grad_(True)
    c = color.double().requires_grad_(True)
    n_i = pix0 // (H * W)
    rem = pix0 % (H * W)
    py, px = rem // W, rem % W
    qx = (px + from1 * (1 - d)).double()
    qy = (py + from1 * d).double()
    i1 = tri[t, (di + 1) % 3].long()
    i2 = tri[t, (di + 2) % 3].long()
    xh, yh = 0.5 * W, 0.5 * H
    fx, fy = qx + 0.5 - xh, qy + 0.5 - yh
    P1, P2 = p[n_i, i1], p[n_i, i2]
    x1, y1 = P1[:, 0] / P1[:, 3] * xh - fx, P1[:, 1] / P1[:, 3] * yh - fy
    x2, y2 = P2[:, 0] / P2[:, 3] * xh - fx, P2[:, 1] / P2[:, 3] * yh - fy
    sw = d.bool()
    X1, Y1 = torch.where(sw, y1, x1), torch.where(sw, x1, y1)
    X2, Y2 = torch.where(sw, y2, x2), torch.where(sw, x2, y2)
    ds = 1.0 - 2.0 * from1.double()
    dc = ds * (X1 * (Y2 - Y1) - Y1 * (X2 - X1)) / (Y2 - Y1)
    a_t = ds * (0.5 - dc)
    a_t = torch.where(clamped, alpha.double(), a_t)  # clamped crossings carry no gradient
    assert torch.allclose(a_t.float(), alpha, atol=2e-4)
    cf = c.reshape(-1, 3)
    recv = torch.where(alpha > 0, pix0, pix1)
    outf = cf.clone().index_add(0, recv, a_t[:, None] * (cf[pix1] - cf[pix0]))
    assert torch.allclose(outf.float().reshape(out.shape), out, atol=1e-5)
    (outf * dy.double().reshape(-1, 3)).sum().backward()
    assert torch.allclose(gc, c.grad.float(), rtol=1e-4, atol=1e-5)
    ref = p.grad.float()
    assert float((gp - ref).abs().max() / ref.abs().max()) < 2e-3  # fp32 oracle vs fp64 autograd


def test_topology():
    tri = torch.tensor([[0, 1, 2], [2, 1, 3], [2, 3, 4]], dtype=torch.int32)
    opp = orc.antialias_topology(tri)
    assert opp.tolist() == [[3, -1, -1], [-1, 4, 0], [-1, -1, 1]]

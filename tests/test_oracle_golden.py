"""The oracle's restatement of models/utils.py + models/ncc_utils.py against fixtures produced by the
reference's own code (oracle/gen_golden.py).  CPU only."""
import numpy as np
import torch

from oracle import refmath as rm

T = lambda a, **k: torch.tensor(np.asarray(a), **k)


def test_tetrahedron_known_answers(golden):
    v, f = T(golden["tet_v"]), T(golden["tet_f"])
    L = rm.compute_laplacian(v, f).to_dense()
    assert torch.equal(L, T(golden["tet_L"]))
    assert torch.allclose(torch.diag(L), -torch.ones(4))
    off = L - torch.diag(torch.diag(L))
    assert torch.allclose(off[off != 0], torch.full((12,), 1.0 / 3.0))
    assert torch.allclose(rm.get_normals(v[None], f)[0], T(golden["tet_normals"]), atol=1e-7)
    assert torch.allclose(rm.laplacian_smoothing(v, f), T(golden["tet_lap"]), atol=1e-7)


def test_get_normals_fwd_bwd(golden):
    v = T(golden["verts"]).requires_grad_(True)
    f = T(golden["faces"]).long()
    wn = T(golden["wn"])
    n = rm.get_normals(v[None].expand(wn.shape[0], -1, -1), f)
    assert torch.equal(n.detach(), T(golden["normals"]))  # same op order -> bit-exact on CPU
    (n * wn).sum().backward()
    assert torch.allclose(v.grad, T(golden["g_normals"]), rtol=1e-5, atol=1e-5)


def test_laplacian_fwd_bwd(golden):
    f = T(golden["faces"]).long()
    for key, lk, gk in (("verts", "lap", "g_lap"), ("alb", "lap_alb", "g_lap_alb")):
        x = T(golden[key]).requires_grad_(True)
        l = rm.laplacian_smoothing(x, f)
        assert torch.allclose(l.detach(), T(golden[lk]), rtol=1e-6)
        l.backward()
        assert torch.allclose(x.grad, T(golden[gk]), rtol=1e-5, atol=1e-9)
    Lx = rm.compute_laplacian(T(golden["verts"]), f).mm(T(golden["verts"]))
    assert torch.allclose(Lx, T(golden["Lx"]), rtol=1e-5, atol=1e-8)


def test_sh_radiance_and_matrix(golden):
    n = T(golden["sh_normals"]).requires_grad_(True)
    c = T(golden["sh_coeff"]).requires_grad_(True)
    r = rm.get_radiance(c, n, 3)
    assert torch.equal(r.detach(), T(golden["radiance"]))
    (r * T(golden["sh_w"])).sum().backward()
    assert torch.allclose(n.grad, T(golden["g_sh_normals"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(c.grad, T(golden["g_sh_coeff"]), rtol=1e-5, atol=1e-6)
    m = rm.get_matrix(n.detach(), 3)
    assert torch.equal(m, T(golden["matrix_t"]))
    assert np.allclose(m.numpy(), golden["matrix_np"], rtol=1e-6, atol=1e-7)
    r1 = rm.get_radiance(T(golden["sh_coeff1"]), n.detach(), 3)
    assert torch.equal(r1, T(golden["radiance1"]))
    # basis at axis-aligned normals
    ax = torch.eye(3)
    mm = rm.get_matrix(ax, 3)
    assert torch.equal(mm[0], torch.tensor([1., 0, 0, 1, 0, 0, -1, 0, 1]))   # +x
    assert torch.equal(mm[1], torch.tensor([1., 1, 0, 0, 0, 0, -1, 0, -1]))  # +y
    assert torch.equal(mm[2], torch.tensor([1., 0, 1, 0, 0, 0, 2, 0, 0]))    # +z


def test_ncc(golden):
    src = T(golden["ncc_src"]).requires_grad_(True)
    ref, msk = T(golden["ncc_ref"]), T(golden["ncc_mask"])
    ncc = rm.NCC(ref, src, torch.ones_like(ref), msk)
    assert torch.allclose(ncc.detach(), T(golden["ncc"]), rtol=1e-5, atol=1e-6)
    (ncc * T(golden["ncc_w"])).sum().backward()
    assert torch.allclose(src.grad, T(golden["g_ncc_src"]), rtol=1e-4, atol=1e-6)
    # known answers: identical patches -> 1 ; constant patches -> 0
    p = torch.rand(1, 5, 49)
    one = rm.NCC(p, p.clone(), torch.ones_like(p), torch.ones_like(p))
    assert torch.allclose(one, torch.ones(5), atol=1e-5)
    zero = rm.NCC(p, torch.full_like(p, 0.3), torch.ones_like(p), torch.ones_like(p))
    assert torch.allclose(zero, torch.zeros(5), atol=1e-6)

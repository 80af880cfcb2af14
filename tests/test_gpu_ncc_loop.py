"""NCC photo-consistency term inside the phase-B iteration (BASELINE.json configs[2]) against its executable
specification oracle.ncc_term (plain PyTorch; the NCC arithmetic itself = models/ncc_utils.py:4-35, pinned by the golden
fixtures of test_gpu_ham.py::test_radiance_matrix_ncc_golden)."""
import numpy as np
import pytest
import torch

from fmhr_b200 import synth
from oracle import compare
from oracle import ham as oham
from oracle import ncc_term as oncc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene():
    return synth.build_scene("small", oham.render_views)


def _setup(scene, n_points, half, weight=10.0, fused=True, **kw):
    from fmhr_b200.ncc_term import NccTerm
    opt = compare.make_optimizer(scene, torch.device("cuda"), **kw)
    gray = torch.tensor(np.asarray(scene["imgs"]).mean(-1).astype(np.float32)).cuda()
    n = scene["imgs"].shape[0]
    term = NccTerm(opt, gray, ref_view=0, src_views=list(range(1, n)), weight=weight, n_points=n_points, half=half, seed=3, fused=fused)
    return opt, term, gray


@pytest.mark.parametrize("half,fused", [(2, False), (5, False), (2, True), (5, True), (6, True)])
def test_ncc_term_matches_the_specification(scene, half, fused):
    """fused: the one-kernel form (fmhr_ncc_term_fused, patches of <= 128 samples; half = 6 -> 169 samples falls back to the
    chain); unfused: the four-call chain, whose materialised patches / masks are compared as well."""
    opt, term, gray = _setup(scene, 3000, half, debug=True, fused=fused)
    n = scene["imgs"].shape[0]
    views = list(range(n))
    opt.step_phase_b(views)       # one iteration with the term enqueued between render and update
    st = oham.HamState(scene)
    verts = (st.vertices_tmp + st.delta).detach().requires_grad_(True)
    loss, ncc, patches, pmask = oncc.ncc_term(verts, st.faces, term.pt_face.cpu(), term.pt_bary.cpu(), st.w2cs, st.projs,
                                              term.view_idx.cpu(), gray.cpu(), st.masks, term.weight, half)
    loss.backward()
    if fused and half <= 5:
        assert term.patches is None, "the fused form must not materialise [Nv,Np,Npx] arrays"
    else:
        assert float((term.patches.cpu() - patches).abs().max()) < 2e-5
        assert float((term.patch_mask.cpu() != pmask).float().mean()) < 1e-4    # (a sample exactly on a pixel border may round apart)
    # NCC divides by sqrt(var_ref * var_src): patches of nearly constant colour amplify rounding, hence the quantile bar
    dn = (term.ncc.cpu() - ncc.detach()).abs()
    assert float(torch.quantile(dn.flatten(), 0.999)) < 1e-3 and abs(float(term.loss) - float(loss)) < 1e-4 * abs(float(loss))
    g, gref = term.grad_delta.cpu(), verts.grad
    err = (g - gref).abs() / gref.abs().max()
    assert float(torch.quantile(err.flatten(), 0.999)) < 1e-3 and float((g - gref).norm() / gref.norm()) < 1e-3, \
        (float(err.max()), float((g - gref).norm() / gref.norm()))


def test_fused_term_equals_the_chain(scene):
    """Same arithmetic in the same order: NCC values and the vertex gradient of the fused kernel against the four-call chain."""
    a, ta, _ = _setup(scene, 4000, 5, fused=True)
    b, tb, _ = _setup(scene, 4000, 5, fused=False)
    views = list(range(scene["imgs"].shape[0]))
    a.step_phase_b(views)
    b.step_phase_b(views)
    # (the fused kernel evaluates the shared fractional sample position once per (point, view), the chain per sample: the
    #  patches agree to ~1e-6, and NCC amplifies that where a patch is nearly constant - hence the quantile)
    dn = (ta.ncc - tb.ncc).abs().flatten()
    assert float(torch.quantile(dn, 0.999)) < 1e-4 and float(dn.max()) < 5e-2, (float(torch.quantile(dn, 0.999)), float(dn.max()))
    assert compare.rel_l2(ta.grad_delta, tb.grad_delta) < 1e-4
    assert abs(float(ta.loss) - float(tb.loss)) < 1e-5 * abs(float(tb.loss))


def test_iteration_with_ncc_term_matches_the_oracle(scene):
    """The whole phase-B gradient w.r.t. delta with the extra term: fused passes + NCC hook vs oracle loss + NCC loss."""
    opt, term, gray = _setup(scene, 3000, 5, debug=True)
    plain = compare.make_optimizer(scene, torch.device("cuda"), debug=True)
    n = scene["imgs"].shape[0]
    views = list(range(n))
    opt.step_phase_b(views)
    plain.step_phase_b(views)
    st = oham.HamState(scene)
    loss, terms = oham.phase_b_forward(st, views)
    l_ncc, _, _, _ = oncc.ncc_term(st.vertices_tmp + st.delta, st.faces, term.pt_face.cpu(), term.pt_bary.cpu(), st.w2cs,
                                   st.projs, term.view_idx.cpu(), gray.cpu(), st.masks, term.weight, 5)
    (loss + l_ncc).backward()
    g, gref = opt.dbg_grad.cpu()[:, :3], st.delta.grad
    assert compare.rel_l2(g, gref) < 2e-4, compare.rel_l2(g, gref)
    # the term did change the gradient, and only the delta gradient (albedo is untouched)
    assert compare.rel_l2(plain.dbg_grad.cpu()[:, :3], gref) > 1e-2
    assert compare.rel_to_max(opt.dbg_grad[:, 3:], plain.dbg_grad[:, 3:]) < 1e-5   # (only the order of the atomics differs)


def test_graph_replay_with_ncc_term_matches_eager(scene):
    a, _, _ = _setup(scene, 2000, 5)
    b, tb, _ = _setup(scene, 2000, 5, use_graphs=True)
    n = scene["imgs"].shape[0]
    for views in (list(range(n)), list(range(n)), [0, 1, 2]):
        la, lb = a.step_phase_b(views).cpu(), b.step_phase_b(views).cpu()
        assert torch.allclose(la, lb, rtol=1e-4, atol=1e-6), (la, lb)
    assert float((a.delta - b.delta).abs().gt(2e-2 * scene["conf"]["lr"]).float().mean()) < 2e-3
    assert np.isfinite(float(tb.loss)) and float(tb.loss) > 0

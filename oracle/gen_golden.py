"""ORACLE (test infrastructure) — generate tests/golden/*.npz from the REFERENCE'S OWN code.

Runs only in the build container (needs /root/reference).  Imports /root/reference/models/utils.py and
models/ncc_utils.py VERBATIM (unrelated missing imports skimage / plyfile / trimesh are stubbed in
sys.modules), evaluates get_normals, get_matrix, get_radiance, laplacian_smoothing,
compute_laplacian and NCC (forward and autograd gradients) on seeded inputs and stores inputs + outputs.

    python -m oracle.gen_golden            # rewrites tests/golden/refmath_*.npz
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference():
    for name in ("skimage", "skimage.measure", "plyfile", "trimesh"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            sys.modules[name] = m
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    sys.modules["plyfile"].PlyData = object
    sys.modules["plyfile"].PlyElement = object
    sys.path.insert(0, REF)
    import models.utils as ru
    import models.ncc_utils as rn
    return ru, rn


def main():
    from fmhr_b200 import synth
    ru, rn = import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    rng = np.random.default_rng(0)

    # known-answer: tetrahedron
    tv = torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=torch.float32)
    tf = torch.tensor([[0, 2, 1], [0, 1, 3], [0, 3, 2], [1, 2, 3]], dtype=torch.int64)
    tet_n = ru.get_normals(tv[None], tf)[0]
    tet_L = ru.compute_laplacian(tv, tf).to_dense()
    tet_lap = ru.laplacian_smoothing(tv, tf, "uniform")

    # mesh: subdiv-1 synthetic hand (3,098 verts) with noise
    v, f = synth.hand_mesh(1, 1, seed=5, noise=2e-3)
    vt = torch.tensor(v, requires_grad=True)
    ft = torch.tensor(f, dtype=torch.int64)
    B = 2
    wn = torch.tensor(rng.normal(size=(B, v.shape[0], 3)).astype(np.float32))
    nrm = ru.get_normals(vt[None].expand(B, -1, -1), ft)
    (nrm * wn).sum().backward()
    g_normals = vt.grad.clone()
    vt.grad = None

    lap = ru.laplacian_smoothing(vt, ft, "uniform")
    lap.backward()
    g_lap = vt.grad.clone()
    Lx = ru.compute_laplacian(vt.detach(), ft).mm(vt.detach())

    alb = torch.tensor(rng.uniform(0.2, 0.9, size=(v.shape[0], 3)).astype(np.float32), requires_grad=True)
    lap_alb = ru.laplacian_smoothing(alb, ft, "uniform")
    lap_alb.backward()

    # SH
    Np = 4096
    nn_ = rng.normal(size=(Np, 3)).astype(np.float32)
    nn_ /= np.linalg.norm(nn_, axis=1, keepdims=True)
    nt = torch.tensor(nn_, requires_grad=True)
    coeff = torch.tensor(rng.normal(size=(Np, 9)).astype(np.float32) * 0.3, requires_grad=True)
    wr = torch.tensor(rng.normal(size=(Np,)).astype(np.float32))
    rad = ru.get_radiance(coeff, nt, 3)
    (rad * wr).sum().backward()
    mat_t = ru.get_matrix(nt.detach(), 3)
    mat_np = ru.get_matrix(nn_, 3)
    coeff1 = torch.tensor(rng.normal(size=(9,)).astype(np.float32))
    rad1 = ru.get_radiance(coeff1, nt.detach(), 3)

    # NCC
    Nv, Npt, Npx = 3, 64, 121
    ref = torch.tensor(rng.uniform(size=(1, Npt, Npx)).astype(np.float32))
    src = torch.tensor(rng.uniform(size=(Nv, Npt, Npx)).astype(np.float32), requires_grad=True)
    msk = torch.tensor((rng.uniform(size=(Nv, Npt, Npx)) > 0.3).astype(np.float32))
    msk[0, 0] = 0  # empty patch -> count forced to 1
    src.data[1, 1] = 0.5  # constant patch -> zero variance forced to 1
    ncc = rn.NCC(ref, src, torch.ones_like(ref), msk)
    wncc = torch.tensor(rng.normal(size=(Nv, Npt)).astype(np.float32))
    (ncc * wncc).sum().backward()

    np.savez_compressed(
        os.path.join(OUT, "refmath_v1.npz"),
        tet_v=tv.numpy(), tet_f=tf.numpy(), tet_normals=tet_n.numpy(), tet_L=tet_L.numpy(), tet_lap=tet_lap.numpy(),
        verts=v, faces=f, wn=wn.numpy(), normals=nrm.detach().numpy(), g_normals=g_normals.numpy(),
        lap=lap.detach().numpy(), g_lap=g_lap.numpy(), Lx=Lx.numpy(),
        alb=alb.detach().numpy(), lap_alb=lap_alb.detach().numpy(), g_lap_alb=alb.grad.numpy(),
        sh_normals=nn_, sh_coeff=coeff.detach().numpy(), sh_w=wr.numpy(), radiance=rad.detach().numpy(),
        g_sh_normals=nt.grad.numpy(), g_sh_coeff=coeff.grad.numpy(), matrix_t=mat_t.numpy(), matrix_np=mat_np,
        sh_coeff1=coeff1.numpy(), radiance1=rad1.numpy(),
        ncc_ref=ref.numpy(), ncc_src=src.detach().numpy(), ncc_mask=msk.numpy(), ncc=ncc.detach().numpy(),
        ncc_w=wncc.numpy(), g_ncc_src=src.grad.numpy(),
    )
    print("wrote", os.path.join(OUT, "refmath_v1.npz"))


def import_reference_repose():
    """/root/reference/repose.py imported verbatim (its unrelated imports - smplx, pyhocon, segment_anything, trimesh -
    are stubbed; `smplx.create` returns a dummy so that get_data's module-level MANO layers construct)."""
    import_reference()
    for name in ("smplx", "pyhocon", "segment_anything", "trimesh.remesh"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pyhocon"].ConfigFactory = object
    sys.modules["segment_anything"].sam_model_registry = {}
    sys.modules["segment_anything"].SamPredictor = object
    sys.modules["smplx"].create = lambda *a, **k: types.SimpleNamespace()
    import repose
    return repose


def gen_repose():
    """tests/golden/repose_v1.npz: the reference's subdivide_weight (repose.py:14-24) on one round of subdivision of the
    synthetic 778-vertex hand with random 16-joint skinning weights."""
    from fmhr_b200 import synth
    rp = import_reference_repose()
    rng = np.random.default_rng(7)
    v0, f0 = synth.base_hand_mesh()
    w0 = rng.dirichlet(np.ones(16) * 0.3, size=v0.shape[0])
    v1, f1 = synth.subdivide_loop(v0, f0, 1)
    w1 = rp.subdivide_weight(w0, f1)
    np.savez_compressed(os.path.join(OUT, "repose_v1.npz"), w0=w0.astype(np.float32), f0=f0.astype(np.int32),
                        f1=f1.astype(np.int32), w1=w1.astype(np.float32))
    print("repose golden:", w1.shape, "row sums", float(w1.sum(1).min()), float(w1.sum(1).max()))


if __name__ == "__main__":
    main()
    gen_repose()

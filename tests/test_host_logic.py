"""Host-side logic that needs no GPU: synthetic inputs, view sharding, and the multi-rank reduction rule
(2 gloo ranks, each running the ORACLE on its view shard, reproduce the single-process reference gradients)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fmhr_b200 import synth
from fmhr_b200.dist import normalisation_scales, shard_views

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mesh_counts_match_the_reference_shapes():
    v, f = synth.base_hand_mesh()
    assert v.shape == (778, 3) and f.shape == (1538, 3)            # MANO template counts
    e, _ = synth.unique_edges(f, 778)
    assert e.shape[0] == 2315                                      # disk topology with a 16-edge boundary
    v3, f3 = synth.hand_mesh(3)
    assert v3.shape == (49281, 3) and f3.shape == (98432, 3)       # mesh_sfs_optim.py:82 iterations=3
    v2, f2 = synth.hand_mesh(1, hands=2)
    assert f2.max() == v2.shape[0] - 1 and f2[f2.shape[0] // 2:].min() == v2.shape[0] // 2  # offset faces, :75-88
    assert v3.dtype == np.float32 and f3.dtype == np.int32


def test_camera_convention():
    w2c, proj = synth.make_cameras(5, 64, 48, np.zeros(3))
    assert w2c.flags["C_CONTIGUOUS"] and proj.flags["C_CONTIGUOUS"]
    P = proj[0].T                                                  # stored transposed (get_data.py:96-97)
    assert P[2, 3] == np.float32(-0.1) and P[3, 2] == 1.0 and P[2, 2] == 0.0 and P[3, 3] == 0.0  # get_data.py:70-73
    R = w2c[0].T[:3, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-5)
    clip = np.array([0.0, 0.0, 0.0, 1.0]) @ w2c[0] @ proj[0]       # the look-at point projects near the image centre
    assert clip[3] > 1.5 and abs(clip[0] / clip[3]) < 0.2 and abs(clip[1] / clip[3]) < 0.2 and clip[2] == np.float32(-0.1)


def test_view_sharding_partitions_the_views():
    for num, world in ((48, 8), (16, 3), (5, 8)):
        shards = [shard_views(num, r, world) for r in range(world)]
        assert sorted(sum(shards, [])) == list(range(num))
    s_photo, s_mask = normalisation_scales(synth.CONF["ih_sfs"], 1000, 48, 512, 334)
    assert s_photo == pytest.approx(30.0 / 3000.0) and s_mask == pytest.approx(400.0 / (48 * 512 * 334))


def _rank_main(rank, world, port, scene, out):
    """Each rank: un-normalised photometric / mask sums over ITS views (oracle math), one gloo all-reduce, then the
    global normalisation - exactly what fmhr_ham_step_render / allreduce_packed / fmhr_ham_step_update do on GPUs."""
    sys.path.insert(0, ROOT)
    import warnings
    warnings.filterwarnings("ignore")
    torch.set_num_threads(2)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fmhr_b200.dist import allreduce_packed
    from oracle import ham as oham
    st = oham.HamState(scene)
    n_all = scene["imgs"].shape[0]
    views = shard_views(n_all, rank, world)
    c = st.conf
    # local un-normalised sums: undo the reference's means
    zero = dict(c, lap_weight=0.0, albedo_weight=0.0, edge_weight=0.0, delta_weight=0.0)
    st.conf = zero
    loss, terms = oham.phase_b_forward(st, views)
    n_valid = terms["n_valid"]
    P_local = len(views) * st.H * st.W
    abs_sum = terms["sfs"] / c["sfs_weight"] * (3.0 * n_valid)
    msk_sum = terms["mask"] / c["mask_weight"] * P_local
    g_abs = torch.autograd.grad(abs_sum, [st.delta, st.albedo], retain_graph=True, allow_unused=True)
    g_msk = torch.autograd.grad(msk_sum, [st.delta], allow_unused=True)
    V = st.delta.shape[0]
    packed = torch.cat([g_abs[0].reshape(-1), g_msk[0].reshape(-1), g_abs[1].reshape(-1),
                        torch.tensor([float(n_valid), float(abs_sum), float(msk_sum), 0.0])]).detach()
    allreduce_packed(packed)                                       # the single collective of the iteration
    s_photo, s_mask = normalisation_scales(c, float(packed[-4]), n_all, st.H, st.W)
    g_delta = s_photo * packed[:3 * V] + 0.5 * s_mask * packed[3 * V:6 * V]
    g_albedo = s_photo * packed[6 * V:9 * V]
    if rank == 0:
        torch.save(dict(g_delta=g_delta.view(V, 3), g_albedo=g_albedo.view(V, 3), n_valid=float(packed[-4]),
                        sfs=c["sfs_weight"] * float(packed[-3]) / (3.0 * float(packed[-4])),
                        mask=c["mask_weight"] * float(packed[-2]) / (n_all * st.H * st.W)), out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_reduction_matches_single_process(tmp_path):
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import ham as oham
    scene = synth.build_scene("tiny", oham.render_views)
    out = str(tmp_path / "rank0.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_rank_main, args=(2, port, scene, out), nprocs=2, join=True)
    got = torch.load(out)
    # single-process reference over ALL views, regularisers switched off the same way
    st = oham.HamState(scene)
    st.conf = dict(st.conf, lap_weight=0.0, albedo_weight=0.0, edge_weight=0.0, delta_weight=0.0)
    loss, terms = oham.phase_b_forward(st, list(range(scene["imgs"].shape[0])))
    loss.backward()
    assert got["n_valid"] == terms["n_valid"]
    assert got["sfs"] == pytest.approx(float(terms["sfs"]), rel=1e-5)
    assert got["mask"] == pytest.approx(float(terms["mask"]), rel=1e-5, abs=1e-9)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    assert rel(got["g_delta"], st.delta.grad) < 1e-4
    assert rel(got["g_albedo"], st.albedo.grad[0]) < 1e-4


def _build_meshlets(tri, verts, tpm, V):
    import ctypes
    import numpy as np
    from fmhr_b200 import _build, _lib
    _build.build()
    lib = _lib.load()
    hp = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)
    T = tri.shape[0]
    nm, nr, mv = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    call = lambda a, b, c: lib.fmhr_meshlets_build_host(hp(tri), hp(verts), V, T, tpm, ctypes.byref(nm), ctypes.byref(nr),
                                                         ctypes.byref(mv), hp(a), hp(b), hp(c))
    assert call(None, None, None) == 0
    vptr = np.zeros(nm.value + 1, dtype=np.int32)
    vrefs = np.zeros(nr.value, dtype=np.int32)
    tri2 = np.zeros((nm.value * tpm, 2), dtype=np.uint32)
    assert call(vptr, vrefs, tri2) == 0
    return nm.value, mv.value, vptr, vrefs, tri2


@pytest.mark.parametrize("tpm", [256, 1024])
def test_meshlets_partition_the_mesh(tpm):
    """fmhr_meshlets_build_host (host-side setup of the coverage kernel): every triangle appears in exactly one meshlet,
    its 10-bit local indices decode to its own vertices in the original corner order, no meshlet exceeds 1024 vertices,
    and the result does not depend on whether vertex positions (Morton order) are supplied."""
    import numpy as np
    from fmhr_b200 import synth
    v, f = synth.base_hand_mesh()
    v, f = synth.subdivide_loop(v, f, 2)
    tri = np.ascontiguousarray(f, dtype=np.int32)
    verts = np.ascontiguousarray(v, dtype=np.float32)
    for use_pos in (True, False):
        M, mx, vptr, vrefs, tri2 = _build_meshlets(tri, verts if use_pos else None, tpm, verts.shape[0])
        assert 0 < mx <= 1024 and vptr[0] == 0 and vptr[-1] == vrefs.shape[0]
        assert M >= (tri.shape[0] + tpm - 1) // tpm
        seen = np.zeros(tri.shape[0], dtype=np.int64)
        for m in range(M):
            loc = vrefs[vptr[m]:vptr[m + 1]]
            assert loc.shape[0] <= 1024 and np.unique(loc).shape[0] == loc.shape[0]
            rec = tri2[m * tpm:(m + 1) * tpm]
            real = rec[:, 0] != 0xFFFFFFFF
            assert np.all(rec[~real, 1] == 0xFFFFFFFF)      # padding is all ones
            tid = rec[real, 1].astype(np.int64)
            seen[tid] += 1
            packed = rec[real, 0]
            corners = np.stack([packed & 1023, (packed >> 10) & 1023, (packed >> 20) & 1023], axis=1).astype(np.int64)
            assert corners.max() < loc.shape[0]
            assert np.array_equal(loc[corners], tri[tid])
        assert np.all(seen == 1)


def test_meshlets_reject_bad_input():
    import ctypes
    import numpy as np
    from fmhr_b200 import _build, _lib
    _build.build()
    lib = _lib.load()
    tri = np.array([[0, 1, 5]], dtype=np.int32)   # vertex 5 out of range for V = 3
    nm, nr, mv = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib.fmhr_meshlets_build_host(ctypes.c_void_p(tri.ctypes.data), None, 3, 1, 1024, ctypes.byref(nm), ctypes.byref(nr),
                                      ctypes.byref(mv), None, None, None)
    assert rc == -1 and b"invalid argument" in lib.fmhr_last_error_string()


def test_result_files_follow_the_reference_formats(tmp_path):
    """fmhr_b200.export: <scan>.pt / <scan>.obj / <scan>_c.obj as mesh_sfs_optim.py:19-28,321-337 writes them (the files
    neural_render.py:90-96 loads): tensor shapes, vertex / face order, BGR->RGB colour clamp(0.5*albedo), flipped winding."""
    from fmhr_b200 import export
    rng = np.random.default_rng(0)
    V, F, num = 7, 5, 3
    verts = rng.normal(size=(V, 3)).astype(np.float32)
    faces = rng.integers(0, V, size=(F, 3)).astype(np.int32)
    albedo = rng.uniform(0, 3, size=(V, 3)).astype(np.float32)   # some entries clamp at 1
    sh = rng.normal(size=(num, 9)).astype(np.float32)
    export.save_ham_results(str(tmp_path), 4, verts, faces, albedo, sh)
    pt = torch.load(str(tmp_path / "4.pt"))
    assert set(pt) == {"sh_coeff", "albedo"} and tuple(pt["albedo"].shape) == (1, V, 3) and tuple(pt["sh_coeff"].shape) == (num, 9)
    assert torch.equal(pt["albedo"][0], torch.tensor(albedo)) and torch.equal(pt["sh_coeff"], torch.tensor(sh))
    v, c, f = export.load_obj(str(tmp_path / "4.obj"))
    assert c is None and np.allclose(v, verts, atol=1e-6) and np.array_equal(f, faces)
    v, c, f = export.load_obj(str(tmp_path / "4_c.obj"))
    assert np.allclose(v, verts, atol=1e-4)
    assert np.array_equal(f, faces[:, [0, 2, 1]])                                   # flipped winding
    assert np.allclose(c, np.clip(0.5 * albedo, 0, 1)[:, ::-1], atol=1e-4)          # BGR -> RGB


def test_mask_boxes_contain_every_set_pixel():
    """Loader metadata of the host-batch path (fmhr_ham_host_u8_submit_boxes): the half-open rectangle of every view must
    contain every pixel with mask byte > 127 - outside it nothing is uploaded - and be tight; an empty mask gives (0,0,0,0)."""
    import numpy as np
    from fmhr_b200.ham import HostStreamingStepper
    rng = np.random.default_rng(0)
    m = np.zeros((5, 40, 56), dtype=np.uint8)
    m[0, 3:17, 10:31] = 255
    m[1, 0, 0] = 200            # a single pixel in the corner
    m[2, 39, 55] = 128          # ... and in the opposite one
    m[3] = (rng.random((40, 56)) > 0.97) * 255
    m[3, 5, 7] = 127            # not set: the loader's rule is > 127
    boxes = HostStreamingStepper.mask_boxes(m)
    assert boxes.dtype == torch.int32 and tuple(boxes.shape) == (5, 4)
    assert boxes[0].tolist() == [3, 17, 10, 31] and boxes[1].tolist() == [0, 1, 0, 1] and boxes[2].tolist() == [39, 40, 55, 56]
    assert boxes[4].tolist() == [0, 0, 0, 0]
    for v in range(5):
        y0, y1, x0, x1 = boxes[v].tolist()
        inside = np.zeros((40, 56), dtype=bool)
        inside[y0:y1, x0:x1] = True
        assert not ((m[v] > 127) & ~inside).any()
        if y1 > y0:
            s = m[v] > 127
            assert s[y0].any() and s[y1 - 1].any() and s[:, x0].any() and s[:, x1 - 1].any()
    assert torch.equal(HostStreamingStepper.mask_boxes(torch.from_numpy(m)), boxes)


def test_demo_camera_convention_matches_the_reference_loader():
    """synth.demo_cameras (RQ decomposition + clip fix-up, get_data.py:62-76,96-97) against the matrices the reference's
    own load_K_Rt_from_P produced for the 16 demo cameras (fixture made by oracle/gen_demo_fixture.py)."""
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "demo1_320x256.npz"))
    w2cs, projs = synth.demo_cameras(fx["world_mats"], fx["scale_mats"], tuple(int(x) for x in fx["cap_res"]))
    assert w2cs.shape == (16, 4, 4) and w2cs.dtype == np.float32
    assert np.abs(w2cs - fx["w2cs"]).max() < 1e-5 and np.abs(projs - fx["projs"]).max() < 1e-5
    assert np.all(projs[:, 3, 2] == -0.1) and np.all(projs[:, 2, 3] == 1.0)  # z_clip = -0.1, w_clip = z_cam
    scene = synth.demo_scene(fx)
    assert scene["vertices"].shape == (49281, 3) and scene["faces"].shape == (98432, 3)
    # the posed hand is in front of every camera and lands on the segmentation
    vh = np.concatenate([scene["vertices"], np.ones((49281, 1), np.float32)], 1)
    inside = []
    for i in range(16):
        clip = vh @ scene["w2cs"][i] @ scene["projs"][i]
        assert clip[:, 3].min() > 0.5
        px = ((clip[:, 0] / clip[:, 3] + 1) * 0.5 * scene["W"]).astype(int).clip(0, scene["W"] - 1)
        py = ((clip[:, 1] / clip[:, 3] + 1) * 0.5 * scene["H"]).astype(int).clip(0, scene["H"] - 1)
        inside.append(float(scene["masks"][i][py, px].mean()))
    assert np.mean(inside) > 0.2 and min(inside) > 0.0  # (some cameras see the hand mostly outside their frame)

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

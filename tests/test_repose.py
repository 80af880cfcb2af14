"""Skinning-weight subdivision (repose.py:14-41 of the reference) against the reference's own
subdivide_weight (golden fixture made by oracle/gen_golden.py:gen_repose from /root/reference/repose.py)."""
import os
import pickle

import numpy as np

from fmhr_b200 import repose, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "repose_v1.npz")


def test_subdivide_weight_matches_the_reference():
    g = np.load(GOLDEN)
    w1 = repose.subdivide_weight(g["w0"].astype(np.float64), g["f1"])
    assert w1.shape == g["w1"].shape
    assert np.abs(w1 - g["w1"]).max() < 1e-7
    # the face layout the reference relies on is the one synth.subdivide_loop produces
    v0, f0 = synth.base_hand_mesh()
    v1, f1 = synth.subdivide_loop(v0, f0, 1)
    assert np.array_equal(f1.astype(np.int32), g["f1"]) and np.array_equal(f0.astype(np.int32), g["f0"])


def test_subdivide_weight_loop_properties(tmp_path):
    rng = np.random.default_rng(3)
    v0, f0 = synth.base_hand_mesh()
    w0 = rng.dirichlet(np.ones(16) * 0.3, size=v0.shape[0])
    v, f, w = repose.subdivide_weight_loop(w0, v0, f0, iterations=3)
    assert v.shape == (49281, 3) and f.shape == (98432, 3) and w.shape == (49281, 16)   # sub3 counts of the reference
    assert np.allclose(w.sum(1), 1.0, atol=1e-12) and w.min() >= 0.0                       # still a partition of unity
    assert np.array_equal(w[: v0.shape[0]], w0)                                           # old vertices keep their weights
    # every inserted vertex of the LAST round carries the mean of its edge's end points
    v2, f2 = synth.subdivide_loop(v0, f0, 2)
    q = f.reshape(-1, 4, 3)
    a, b, ab = q[:, 0, 0], q[:, 1, 1], q[:, 0, 1]
    assert a.max() < v2.shape[0] and ab.min() >= v2.shape[0]
    assert np.allclose(w[ab], 0.5 * (w[a] + w[b]), atol=1e-12)
    out = repose.save_sub_weights(str(tmp_path / "mano_weight_sub3.pkl"), {"right": (w0, v0, f0)})
    with open(tmp_path / "mano_weight_sub3.pkl", "rb") as fh:
        pk = pickle.load(fh)
    assert np.array_equal(pk["right"]["faces"], out["right"]["faces"]) and pk["right"]["weights"].shape == (49281, 16)

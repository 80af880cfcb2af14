"""Setup geometry of the T-pose export (repose.py:14-41 of the reference), host side, numpy.

The reference subdivides the MANO skinning weights alongside the mesh (three rounds of Loop subdivision,
`mano/mano_weight_sub3.pkl`): every vertex that a round inserts on an edge takes the mean of the weights of the edge's
two end points.  `subdivide_weight` relies on the face layout of trimesh.remesh.subdivide_loop - the four children of
parent face (a, b, c) are consecutive: (a, ab, ca), (ab, b, bc), (ca, bc, c), (ab, bc, ca) - which is also the layout of
fmhr_b200.synth.subdivide_loop; it does not depend on the ORDER in which the new vertices are numbered.
The linear-blend-skinning T-pose itself (repose.py:43-99) needs the licensed MANO model and is out of scope.
"""
import pickle

import numpy as np

from .synth import subdivide_loop


def subdivide_weight(weights, faces):
    """repose.py:14-24, vectorised: weights [V0,J] of the vertices before the round, faces [4F,3] after it ->
    weights [V1,J].  (The reference walks the parent faces in order and overwrites; all writes to one midpoint carry the
    same value, so the order is immaterial.)"""
    weights = np.asarray(weights)
    faces = np.asarray(faces)
    if faces.shape[0] % 4:
        raise RuntimeError("fmhr_b200.repose: faces must hold four consecutive children per parent face")
    out = np.zeros((int(faces.max()) + 1, weights.shape[1]), dtype=np.float64)
    out[:weights.shape[0]] = weights
    q = faces.reshape(-1, 4, 3)
    a, b, c = q[:, 0, 0], q[:, 1, 1], q[:, 2, 2]          # the parent's corners
    ab, ca, bc = q[:, 0, 1], q[:, 0, 2], q[:, 1, 2]       # the three inserted vertices
    if max(a.max(), b.max(), c.max()) >= weights.shape[0]:
        raise RuntimeError("fmhr_b200.repose: the first child of every parent face must start at an old vertex")
    out[ab] = (out[a] + out[b]) / 2
    out[ca] = (out[a] + out[c]) / 2
    out[bc] = (out[b] + out[c]) / 2
    return out


def subdivide_weight_loop(weights, vertices, faces, iterations=3):
    """repose.py:26-30: `iterations` rounds of Loop subdivision carrying the skinning weights along."""
    for _ in range(iterations):
        vertices, faces = subdivide_loop(vertices, faces, 1)
        weights = subdivide_weight(weights, faces)
    return vertices, faces, weights


def save_sub_weights(path, hands):
    """repose.py:32-41: `hands` = {'left' | 'right': (lbs_weights [778,16], v_template [778,3], faces [1538,3])} ->
    pickle {'left' | 'right': {'faces': [F,3], 'weights': [V,16]}} as mesh_sfs_optim.py:347-365 reads it."""
    out = {}
    for hand_type, (w, v, f) in hands.items():
        _, faces_tmp, new_weights = subdivide_weight_loop(np.asarray(w), np.asarray(v), np.asarray(f, dtype=np.int64), 3)
        out[hand_type] = {"faces": faces_tmp, "weights": new_weights}
    with open(path, "wb") as fh:
        pickle.dump(out, fh)
    return out

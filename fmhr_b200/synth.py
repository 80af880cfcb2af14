"""Seeded synthetic inputs of the shapes BASELINE.json names (host-side setup; numpy only).

The reference's inputs are licensed (MANO, InterHand2.6M) or absent (demo mesh), so the
bench and the tests run on a procedural stand-in with the SAME counts and conventions:

* mesh: a hand-sized open "mitten" with exactly MANO's topology counts (778 verts / 1538 faces, one
  16-edge wrist boundary), Loop-subdivided k times -> 49,281 / 98,432 at k=3 and 196,993 / 393,728
  at k=4 (mesh_sfs_optim.py:82,106 use trimesh.remesh.subdivide_loop(iterations=3)); 4 children per
  parent face stay consecutive as in trimesh.
* cameras: the reference's clip convention (get_data.py:62-76,96-97): row-vector matrices stored
  transposed, z_clip = -0.1, w_clip = z_cam, no y flip.
"""
import math

import numpy as np


# ------------------------------------------------------------------------------------------------
# mesh
# ------------------------------------------------------------------------------------------------
def _ring_sizes(total=778, boundary=16):
    """Ring vertex counts from the wrist boundary to a single tip vertex, summing to `total`."""
    sizes = [boundary, 24, 32, 40]
    tail = [40, 32, 24, 16, 8, 1]
    body = total - sum(sizes) - sum(tail)
    n48 = body // 48
    sizes += [48] * n48
    rem = body - 48 * n48
    if rem:
        sizes.append(rem)  # one odd-sized ring absorbs the remainder
    sizes += tail
    assert sum(sizes) == total
    return sizes


def _zip_rings(start_a, na, start_b, nb):
    """Triangulate the band between ring A (na verts) and ring B (nb verts), both ccw, by angle."""
    faces = []
    if nb == 1:
        for i in range(na):
            faces.append((start_a + i, start_a + (i + 1) % na, start_b))
        return faces
    i = j = 0
    while i < na or j < nb:
        # next angle on each ring
        ta = (i + 1) / na
        tb = (j + 1) / nb
        a0, b0 = start_a + i % na, start_b + j % nb
        if j >= nb or (i < na and ta <= tb):
            faces.append((a0, start_a + (i + 1) % na, b0))
            i += 1
        else:
            faces.append((a0, start_b + (j + 1) % nb, b0))
            j += 1
    return faces


def base_hand_mesh(mirror=False):
    """778-vertex / 1538-face open mitten, world units (~0.4 long, like the x2-scaled demo hand)."""
    sizes = _ring_sizes()
    R = len(sizes)
    verts = []
    starts = []
    for r, n in enumerate(sizes):
        starts.append(len(verts))
        t = r / (R - 1)  # 0 wrist .. 1 finger tip
        y = -0.2 + 0.4 * t
        # cross-section radii: wrist narrow, palm wide, fingers taper
        rx = 0.035 + 0.065 * math.sin(math.pi * min(t * 1.15, 1.0)) ** 0.8
        rz = 0.018 + 0.022 * math.sin(math.pi * min(t * 1.1, 1.0))
        if n == 1:
            verts.append((0.0, y + 0.004, 0.0))
            continue
        for i in range(n):
            a = 2.0 * math.pi * (i + 0.5 * (r % 2)) / n
            # four finger-like lobes towards the tip, a thumb bulge on one side
            lobes = 1.0 + 0.18 * max(t - 0.5, 0.0) * 2.0 * math.cos(4.0 * a)
            thumb = 0.03 * math.exp(-((t - 0.35) / 0.12) ** 2) * max(math.cos(a), 0.0) ** 2
            x = (rx * lobes + thumb) * math.cos(a)
            z = rz * lobes * math.sin(a) + 0.01 * math.sin(3.0 * math.pi * t)
            verts.append((x, y, z))
    faces = []
    for r in range(R - 1):
        faces += _zip_rings(starts[r], sizes[r], starts[r + 1], sizes[r + 1])
    v = np.asarray(verts, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int64)[:, [0, 2, 1]]  # outward-facing winding
    if mirror:
        v = v * np.array([-1.0, 1.0, 1.0])
        f = f[:, [0, 2, 1]]
    assert v.shape[0] == 778 and f.shape[0] == 1538, (v.shape, f.shape)
    return v, f


def unique_edges(faces, n_verts):
    """Sorted unique undirected edges [E,2] and, per face, the edge id opposite... of edges (1,2),(2,0),(0,1)."""
    e = np.concatenate([faces[:, [1, 2]], faces[:, [2, 0]], faces[:, [0, 1]]], axis=0)
    e = np.sort(e, axis=1)
    h = e[:, 0] * n_verts + e[:, 1]
    uh, inv = np.unique(h, return_inverse=True)
    edges = np.stack([uh // n_verts, uh % n_verts], axis=1)
    return edges, inv.reshape(3, -1).T  # face -> edge ids of (12, 20, 01)


def subdivide_loop(verts, faces, iterations=1):
    """Loop subdivision (Loop 1987 weights, boundary rules), 4 consecutive children per face."""
    v = np.asarray(verts, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int64)
    for _ in range(iterations):
        nv = v.shape[0]
        edges, f2e = unique_edges(f, nv)
        ne = edges.shape[0]
        # opposite vertices per edge
        opp_sum = np.zeros((ne, 3))
        opp_cnt = np.zeros(ne, dtype=np.int64)
        for k, (ek, ov) in enumerate(((0, 0), (1, 1), (2, 2))):
            np.add.at(opp_sum, f2e[:, ek], v[f[:, ov]])
            np.add.at(opp_cnt, f2e[:, ek], 1)
        interior = opp_cnt == 2
        a, b = v[edges[:, 0]], v[edges[:, 1]]
        odd = 0.5 * (a + b)
        odd[interior] = 0.375 * (a[interior] + b[interior]) + 0.125 * opp_sum[interior]
        # even vertices
        nb_sum = np.zeros((nv, 3))
        nb_cnt = np.zeros(nv, dtype=np.int64)
        np.add.at(nb_sum, edges[:, 0], b)
        np.add.at(nb_sum, edges[:, 1], a)
        np.add.at(nb_cnt, edges[:, 0], 1)
        np.add.at(nb_cnt, edges[:, 1], 1)
        k = np.maximum(nb_cnt, 1).astype(np.float64)
        beta = (1.0 / k) * (0.625 - (0.375 + 0.25 * np.cos(2.0 * np.pi / k)) ** 2)
        even = (1.0 - k * beta)[:, None] * v + beta[:, None] * nb_sum
        bedges = edges[~interior]
        if bedges.shape[0]:
            bsum = np.zeros((nv, 3))
            isb = np.zeros(nv, dtype=bool)
            np.add.at(bsum, bedges[:, 0], v[bedges[:, 1]])
            np.add.at(bsum, bedges[:, 1], v[bedges[:, 0]])
            isb[bedges.reshape(-1)] = True
            even[isb] = 0.75 * v[isb] + 0.125 * bsum[isb]
        m = f2e + nv  # midpoints of edges (12, 20, 01) -> m[:,0] between v1,v2 ; m[:,1] between v2,v0 ; m[:,2] between v0,v1
        f = np.stack([
            f[:, 0], m[:, 2], m[:, 1],
            m[:, 2], f[:, 1], m[:, 0],
            m[:, 1], m[:, 0], f[:, 2],
            m[:, 2], m[:, 0], m[:, 1]], axis=1).reshape(-1, 3)
        v = np.concatenate([even, odd], axis=0)
    return v, f


def hand_mesh(subdiv=3, hands=1, seed=0, noise=2e-4):
    """float32 vertices [V,3] / int32 faces [F,3]; two hands are concatenated with offset face indices
    (mesh_sfs_optim.py:75-88)."""
    vs, fs = [], []
    off = 0
    for h in range(hands):
        v, f = base_hand_mesh(mirror=(h == 1))
        v, f = subdivide_loop(v, f, subdiv)
        if h == 1:
            v = v + np.array([0.24, 0.0, 0.02])
        vs.append(v)
        fs.append(f + off)
        off += v.shape[0]
    v = np.concatenate(vs, 0)
    f = np.concatenate(fs, 0)
    rng = np.random.default_rng(seed)
    v = v + rng.normal(0.0, noise, v.shape)
    return v.astype(np.float32), f.astype(np.int32)


# ------------------------------------------------------------------------------------------------
# cameras (reference convention)
# ------------------------------------------------------------------------------------------------
def make_cameras(n, H, W, center, extent=0.4, radius=1.9, frac=0.45, seed=1):
    """n golden-spiral cameras looking at `center`.  Returns (w2cs[n,4,4], projs[n,4,4]) float32, both
    TRANSPOSED for row-vector use exactly like get_data.py:96-97."""
    rng = np.random.default_rng(seed)
    f = frac * H * radius / extent
    w2cs, projs = [], []
    ga = math.pi * (3.0 - math.sqrt(5.0))
    for i in range(n):
        zc = 1.0 - 2.0 * (i + 0.5) / n
        rr = math.sqrt(max(0.0, 1.0 - zc * zc))
        th = ga * i
        d = np.array([rr * math.cos(th), zc * 0.6, rr * math.sin(th)])  # flatten towards the equator
        d /= np.linalg.norm(d)
        eye = center + radius * (1.0 + 0.05 * rng.uniform(-1, 1)) * d
        fwd = center - eye
        fwd /= np.linalg.norm(fwd)
        up = np.array([0.0, 1.0, 0.0]) if abs(fwd[1]) < 0.95 else np.array([1.0, 0.0, 0.0])
        right = np.cross(fwd, up)
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        Rm = np.stack([right, down, fwd], 0)  # world -> cam (x right, y down, z forward)
        w2c = np.eye(4)
        w2c[:3, :3] = Rm
        w2c[:3, 3] = -Rm @ eye
        cx = W / 2.0 + rng.uniform(-4, 4)
        cy = H / 2.0 + rng.uniform(-4, 4)
        proj = np.zeros((4, 4))
        proj[0, 0] = f / (W / 2.0)
        proj[0, 2] = cx / (W / 2.0) - 1.0
        proj[1, 1] = f / (H / 2.0)
        proj[1, 2] = cy / (H / 2.0) - 1.0
        proj[2, 3] = -0.1
        proj[3, 2] = 1.0
        w2cs.append(w2c.T)
        projs.append(proj.T)
    # np.stack of .T views yields Fortran-ordered memory; downstream code hands raw pointers to the C ABI
    return (np.ascontiguousarray(np.stack(w2cs), dtype=np.float32),
            np.ascontiguousarray(np.stack(projs), dtype=np.float32))


def smooth_albedo(n_verts, faces, seed=3, steps=5):
    """Per-vertex BGR albedo U(0.3,0.9) smoothed by `steps` uniform-Laplacian averaging passes."""
    rng = np.random.default_rng(seed)
    alb = rng.uniform(0.3, 0.9, (n_verts, 3))
    edges, _ = unique_edges(faces.astype(np.int64), n_verts)
    deg = np.zeros(n_verts)
    np.add.at(deg, edges[:, 0], 1)
    np.add.at(deg, edges[:, 1], 1)
    for _ in range(steps):
        s = np.zeros_like(alb)
        np.add.at(s, edges[:, 0], alb[edges[:, 1]])
        np.add.at(s, edges[:, 1], alb[edges[:, 0]])
        alb = 0.5 * alb + 0.5 * s / np.maximum(deg, 1)[:, None]
    return alb.astype(np.float32)


def sh_lighting(n, seed=3):
    rng = np.random.default_rng(seed + 100)
    base = np.array([0.8, 0.1, 0.3, 0.1, 0.0, 0.0, 0.05, 0.0, 0.0])
    return (base[None] + rng.normal(0.0, 0.02, (n, 9))).astype(np.float32)


# weights / learning rates per conf file
CONF = {
    "ih_sfs": dict(sfs_weight=30.0, lap_weight=1000.0, albedo_weight=2.0, mask_weight=200.0, edge_weight=1e6,
                   delta_weight=5e4, lr=5e-4, albedo_lr=2e-2, sh_lr=5e-3, batch=32),
    "demo_sfs": dict(sfs_weight=50.0, lap_weight=2000.0, albedo_weight=1.0, mask_weight=1000.0, edge_weight=5e5,
                     delta_weight=5e4, lr=1e-4, albedo_lr=5e-3, sh_lr=5e-3, batch=8),
    "ih_sfsseq": dict(sfs_weight=30.0, lap_weight=2000.0, albedo_weight=0.0, mask_weight=1000.0, edge_weight=5e5,
                      delta_weight=1e4, lr=1e-4, albedo_lr=5e-3, sh_lr=5e-3, batch=32),
}

# BASELINE.json configs -> shapes
WORKLOADS = {
    "demo_reduced": dict(n=16, H=256, W=320, subdiv=3, hands=1, conf="demo_sfs"),
    "demo_full": dict(n=16, H=1024, W=1280, subdiv=3, hands=1, conf="demo_sfs"),
    "interhand_48x512x334": dict(n=48, H=512, W=334, subdiv=3, hands=1, conf="ih_sfs"),
    "capture_16x1024x1024": dict(n=16, H=1024, W=1024, subdiv=3, hands=1, conf="demo_sfs"),
    "two_hands_48x512x334": dict(n=48, H=512, W=334, subdiv=3, hands=2, conf="ih_sfs"),
    "stress_128x2048x2048": dict(n=128, H=2048, W=2048, subdiv=4, hands=1, conf="ih_sfs"),
    "tiny": dict(n=4, H=96, W=80, subdiv=1, hands=1, conf="ih_sfs"),
    "small": dict(n=6, H=160, W=128, subdiv=2, hands=1, conf="ih_sfs"),
    # triangles larger than a pixel: the regime where antialias finds silhouette edges (it only inspects the edges
    # of the triangle visible in the pixel, so micropolygon meshes rarely blend)
    "coarse": dict(n=4, H=288, W=240, subdiv=0, hands=1, conf="demo_sfs"),
    "coarse8": dict(n=8, H=288, W=240, subdiv=0, hands=1, conf="demo_sfs"),    # one view per rank on an 8-GPU box
    "coarse16": dict(n=16, H=288, W=240, subdiv=0, hands=1, conf="demo_sfs"),  # two views per rank on an 8-GPU box
}


def build_scene(workload, render_fn, n_views=None, view_offset=0, view_stride=1, camera_seed=1):
    """Assemble one HAM problem.  `render_fn(vertices, faces, albedo, sh, w2cs, projs, H, W)` must return
    (img[n,H,W,3], coverage[n,H,W], aa_coverage[n,H,W]) as numpy float32; the caller passes the product
    renderer (bench) or the oracle renderer (CPU tests).  Views [view_offset::view_stride] of the
    workload's camera set are kept (multi-GPU sharding)."""
    wl = dict(WORKLOADS[workload]) if isinstance(workload, str) else dict(workload)
    n_all, H, W = wl["n"], wl["H"], wl["W"]
    verts, faces = hand_mesh(wl["subdiv"], wl["hands"], seed=0)
    center = verts.mean(0).astype(np.float64)
    extent = float(verts[:, 1].max() - verts[:, 1].min())
    w2cs, projs = make_cameras(n_all, H, W, center, extent=extent, seed=camera_seed)
    sh = sh_lighting(n_all)
    sel = np.arange(view_offset, n_all, view_stride)
    if n_views is not None:
        sel = sel[:n_views]
    w2cs, projs, sh_true = np.ascontiguousarray(w2cs[sel]), np.ascontiguousarray(projs[sel]), sh[sel]
    rng = np.random.default_rng(2)
    target_verts = (verts + rng.normal(0.0, 1e-3, verts.shape)).astype(np.float32)
    alb_true = smooth_albedo(verts.shape[0], faces)
    img, cov, _ = render_fn(target_verts, faces, alb_true, sh_true, w2cs, projs, H, W)
    _, _, valid_mask = render_fn(verts, faces, alb_true, sh_true, w2cs, projs, H, W)
    # initial state as after the reference's init + phase A (mesh_sfs_optim.py:124-240): near-mean albedo, fitted SH
    # (not exactly constant: on a constant albedo the uniform Laplacian is pure rounding noise and its
    #  sub-gradient direction is undefined, in the reference as well; phase A leaves it non-constant)
    alb0 = (0.7 * alb_true.mean(0, keepdims=True) + 0.3 * alb_true).astype(np.float32)
    sh0 = (sh_true + np.random.default_rng(4).normal(0, 0.01, sh_true.shape)).astype(np.float32)
    conf = dict(CONF[wl["conf"]])
    return dict(vertices=verts, faces=faces, w2cs=w2cs, projs=projs, imgs=img, masks=cov, valid_masks=valid_mask,
                albedo=alb0, sh_coeffs=sh0, H=H, W=W, conf=conf, n_total=n_all, workload=wl)


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs[0]: the reference's demo capture (demo_data/1) through the loader convention
# ------------------------------------------------------------------------------------------------
def demo_cameras(world_mats, scale_mats, cap_res=(1280, 1024)):
    """The camera convention of get_demo_data (get_data.py:62-76,96-97) without OpenCV: P = world_mat @ scale_mat is
    split into K [R | t] (RQ decomposition, what cv2.decomposeProjectionMatrix does inside load_K_Rt_from_P,
    models/utils.py:29-47), K is normalised by K[2,2], w2c = inverse of the camera-to-world pose, and the projection is
    rewritten to the reference's clip convention (x, y scaled to NDC at the capture resolution, z_clip = -0.1,
    w_clip = z_cam).  Returns (w2cs[n,4,4], projs[n,4,4]) float32, TRANSPOSED for row-vector use."""
    w2cs, projs = [], []
    for Wm, Sm in zip(np.asarray(world_mats, dtype=np.float32), np.asarray(scale_mats, dtype=np.float32)):
        P = (Wm @ Sm)[:3].astype(np.float64)
        M = P[:, :3]
        # RQ decomposition M = K R with K upper triangular, positive diagonal
        flip = np.flipud(np.eye(3))
        q, r = np.linalg.qr((flip @ M).T)
        K = flip @ r.T @ flip
        R = flip @ q.T
        sgn = np.diag(np.sign(np.diag(K)))
        K, R = K @ sgn, sgn @ R
        if np.linalg.det(R) < 0:
            R = -R
            K = -K  # cancels in K / K[2,2] below except for the overall sign of P, which is projective
        c = -np.linalg.solve(M, P[:, 3])  # camera centre: P [c, 1]^T = 0
        K = K / K[2, 2]
        pose = np.eye(4, dtype=np.float32)
        pose[:3, :3] = R.T
        pose[:3, 3] = c
        w2c = np.linalg.inv(pose)
        proj = np.eye(4)
        proj[:3, :3] = K
        proj[0, 0] = proj[0, 0] / (cap_res[0] / 2.)
        proj[0, 2] = proj[0, 2] / (cap_res[0] / 2.) - 1.
        proj[1, 1] = proj[1, 1] / (cap_res[1] / 2.)
        proj[1, 2] = proj[1, 2] / (cap_res[1] / 2.) - 1.
        proj[2, 2] = 0.
        proj[2, 3] = -0.1
        proj[3, 2] = 1.
        proj[3, 3] = 0.
        w2cs.append(w2c.astype(np.float32).T)
        projs.append(proj.astype(np.float32).T)
    return (np.ascontiguousarray(np.stack(w2cs), dtype=np.float32),
            np.ascontiguousarray(np.stack(projs), dtype=np.float32))


def fit_hand_to_keypoints(kp, subdiv=3, seed=0, noise=2e-4):
    """The synthetic right-hand mesh posed on 21 triangulated keypoints (MediaPipe order: 0 wrist, 5 / 9 / 17 index /
    middle / little knuckle, 12 middle finger tip): wrist -> middle finger tip is the long axis, the knuckle line gives
    the palm plane.  Stands in for the demo's MANO fit, which is not shipped."""
    kp = np.asarray(kp, dtype=np.float64)
    v, f = base_hand_mesh()
    v, f = subdivide_loop(v, f, subdiv)
    wrist, tip = kp[0], kp[12]
    ay = tip - wrist
    length = np.linalg.norm(ay)
    ay = ay / length
    ax = kp[5] - kp[17]
    ax = ax - ay * float(ax @ ay)
    ax = ax / np.linalg.norm(ax)
    az = np.cross(ax, ay)
    s = length / 0.4  # the base mesh spans y in [-0.2, 0.2]
    local = v + np.array([0.0, 0.2, 0.0])
    world = wrist[None] + s * (local[:, 0:1] * ax[None] + local[:, 1:2] * ay[None] + local[:, 2:3] * az[None])
    rng = np.random.default_rng(seed)
    world = world + rng.normal(0.0, noise, world.shape)
    return world.astype(np.float32), f.astype(np.int32)


def demo_scene(fixture, hand="right", subdiv=3):
    """Config 1 on the real demo capture: `fixture` = tests/golden/demo1_320x256.npz (or the dict loaded from it; made
    by the committed fixture generator from the reference's demo_data/1).  Images / masks / gray images as the loader
    returns them (uint8 / 255, BGR, zero outside the mask; mask = byte > 127), cameras in the loader's transposed
    clip convention, the synthetic hand posed on the capture's 3-D keypoints, conf/demo_sfs.conf weights.
    valid_masks, sh_coeffs and albedo are the job of the HAM initialisation (mesh_sfs_optim.py:124-177); placeholders
    (segmentation mask, zero lighting, grey albedo) are returned so the dict has the usual keys."""
    fx = np.load(fixture) if isinstance(fixture, str) else fixture
    kp = np.asarray(fx["keypoints_3d"])
    verts, faces = fit_hand_to_keypoints(kp[21:42] if hand == "right" else kp[0:21], subdiv)
    imgs = fx["imgs_u8"].astype(np.float32) / np.float32(255.0)
    gray = fx["gray_u8"].astype(np.float32) / np.float32(255.0)
    masks = (fx["masks_u8"] > 127).astype(np.float32)
    n, H, W = masks.shape
    w2cs, projs = demo_cameras(fx["world_mats"], fx["scale_mats"], tuple(int(x) for x in fx["cap_res"]))
    return dict(vertices=verts, faces=faces, w2cs=w2cs, projs=projs, imgs=imgs, grayimgs=gray, masks=masks,
                valid_masks=masks.copy(), albedo=np.full((verts.shape[0], 3), 0.5, dtype=np.float32),
                sh_coeffs=np.zeros((n, 9), dtype=np.float32), H=H, W=W, conf=dict(CONF["demo_sfs"]), n_total=n,
                workload=dict(n=n, H=H, W=W, subdiv=subdiv, hands=1, conf="demo_sfs"))

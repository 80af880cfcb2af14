"""ORACLE (test infrastructure) — run the reference's OWN script, unchanged, and keep its results as a golden fixture.

Runs only in the build container (needs /root/reference).  /root/reference/mesh_sfs_optim.py is imported verbatim and its
main() is executed from the first line to the last (initialisation :124-177, phase A :193-240, phase B :242-317, result
files :321-343, T-pose export :346-388) on a tiny synthetic capture:

  * `nvdiffrast.torch`   -> oracle.raster (the CPU restatement of rasterize / interpolate / antialias; the real
                            dependency is absent and needs a GL context),
  * `models.utils`       -> the reference's own file (imported verbatim),
  * `trimesh`, `pyhocon`, `get_data`, `repose` -> thin stand-ins for the I/O the script does through them (OBJ read /
                            write, `key = value` conf file, the loader's tensor tuple, identity pose for the T-pose),
  * `Tensor.cuda()`      -> identity (this container has no GPU; the script is device-agnostic otherwise).

The permutations come from torch.manual_seed(SEED) + the script's own torch.randperm calls.  Inputs and outputs go to
tests/golden/refrun_v1.npz; tests replay the same run with oracle.ham (CPU: pins the restated loop to the real script)
and with fmhr_b200.ham.HamOptimizer (GPU: the product against the reference script's results).

    python -m oracle.gen_reference_run
"""
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch

from .gen_golden import OUT, REF, import_reference

SEED = 123
CONF = dict(data_type="demo", input_mesh_dire="mano_out", out_mesh_dire="demo_sfs", num=4, w=80, h=96, epoch_albedo=2,
            epoch_sfs=2, sfs_weight=50, albedo_weight=1, lap_weight=2000, mask_weight=1000, edge_weight=500000,
            delta_weight=50000, degree=3, batch=2, albedo_lr=0.005, lr=0.0001, sh_lr=0.005)


def tube_mesh(rings=6, segs=8):
    """A small open tube with a closed tip (hand-sized, like the synthetic mitten but ~50 vertices), so that the three
    rounds of Loop subdivision the script applies (mesh_sfs_optim.py:82) stay CPU-sized."""
    verts, faces = [], []
    for r in range(rings):
        t = r / (rings - 1)
        rad = 0.06 * (1.0 - 0.55 * t * t)
        for s in range(segs):
            a = 2 * np.pi * (s + 0.5 * (r % 2)) / segs
            verts.append((rad * np.cos(a) * 1.4, -0.15 + 0.3 * t, rad * np.sin(a) * 0.8))
    tip = len(verts)
    verts.append((0.0, 0.17, 0.0))
    for r in range(rings - 1):
        for s in range(segs):
            a0, a1 = r * segs + s, r * segs + (s + 1) % segs
            b0, b1 = a0 + segs, a1 + segs
            faces += [(a0, b0, a1), (a1, b0, b1)] if r % 2 == 0 else [(a0, b1, a1), (a0, b0, b1)]
    top = (rings - 1) * segs
    for s in range(segs):
        faces.append((top + s, tip, top + (s + 1) % segs))
    return np.asarray(verts, dtype=np.float64), np.asarray(faces, dtype=np.int64)


def make_inputs():
    """Base mesh + the loader tuple (imgs BGR [num,H,W,3], grayimgs, masks, w2cs, projs) of a 4-view synthetic capture."""
    from fmhr_b200 import synth
    from . import ham as oham
    v0, f0 = tube_mesh()
    v3, f3 = synth.subdivide_loop(v0, f0, 3)
    v3, f3 = v3.astype(np.float32), f3.astype(np.int32)
    H, W, n = CONF["h"], CONF["w"], CONF["num"]
    w2cs, projs = synth.make_cameras(n, H, W, v3.mean(0).astype(np.float64), extent=0.32, radius=1.9, frac=0.6, seed=11)
    rng = np.random.default_rng(5)
    target = (v3 + rng.normal(0, 1.5e-3, v3.shape)).astype(np.float32)
    alb = synth.smooth_albedo(v3.shape[0], f3, seed=9)
    sh = synth.sh_lighting(n, seed=9)
    img, cov, _ = oham.render_views(target, f3, alb, sh, w2cs, projs, H, W)
    img_u8 = np.clip(np.rint(img.astype(np.float64) * 255), 0, 255).astype(np.uint8)
    gray_u8 = np.clip(np.rint(img_u8.astype(np.float64).mean(-1)), 0, 255).astype(np.uint8)
    return dict(base_verts=v0.astype(np.float32), base_faces=f0.astype(np.int32), imgs_u8=img_u8, gray_u8=gray_u8,
                masks_u8=(cov > 0).astype(np.uint8) * 255, w2cs=w2cs, projs=projs)


class _Conf:
    def __init__(self, d):
        self.d = d

    def get_string(self, k):
        return str(self.d[k])

    def get_int(self, k):
        return int(self.d[k])

    def get_float(self, k):
        return float(self.d[k])


def read_conf(path):
    """`key = value` lines (the flat HOCON the reference's conf/*.conf files are)."""
    d = {}
    for line in open(path):
        if "=" in line:
            k, v = line.split("=", 1)
            d[k.strip()] = v.strip()
    return _Conf(d)


def install_stubs(inputs, dr_module):
    """sys.modules stand-ins for the script's I/O dependencies; returns the dict of modules installed."""
    from fmhr_b200 import export, synth

    class Trimesh:
        def __init__(self, vertices=None, faces=None, process=False, maintain_order=True):
            self.vertices, self.faces = np.asarray(vertices), np.asarray(faces)

        def export(self, path):
            export.save_obj_mesh(path, self.vertices, self.faces)

    def load(path, process=False, maintain_order=True):
        v, _, f = export.load_obj(path)
        return Trimesh(v, f)

    trimesh = types.ModuleType("trimesh")
    trimesh.Trimesh, trimesh.load = Trimesh, load
    remesh = types.ModuleType("trimesh.remesh")
    remesh.subdivide_loop = lambda v, f, iterations=1: synth.subdivide_loop(np.asarray(v), np.asarray(f), iterations)
    trimesh.remesh = remesh
    pyhocon = types.ModuleType("pyhocon")
    pyhocon.ConfigFactory = types.SimpleNamespace(parse_file=read_conf)

    def get_demo_data(data_path, scan_id, num, res=(1280, 1024), return_ray=False, with_mask=True, use_mask=False):
        assert (res[1], res[0]) == inputs["imgs_u8"].shape[1:3] and num == inputs["imgs_u8"].shape[0]
        t = torch.from_numpy
        # the loader's arithmetic (get_data.py:91-99): uint8 / 255. in float64, then .float()
        return (t(inputs["imgs_u8"] / 255.).float(), t(inputs["gray_u8"] / 255.).float(),
                t((inputs["masks_u8"] > 127).astype(np.float32)), t(inputs["w2cs"].copy()), t(inputs["projs"].copy()))

    get_data = types.ModuleType("get_data")
    get_data.get_demo_data, get_data.get_interhand_data = get_demo_data, None
    repose = types.ModuleType("repose")
    repose.lbs_tpose = lambda pose, shape, weights, verts, hand_type="right": verts  # (needs the licensed MANO model)
    nvd = types.ModuleType("nvdiffrast")
    nvd.torch = dr_module
    mods = {"trimesh": trimesh, "trimesh.remesh": remesh, "pyhocon": pyhocon, "get_data": get_data, "repose": repose,
            "nvdiffrast": nvd, "nvdiffrast.torch": dr_module}
    sys.modules.update(mods)
    return mods


def prepare_workdir(workdir, inputs):
    """The files main() reads: conf, <out>/mano_out/1.obj + 1.pt, mano/mano_weight_sub3.pkl."""
    from fmhr_b200 import export
    from fmhr_b200 import repose as frepose
    os.makedirs(os.path.join(workdir, "demo_out", "mano_out"), exist_ok=True)
    os.makedirs(os.path.join(workdir, "mano"), exist_ok=True)
    conf = os.path.join(workdir, "demo_sfs.conf")
    with open(conf, "w") as f:
        for k, v in CONF.items():
            f.write("%s = %s\n" % (k, v))
    export.save_obj_mesh(os.path.join(workdir, "demo_out", "mano_out", "1.obj"), inputs["base_verts"], inputs["base_faces"])
    torch.save([{"type": "right", "pose": torch.zeros(48), "shape": torch.zeros(10), "trans": torch.zeros(3), "scale": 1}],
               os.path.join(workdir, "demo_out", "mano_out", "1.pt"))
    rng = np.random.default_rng(2)
    w0 = rng.dirichlet(np.ones(16) * 0.3, size=inputs["base_verts"].shape[0])
    frepose.save_sub_weights(os.path.join(workdir, "mano", "mano_weight_sub3.pkl"),
                             {"right": (w0, inputs["base_verts"].astype(np.float64), inputs["base_faces"].astype(np.int64))})
    return conf


def run_reference_script(script_dir, inputs, dr_module, workdir, to_device=None):
    """Imports <script_dir>/mesh_sfs_optim.py verbatim and runs main() in `workdir`.  Returns the result files' content."""
    from fmhr_b200 import export
    saved_modules = {k: sys.modules.get(k) for k in ("trimesh", "trimesh.remesh", "pyhocon", "get_data", "repose",
                                                     "nvdiffrast", "nvdiffrast.torch", "mesh_sfs_optim")}
    install_stubs(inputs, dr_module)
    conf = prepare_workdir(workdir, inputs)
    cwd, saved_cuda = os.getcwd(), torch.Tensor.cuda
    sys.path.insert(0, script_dir)
    try:
        if to_device is None:
            torch.Tensor.cuda = lambda self, *a, **k: self
        os.chdir(workdir)
        sys.modules.pop("mesh_sfs_optim", None)
        import mesh_sfs_optim
        torch.manual_seed(SEED)
        mesh_sfs_optim.main(conf, 1, "demo_data")
    finally:
        os.chdir(cwd)
        torch.Tensor.cuda = saved_cuda
        sys.path.remove(script_dir)
        for k, m in saved_modules.items():
            if m is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = m
    out = os.path.join(workdir, "demo_out", "demo_sfs", "1")
    pt = torch.load(os.path.join(out, "1.pt"))
    v, _, f = export.load_obj(os.path.join(out, "1.obj"))
    vo, _, _ = export.load_obj(os.path.join(out, "ori_1.obj"))
    vt, _, _ = export.load_obj(os.path.join(out, "1_right_tpose.obj"))
    vc, cc, fc = export.load_obj(os.path.join(out, "1_c.obj"))
    files = sorted(os.listdir(out)) + sorted(os.listdir(os.path.join(out, "rerender")))
    return dict(sh_coeff=pt["sh_coeff"].detach().cpu().numpy(), albedo=pt["albedo"].detach().cpu().numpy(),
                vertices=v.astype(np.float32), faces=f.astype(np.int32), ori_vertices=vo.astype(np.float32),
                tpose_vertices=vt.astype(np.float32), color_rgb=cc.astype(np.float32), color_faces=fc.astype(np.int32),
                files=np.array(files))


def main():
    import_reference()  # the reference's models/ package (stubs for skimage / plyfile), verbatim
    from . import raster as oraster
    inputs = make_inputs()
    with tempfile.TemporaryDirectory() as tmp:
        res = run_reference_script(REF, inputs, oraster, tmp)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "refrun_v1.npz")
    np.savez_compressed(path, seed=SEED, conf=np.array(sorted("%s=%s" % kv for kv in CONF.items())),
                        **{"in_" + k: v for k, v in inputs.items()}, **{"out_" + k: v for k, v in res.items()})
    print(path, os.path.getsize(path), "bytes;", res["vertices"].shape, "files:", list(res["files"]))


if __name__ == "__main__":
    main()

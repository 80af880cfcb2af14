"""ORACLE (test infrastructure) — tests/golden/demo1_320x256.npz: BASELINE.json configs[0] on the REAL demo fixture.

Runs only in the build container (needs /root/reference).  Follows get_demo_data (get_data.py:49-99) on
demo_data/1 with the reference's OWN load_K_Rt_from_P (models/utils.py:29-47, imported verbatim):
  * cameras: P = world_mat @ scale_mat, (K, w2c) = load_K_Rt_from_P(P[:3]), projection fix-up of get_data.py:66-73 at
    the capture resolution 1280x1024 (the NDC mapping does not depend on the raster size), both matrices transposed
    (get_data.py:96-97);
  * images: cv2.imread (BGR), zeroed outside the mask (>127.5), gray = cv2.cvtColor, then cv2.resize to 320x256
    (BASELINE.json: "reference CPU path at reduced resolution"), masks nearest-resized - all kept as the uint8 the
    loader divides by 255;
  * the 42 triangulated 3-D keypoints (pose_optim.py output shipped with the demo), used to place the synthetic
    right-hand mesh (the demo's MANO fit is not shipped).

    python -m oracle.gen_demo_fixture
"""
import os

import cv2
import numpy as np

from .gen_golden import OUT, REF, import_reference

NUM, CAP_RES, RES = 16, (1280, 1024), (320, 256)


def main():
    ru, _ = import_reference()
    d = os.path.join(REF, "demo_data", "1")
    cam = np.load(os.path.join(d, "camera", "param.npz"))
    world_mats = np.stack([cam["world_mat_%d" % i].astype(np.float32) for i in range(NUM)])
    scale_mats = np.stack([cam["scale_mat_%d" % i].astype(np.float32) for i in range(NUM)])
    imgs, grays, masks, w2cs, projs = [], [], [], [], []
    for i in range(NUM):
        P = world_mats[i] @ scale_mats[i]
        proj, w2c = ru.load_K_Rt_from_P(P[:3])
        proj[0, 0] = proj[0, 0] / (CAP_RES[0] / 2.)
        proj[0, 2] = proj[0, 2] / (CAP_RES[0] / 2.) - 1.
        proj[1, 1] = proj[1, 1] / (CAP_RES[1] / 2.)
        proj[1, 2] = proj[1, 2] / (CAP_RES[1] / 2.) - 1.
        proj[2, 2] = 0.
        proj[2, 3] = -0.1
        proj[3, 2] = 1.
        proj[3, 3] = 0.
        projs.append(proj.astype(np.float32))
        w2cs.append(w2c.astype(np.float32))
        img = cv2.imread(os.path.join(d, "img", "%02d.png" % i))
        mask = cv2.imread(os.path.join(d, "mask", "%02d.png" % i))[:, :, 0]
        mask = (mask > 127.5).astype(np.float32)
        img[mask == 0] = 0
        gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        imgs.append(cv2.resize(img, RES))
        grays.append(cv2.resize(gray, RES))
        masks.append((cv2.resize(mask, RES, interpolation=cv2.INTER_NEAREST) > 0).astype(np.uint8) * 255)
    out = dict(
        imgs_u8=np.stack(imgs), gray_u8=np.stack(grays), masks_u8=np.stack(masks),
        w2cs=np.ascontiguousarray(np.stack(w2cs).transpose(0, 2, 1)),
        projs=np.ascontiguousarray(np.stack(projs).transpose(0, 2, 1)),
        world_mats=world_mats, scale_mats=scale_mats,
        keypoints_3d=np.loadtxt(os.path.join(d, "keypoints_3d_1.xyz")).astype(np.float32),
        cap_res=np.array(CAP_RES), res=np.array(RES))
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "demo1_320x256.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes; mask coverage", float((out["masks_u8"] > 0).mean()))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generates tests/golden/camera.npz by running the reference's own utils/camera.py (imported from /root/reference in
the build container; needs cv2, which that file imports).  Usage: python tests/golden/make_golden_camera.py"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MOCO_REFERENCE", "/root/reference")


def main():
    spec = importlib.util.spec_from_file_location("ref_camera", os.path.join(REF, "utils", "camera.py"))
    cam_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cam_mod)
    rng = np.random.default_rng(7)
    out = {}
    for name, (H, W) in {"a": (12, 20), "b": (33, 17)}.items():
        K = np.array([[W * 1.3, 0.0, W / 2 - 0.25], [0.0, W * 1.1, H / 2 + 0.5], [0.0, 0.0, 1.0]])
        # random rotation (QR) + translation
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        c2w = np.eye(4)
        c2w[:3, :3] = q
        c2w[:3, 3] = rng.normal(size=3) * 2.0
        aabb = np.array([[-0.5, -1.0, -0.3], [0.5, 1.0, 0.3]])
        verts = cam_mod.convert_AABB_to_verts(aabb)
        cam = cam_mod.Camera((H, W), K)
        cam.c2w = c2w
        idx = 0.375
        rays = cam.make_rays(verts, idx)
        rays_o_cam, rays_d_cam = cam_mod.gen_rays(cam.directions, None)
        out.update({f"{name}_H": H, f"{name}_W": W, f"{name}_K": K, f"{name}_c2w": c2w, f"{name}_verts": verts,
                    f"{name}_idx": idx, f"{name}_rays": rays.numpy(), f"{name}_dirs": cam.directions.numpy(),
                    f"{name}_rays_d_cam": rays_d_cam.numpy()})
    np.savez_compressed(os.path.join(HERE, "camera.npz"), **out)
    print("wrote camera.npz", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()

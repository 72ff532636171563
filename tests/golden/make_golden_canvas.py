#!/usr/bin/env python
"""Generates tests/golden/canvas.npz by running the reference's own MoCoFlowTrainer.render
(trainer/trainer_moco_flow.py:226-265, imported from /root/reference in the build container) on canned network outputs.

The method is called unbound on a stand-in ``self`` whose ``forward`` returns slices of the canned per-ray results
(the chunk loop of :238-247 runs as written); only the image / depth assembly after it (:249-263) is under test.

Two accommodations, neither touching the code under test:
  * the trainer module imports packages this container does not have (imageio, mcubes, trimesh, knn_cuda, plyfile,
    kornia, tensorboardX); none is used by ``render`` -- they are replaced by empty stand-in modules for the import;
  * ``np.float`` (used at :255, removed in numpy 1.24) is restored as the alias of ``float`` it used to be.

Usage: python tests/golden/make_golden_canvas.py
"""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MOCO_REFERENCE", "/root/reference")
ABSENT = {"imageio", "mcubes", "trimesh", "knn_cuda", "plyfile", "kornia", "tensorboardX"}


class _StandIns(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in ABSENT:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__path__, m.__spec__, m.__name__ = [], spec, spec.name
        return m

    def exec_module(self, module):
        pass


class _Self:
    """What ``render`` reads from the trainer: config['model']['N_rand'], device, forward()."""

    def __init__(self, n_rand, results):
        self.config = {"model": {"N_rand": n_rand}}
        self.device = torch.device("cpu")
        self._results, self._pos = results, 0

    def forward(self, rays, background=None, use_nof=True, test_time=False):
        n = rays.shape[0]
        out = {k: v[self._pos:self._pos + n] for k, v in self._results.items()}
        self._pos += n
        return out


def main():
    if not hasattr(np, "float"):
        np.float = float
    sys.path.insert(0, REF)
    sys.meta_path.append(_StandIns())
    render = importlib.import_module("trainer.trainer_moco_flow").MoCoFlowTrainer.render
    rng = np.random.default_rng(11)
    out = {}
    for name, (P, typ, n_rand) in {"fine": (96, "fine", 32), "coarse": (61, "coarse", 1024)}.items():
        rays_msk = rng.random(P) < 0.6
        n = int(rays_msk.sum())
        rays = torch.from_numpy(rng.normal(size=(P, 9)).astype("float32"))
        background = torch.from_numpy(rng.uniform(0, 1, size=(P, 3)).astype("float32"))
        rgb = torch.from_numpy(rng.uniform(0, 1, size=(n, 3)).astype("float32"))
        depth = torch.from_numpy(rng.uniform(2, 4, size=(n,)).astype("float32"))
        opacity = rng.uniform(0, 1, size=(n,)).astype("float32")
        opacity[rng.random(n) < 0.3] = 0.0               # empty rays: the background shows, depth stays 8
        opacity = torch.from_numpy(opacity)
        canned = {f"rgb_{typ}": rgb, f"depth_{typ}": depth, f"opacity_{typ}": opacity}
        if typ == "fine":                                # a coarse set beside it must be ignored by the assembly
            canned.update(rgb_coarse=rgb * 0.5, depth_coarse=depth + 1, opacity_coarse=opacity * 0.5)
        fake = _Self(n_rand, canned)
        res = render(fake, rays, background, rays_msk=rays_msk, use_nof=True, test_time=True)
        assert fake._pos == n
        out.update({f"{name}_background": background.numpy(), f"{name}_rays_msk": rays_msk, f"{name}_rgb": rgb.numpy(),
                    f"{name}_depth": depth.numpy(), f"{name}_opacity": opacity.numpy(),
                    f"{name}_img": res[f"rgb_{typ}"].numpy(), f"{name}_depth_img": res[f"depth_{typ}"].numpy()})
    np.savez_compressed(os.path.join(HERE, "canvas.npz"), **out)
    print("wrote canvas.npz", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()

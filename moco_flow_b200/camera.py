"""Ray generation and canvas assembly on the device.

Mirrors the reference's host-side helpers on either side of ``render_rays``: ``utils/camera.py:29-82``
(``gen_ray_directions`` + ``gen_rays``), ``Camera.make_rays`` (``:134-148``) and the masked gather / canvas scatter of
``MoCoFlowTrainer.render`` (``trainer/trainer_moco_flow.py:226-268``), which the reference runs with numpy on the CPU
(including a device->host copy of the opacities).  Here a frame's rays are produced by one kernel from the 3x4 pose
and the finished image by two, so an inference frame needs 48 bytes of host->device traffic instead of 36 B/pixel.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L


def near_far_from_aabb(aabb_verts, c2w) -> Tuple[float, float]:
    """utils/camera.py:137-139: min / max distance from the camera centre to the 8 box corners (host, 8 points)."""
    d = np.sqrt(np.sum((np.asarray(aabb_verts) - np.asarray(c2w)[:3, 3]) ** 2, axis=-1))
    return float(d.min()), float(d.max())


def make_rays(H: int, W: int, focal: float, center: Sequence[float], c2w, near: float, far: float, idx: float,
              pixel_index: Optional[torch.Tensor] = None, device=None) -> torch.Tensor:
    """Rows ``[o(3) d(3) near far idx]`` (N, 9) for all H*W pixels or for ``pixel_index`` (int64, device)."""
    if pixel_index is not None:
        device = pixel_index.device
        if pixel_index.dtype != torch.int64 or not pixel_index.is_contiguous():
            raise ValueError("pixel_index must be a contiguous int64 tensor")
    device = torch.device(device if device is not None else "cuda")
    if device.type != "cuda":
        raise RuntimeError("moco_flow_b200.camera runs on CUDA only (no CPU fallback)")
    n = int(pixel_index.numel()) if pixel_index is not None else H * W
    rays = torch.empty(n, 9, device=device)
    pose = None
    if c2w is not None:
        pose = L.f32_array(np.asarray(c2w, dtype=np.float64)[:3, :4].reshape(-1), 12)
    with torch.cuda.device(device):
        L.check(L.lib().mcf_make_rays(C.c_int(H), C.c_int(W), C.c_float(focal), C.c_float(center[0]),
                                      C.c_float(center[1]), pose, C.c_float(near), C.c_float(far), C.c_float(idx),
                                      L.ptr(pixel_index), C.c_longlong(n), L.ptr(rays), C.c_int(9), L.stream_ptr()),
                "mcf_make_rays")
    return rays


def scatter_canvas(background: torch.Tensor, pixel_index: Optional[torch.Tensor], rgb: torch.Tensor,
                   depth: torch.Tensor, opacity: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """trainer/trainer_moco_flow.py:247-262 without the host round trip: returns ``(img (P,3), depth (P,))``."""
    for t in (background, rgb, depth, opacity):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError("scatter_canvas needs contiguous CUDA float32 tensors (no CPU fallback)")
    P = background.shape[0]
    n = rgb.shape[0]
    img = torch.empty(P, 3, device=background.device)
    dep = torch.empty(P, device=background.device)
    with torch.cuda.device(background.device):
        L.check(L.lib().mcf_canvas_scatter(L.ptr(background), C.c_longlong(P), L.ptr(pixel_index), C.c_longlong(n),
                                           L.ptr(rgb), L.ptr(depth), L.ptr(opacity), L.ptr(img), L.ptr(dep),
                                           L.stream_ptr()), "mcf_canvas_scatter")
    L.LAUNCHES += 1  # init + scatter
    return img, dep


class Camera:
    """utils/camera.py:98-148 (``size`` = (H, W), ``K`` 3x3 intrinsics, ``c2w`` set by the caller)."""

    def __init__(self, size, K, D=None, device=None):
        self.size = size
        self.K = np.asarray(K)
        self.D = D
        self.c2w = None
        self.device = device
        self.focal = [float(self.K[0][0]), float(self.K[1][1])]
        self.center = [float(self.K[0][2]), float(self.K[1][2])]

    def make_rays(self, aabb_verts, idx, pixel_index: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert self.c2w is not None, 'Camera is not initialized'
        near, far = near_far_from_aabb(aabb_verts, self.c2w)
        return make_rays(self.size[0], self.size[1], self.focal[0], self.center, np.asarray(self.c2w)[:3, :4], near, far,
                         float(idx), pixel_index, self.device)

"""SMPL correspondence sampling on the device.

``get_frame_correspondence`` (datasets/moco_flow_dataset.py:87-142) finds, for every query point, the nearest posed
SMPL vertex with ``knn_cuda.KNN(k=1)`` -- the only native CUDA dependency of the reference -- and maps the point with
that vertex's source->target transform.  ``nearest_vertex`` is that step as one brute-force kernel;
``split_correspondences`` reproduces the inside / outside split the trainer consumes.  The SMPL model itself (vertex
posing, per-vertex transforms) is data-gated and stays with the caller.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L


def nearest_vertex(verts: torch.Tensor, query: torch.Tensor, trans: Optional[torch.Tensor] = None,
                   thickness: float = 0.2):
    """verts (V,3), query (N,3), trans (V,4,4) or None -> dist (N,), ind (N,) int64, cano (N,3) or None, inside (N,) bool."""
    for t in (verts, query) + ((trans,) if trans is not None else ()):
        if not (t.is_cuda and t.dtype == torch.float32):
            raise RuntimeError("nearest_vertex needs CUDA float32 tensors (no CPU fallback)")
    verts, query = verts.contiguous(), query.contiguous()
    V, N = verts.shape[0], query.shape[0]
    if trans is not None:
        if tuple(trans.shape) != (V, 4, 4):
            raise ValueError("trans must be (V, 4, 4)")
        trans = trans.contiguous()
    dev = query.device
    dist = torch.empty(N, device=dev)
    ind = torch.empty(N, dtype=torch.int64, device=dev)
    cano = torch.empty(N, 3, device=dev) if trans is not None else None
    inside = torch.empty(N, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib().mcf_nearest_vertex(L.ptr(verts), C.c_int(V), L.ptr(trans), L.ptr(query), C.c_longlong(N),
                                           C.c_float(thickness), L.ptr(dist), L.ptr(ind), L.ptr(cano), L.ptr(inside),
                                           L.stream_ptr()), "mcf_nearest_vertex")
    return dist, ind, cano, inside.bool()


def split_correspondences(query: torch.Tensor, cano: torch.Tensor, inside: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """datasets/moco_flow_dataset.py:124-133: rows ``[query(3) | canonical(3)]`` of the points nearer than ``thickness``
    to the surface, and of the others."""
    both = torch.cat([query.view(-1, 3), cano.view(-1, 3)], dim=-1)
    return both[inside], both[~inside]

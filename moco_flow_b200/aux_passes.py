"""Auxiliary point-batch passes that reuse the fused MLP kernels.

``density_grid`` is the occupancy pass of ``MoCoFlowTrainer.visualize_mesh`` (trainer/trainer_moco_flow.py:484-531):
sigma of the (fine) NeRF on an N^3 lattice over [-1.5, 1.5]^3, optionally warped first by the backward flow network of
one frame, clamped at 0 and reshaped to (N, N, N) -- the volume the reference hands to marching cubes (``mcubes``, a CPU
library that stays with the caller).  The reference embeds every chunk, zero-pads it and calls the modules through their
dense-input interface; here the points go straight into the fused chains (encoder inside the first layer's operand
builder, sigma head on CUDA cores), chunk after chunk, nothing but the (N^3,) result ever exists in HBM.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch


def lattice_points(n_grid: int, device, lo: float = -1.5, hi: float = 1.5) -> torch.Tensor:
    """``np.stack(np.meshgrid(x, y, z), -1).reshape(-1, 3)`` of trainer/trainer_moco_flow.py:494-497 (numpy's default
    'xy' indexing: the FIRST output axis runs over y), float64 linspace cast to float32."""
    t = torch.linspace(lo, hi, n_grid, dtype=torch.float64, device=device)
    x, y, z = torch.meshgrid(t, t, t, indexing="xy")
    return torch.stack([x, y, z], -1).reshape(-1, 3).float()


@torch.no_grad()
def density_grid(nerf_model, nerf_embedding_xyz, n_grid: int = 256, frame_idx: int = -1, nof_model=None,
                 nof_embeddings: Optional[Sequence] = None, num_frames: Optional[int] = None,
                 chunk: int = 1 << 20) -> torch.Tensor:
    """(N, N, N) tensor ``max(sigma, 0)`` on the device of ``nerf_model`` (trainer/trainer_moco_flow.py:484-516).

    ``frame_idx != -1`` warps the lattice with ``nof_model`` (the backward flow network) conditioned on that frame,
    exactly as ``forward_nof(xyz, tensor([frame_idx]), 'bw_NoF')`` does (:507-508, :160-187: the index is mapped to
    ``idx * 2 / num_frames - 1`` before it is embedded)."""
    dev = next(nerf_model.parameters()).device
    pts = lattice_points(n_grid, dev)
    out = torch.empty(pts.shape[0], device=dev)
    ind_feat = None
    if frame_idx != -1:
        if nof_model is None or nof_embeddings is None or not num_frames:
            raise ValueError("frame_idx != -1 needs nof_model, nof_embeddings=[xyz, ind] and num_frames")
        ind = torch.full((1, 1), float(frame_idx) * 2.0 / float(num_frames) - 1.0, device=dev)
        ind_feat = nof_embeddings[1](ind)                      # one "ray" = the whole chunk shares the frame index
    for b in range(0, pts.shape[0], chunk):
        x = pts[b:b + chunk]
        n = x.shape[0]
        if ind_feat is not None:
            x = nof_model.evaluate(xyz=x, pe=nof_embeddings[0], ray_feat=ind_feat, rows_per_ray=n)
        sig = nerf_model.evaluate(xyz=x, pe=nerf_embedding_xyz, ray_feat=None, rows_per_ray=n, sigma_only=True)
        out[b:b + n] = sig.view(-1)
    return out.clamp_min_(0.0).view(n_grid, n_grid, n_grid)

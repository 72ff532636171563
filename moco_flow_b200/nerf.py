"""Canonical NeRF MLP with the reference's interface (models/nerf.py:5-102), executed as one fused
tcgen05 chain kernel.

Parameters keep the reference's ``state_dict`` names and shapes (``xyz_encoding_{i}.0.*``,
``xyz_encoding_final.*``, ``extra_encoding.0.*``, ``sigma.*``, ``rgb.0.*``) so reference checkpoints
load unchanged; the ``nn.Sequential`` containers exist only to carry those names.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _lib as L
from . import ops
from . import plans as P
from .mlp import FusedMLP, fold_bias, setup_input

_EXTRA_TYPES = ("none", "ind", "dir", "latent_code")


class NeRF(FusedMLP):
    def __init__(self, D=8, W=256, in_channels_xyz=33, skips=[4], extra_feat_type="none", extra_feat_dim=0):
        super().__init__()
        assert extra_feat_type in _EXTRA_TYPES, f"extra_feat_type {extra_feat_type} for NeRF model not supported!!!"
        self.D, self.W, self.in_channels_xyz, self.skips = D, W, in_channels_xyz, skips
        self.extra_feat_type, self.extra_feat_dim = extra_feat_type, extra_feat_dim
        for i in range(D):
            fan_in = in_channels_xyz if i == 0 else W + (in_channels_xyz if i in skips else 0)
            setattr(self, f"xyz_encoding_{i+1}", nn.Sequential(nn.Linear(fan_in, W), nn.ReLU(True)))
        self.xyz_encoding_final = nn.Linear(W, W)
        if extra_feat_type == "latent_code":
            self.app_code = torch.randn(1000, extra_feat_dim, requires_grad=True)
        tail_in = W + (extra_feat_dim if extra_feat_type != "none" else 0)
        self.extra_encoding = nn.Sequential(nn.Linear(tail_in, W // 2), nn.ReLU(True))
        self.sigma = nn.Linear(W, 1)
        self.rgb = nn.Sequential(nn.Linear(W // 2, 3), nn.Sigmoid())

    # -- shape helpers ------------------------------------------------------------------------
    def _extra_dim(self) -> int:
        return self.extra_feat_dim if self.extra_feat_type != "none" else 0

    def _plan(self, sigma_only: bool, training: bool) -> ops.PackedPlan:
        return self._packed(("fwd", sigma_only, training),
                            lambda: P.nerf_forward_plan(self.D, self.W, self.in_channels_xyz, tuple(self.skips),
                                                        self._extra_dim(), sigma_only, training))

    # -- fused evaluation ----------------------------------------------------------------------
    def evaluate(self, *, xyz: Optional[torch.Tensor] = None, pe=None, dense: Optional[torch.Tensor] = None,
                 ray_feat: Optional[torch.Tensor] = None, rows_per_ray: int = 1, sigma_only: bool = False):
        """rgb-sigma (M,4) or sigma (M,1) for M points.

        Either ``xyz`` (M,3) + the xyz ``Embedding`` (encoding fused into layer 1), or ``dense`` (M, >=in_channels_xyz)
        already-embedded rows.  ``ray_feat`` (M/rows_per_ray, E') holds the per-ray extra feature (index or direction
        embedding, E' <= extra_feat_dim; narrower means zero-padded as in models/rendering.py:135-136,140-141).
        """
        from .autograd_mlp import nerf_apply  # late import: backward machinery
        return nerf_apply(self, xyz, pe, dense, ray_feat, rows_per_ray, sigma_only)

    def forward(self, inputs, sigma_only=False, img_ind=None):
        """inputs: (B, in_channels_xyz [+ extra_feat_dim]) embedded rows, as in the reference."""
        if self.extra_feat_type == "latent_code" and not sigma_only:
            raise NotImplementedError("NeRF model does not support latent code yet!!!")
        ops._need_cuda(inputs)
        if inputs.stride(-1) != 1:
            inputs = inputs.contiguous()
        cx, E = self.in_channels_xyz, self._extra_dim()
        feat = None if (sigma_only or E == 0) else inputs[:, cx:cx + E]
        return self.evaluate(dense=inputs, ray_feat=feat, rows_per_ray=1, sigma_only=sigma_only)

"""Neural optical-flow MLP ("NoF") with the reference's interface (models/nof.py:6-85), executed as
one fused tcgen05 chain kernel whose last epilogue applies the log-quaternion / pivot / translation
transform (kornia 0.6.5 semantics of quaternion_log_to_exp + quaternion_to_rotation_matrix) in fp32.

``state_dict`` names follow the reference: ``nof_encoding_{i}.0.*`` and ``nof_encoding_final.*``.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import ops
from . import plans as P
from .mlp import FusedMLP


class NoF(FusedMLP):
    def __init__(self, D=8, W=256, in_channels_xyz=33, skips=[4], extra_feat_type="ind", extra_feat_dim=0,
                 use_quat=False):
        super().__init__()
        assert extra_feat_type in ["ind", "latent_code"], \
            f"extra_feat_type {extra_feat_type} for NoF model not supported!!!"
        self.D, self.W, self.in_channels_xyz, self.skips = D, W, in_channels_xyz, skips
        self.use_quat, self.extra_feat_type, self.extra_feat_dim = use_quat, extra_feat_type, extra_feat_dim
        if extra_feat_type == "latent_code":
            self.time_code = torch.randn(1000, extra_feat_dim, requires_grad=True)
        width_in = in_channels_xyz + extra_feat_dim
        for i in range(D):
            fan_in = width_in if i == 0 else W + (width_in if i in skips else 0)
            setattr(self, f"nof_encoding_{i+1}", nn.Sequential(nn.Linear(fan_in, W), nn.ReLU(True)))
        # 3 (log quaternion) + 3 (pivot) + 3 (translation), or a 3-vector residual flow
        self.nof_encoding_final = nn.Linear(W, 9 if use_quat else 3)

    def _plan(self, training: bool) -> ops.PackedPlan:
        return self._packed(("fwd", training),
                            lambda: P.nof_forward_plan(self.D, self.W, self.in_channels_xyz, tuple(self.skips),
                                                       self.extra_feat_dim, bool(self.use_quat), training))

    def evaluate(self, *, xyz: torch.Tensor, pe=None, dense: Optional[torch.Tensor] = None,
                 ray_feat: Optional[torch.Tensor] = None, rows_per_ray: int = 1):
        """Warped positions (M,3).  ``xyz`` (M,3) are the points being warped; they are encoded in-kernel with ``pe``
        unless ``dense`` (M, >= in_channels_xyz) carries already-embedded rows.  ``ray_feat``: (M/rows_per_ray, E)."""
        from .autograd_mlp import nof_apply
        return nof_apply(self, xyz, pe, dense, ray_feat, rows_per_ray)

    def forward(self, inputs, xyz, img_ind=None):
        """inputs: (N, in_channels_xyz + extra_feat_dim) embedded rows; xyz: (N,3) -> (N,3)."""
        if self.extra_feat_type == "latent_code":
            raise NotImplementedError("NoF model does not support latent code yet!!!")
        ops._need_cuda(inputs, xyz)
        if inputs.stride(-1) != 1:
            inputs = inputs.contiguous()
        cx, E = self.in_channels_xyz, self.extra_feat_dim
        return self.evaluate(xyz=xyz, dense=inputs, ray_feat=inputs[:, cx:cx + E] if E > 0 else None, rows_per_ray=1)

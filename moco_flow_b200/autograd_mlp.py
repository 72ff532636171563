"""Forward launches (and autograd wiring) of the fused NeRF / NoF chains."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L
from . import ops
from . import plans as P
from .mlp import fold_bias, setup_input


def _rows(xyz, dense) -> int:
    return int(dense.shape[0] if dense is not None else xyz.shape[0])


def _wants_grad(model, *tensors) -> bool:
    if not torch.is_grad_enabled():
        return False
    if any(t is not None and t.requires_grad for t in tensors):
        return True
    return any(p.requires_grad for p in model.parameters())


# ------------------------------------------------------------------------------------------------
# NeRF
# ------------------------------------------------------------------------------------------------
def nerf_forward_launch(model, xyz, pe, dense, ray_feat, rows_per_ray: int, sigma_only: bool, training: bool):
    M = _rows(xyz, dense)
    if M % rows_per_ray != 0:
        raise ValueError("row count must be a multiple of rows_per_ray")
    R = M // rows_per_ray
    dev = (dense if dense is not None else xyz).device
    pp = model._plan(sigma_only, training)
    cp = ops.chain_params(pp, M, rows_per_ray, R)
    keep = setup_input(cp, None if xyz is None else xyz.detach().contiguous(), pe,
                       None if dense is None else dense.detach(), model.in_channels_xyz)
    E = model._extra_dim()
    if not sigma_only and E > 0:
        if ray_feat is None or ray_feat.shape[0] != R:
            raise ValueError("ray_feat with one row per ray is required")
        lin = model.extra_encoding[0]
        rb = fold_bias(lin.weight, model.W, lin.bias, ray_feat[:, :E])
        cp.raybias[0] = rb.data_ptr()
        keep.append(rb)
    out = torch.empty(M, 1 if sigma_only else 4, device=dev)
    cp.out, cp.out_stride, cp.sigma_col = out.data_ptr(), out.shape[1], (0 if sigma_only else 3)
    flops = ops.linear_flops(model)
    if sigma_only:
        flops -= 2.0 * (model.W * model.W + (model.W + E) * (model.W // 2) + (model.W // 2) * 3)
    ops.launch_chain(cp, "nerf_fwd", flops)
    return out, keep


def nerf_apply(model, xyz, pe, dense, ray_feat, rows_per_ray, sigma_only):
    if _wants_grad(model, xyz, dense):
        from .backward_mlp import nerf_autograd
        return nerf_autograd(model, xyz, pe, dense, ray_feat, rows_per_ray, sigma_only)
    out, _ = nerf_forward_launch(model, xyz, pe, dense, ray_feat, rows_per_ray, sigma_only, training=False)
    return out


# ------------------------------------------------------------------------------------------------
# NoF
# ------------------------------------------------------------------------------------------------
def nof_forward_launch(model, xyz, pe, dense, ray_feat, rows_per_ray: int, training: bool):
    M = _rows(xyz, dense)
    if M % rows_per_ray != 0:
        raise ValueError("row count must be a multiple of rows_per_ray")
    R = M // rows_per_ray
    dev = xyz.device
    pp = model._plan(training)
    cp = ops.chain_params(pp, M, rows_per_ray, R)
    keep = setup_input(cp, xyz.detach().contiguous(), pe, None if dense is None else dense.detach(),
                       model.in_channels_xyz)
    E, cx = model.extra_feat_dim, model.in_channels_xyz
    if E > 0:
        check_nof_ray_feat(model, ray_feat, R)
        for k, i in enumerate(P.folded_layers(model.D, tuple(model.skips))):
            lin = getattr(model, f"nof_encoding_{i+1}")[0]
            rb = fold_bias(lin.weight, cx, lin.bias, ray_feat[:, :E])
            cp.raybias[k] = rb.data_ptr()
            keep.append(rb)
    out = torch.empty(M, 3, device=dev)
    cp.out, cp.out_stride, cp.use_quat = out.data_ptr(), 3, int(model.use_quat)
    head_save = None
    if training:
        head_save = torch.empty(M, 12, device=dev)
        cp.head_save = head_save.data_ptr()
    ops.launch_chain(cp, "nof_fwd", ops.linear_flops(model))
    return out, keep, head_save


def check_nof_ray_feat(model, ray_feat, n_rays: int) -> None:
    """The reference concatenates the index embedding without padding (models/rendering.py:71-74), so a width other than
    ``extra_feat_dim`` fails in its first Linear; a silent zero-pad / truncation here would hide that."""
    if ray_feat is None or ray_feat.shape[0] != n_rays:
        raise ValueError("ray_feat with one row per ray is required")
    if ray_feat.shape[1] != model.extra_feat_dim:
        raise RuntimeError(f"NoF expects a {model.extra_feat_dim}-wide per-ray feature, got {ray_feat.shape[1]} "
                           "(mat1 and mat2 shapes cannot be multiplied in the reference)")


def nof_apply(model, xyz, pe, dense, ray_feat, rows_per_ray):
    if _wants_grad(model, xyz, dense):
        from .backward_mlp import nof_autograd
        return nof_autograd(model, xyz, pe, dense, ray_feat, rows_per_ray)
    out, _, _ = nof_forward_launch(model, xyz, pe, dense, ray_feat, rows_per_ray, training=False)
    return out

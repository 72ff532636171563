"""Launch logic shared by the NeRF and NoF modules: plan cache, weight packing, per-ray bias folding
and the chain-kernel launch.  See plans.py for the layer programs and csrc/chain.cu for the kernel.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import _lib as L
from . import ops
from . import plans as P


class FusedMLP(nn.Module):
    """Base class: owns the per-(mode) packed plans of one module."""

    def _plan_cache(self) -> Dict[tuple, ops.PackedPlan]:
        cache = self.__dict__.get("_mcf_plans")
        if cache is None:
            cache = {}
            self.__dict__["_mcf_plans"] = cache
        return cache

    def _param_dict(self) -> Dict[str, torch.Tensor]:
        return {k: v for k, v in self.named_parameters()}

    def _packed(self, key: tuple, build) -> ops.PackedPlan:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("moco_flow_b200 modules run on CUDA only (no CPU fallback); call .cuda() first")
        cache = self._plan_cache()
        full_key = key + (dev.index,)
        if full_key not in cache:
            cache[full_key] = ops.PackedPlan(build(), dev)
        pp = cache[full_key]
        pp.repack(self._param_dict())
        return pp

    def _apply(self, fn, *a, **k):  # parameters moved/cast: drop packed copies
        self.__dict__.pop("_mcf_plans", None)
        return super()._apply(fn, *a, **k)


def feat_view(feat: torch.Tensor) -> torch.Tensor:
    if feat.stride(-1) != 1:
        feat = feat.contiguous()
    return feat


def fold_bias(weight: torch.Tensor, col_off: int, bias: torch.Tensor, feat: torch.Tensor) -> torch.Tensor:
    """Per-row bias b + W[:, col_off:col_off+E] feat^T for per-ray constant input columns."""
    import ctypes as C
    feat = feat_view(feat.detach())
    weight, bias = weight.detach(), bias.detach()
    R, E = feat.shape
    N = weight.shape[0]
    out = torch.empty(R, N, device=feat.device)
    L.check(L.lib().mcf_ray_bias(L.ptr(weight), C.c_int(weight.stride(0)), C.c_int(col_off), L.ptr(bias),
                                 L.ptr(feat), C.c_int(feat.stride(0)), C.c_int(E), C.c_int(R), C.c_int(N),
                                 L.ptr(out), L.stream_ptr()), "mcf_ray_bias")
    return out


def setup_input(cp: L.ChainParams, xyz: Optional[torch.Tensor], pe, dense: Optional[torch.Tensor], cx: int) -> list:
    """Fills the prologue fields; returns tensors that must stay alive until the launch."""
    keep = []
    if dense is not None:
        cp.prologue = L.PRO_DENSE
        cp.dense, cp.dense_stride, cp.dense_cols = dense.data_ptr(), dense.stride(0), cx
        keep.append(dense)
    else:
        if pe.out_channels > cx:
            raise RuntimeError(f"xyz embedding has {pe.out_channels} channels but the model expects at most {cx}")
        if pe.in_channels != 3:
            raise ValueError("the fused xyz encoder expects 3 input channels")
        cp.prologue = L.PRO_PE_XYZ
        ops.set_pe(cp, pe.frequencies(), pe.multipliers(), cx)
    if xyz is not None:
        cp.xyz = xyz.data_ptr()
        keep.append(xyz)
    return keep

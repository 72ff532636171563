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


# Per-call memo (rendering.render_rays opens it): within one render call the per-ray embeddings and the folded per-ray
# biases depend only on (rays, weights), not on the sample positions, so the coarse and the fine pass and the repeated
# evaluations of one flow network share them -- 6 distinct results instead of 22 launches in a training step.
_MEMO = None


class call_memo:
    """Context manager: memoise per-ray features / folded biases for the duration of one rendering call."""

    def __enter__(self):
        global _MEMO
        self.outer = _MEMO
        if _MEMO is None:
            _MEMO = {}
        return self

    def __exit__(self, *exc):
        global _MEMO
        if self.outer is None:
            _MEMO = None
        return False


def memoised(key: tuple, keep_alive: tuple, make):
    """``make()`` once per key while a ``call_memo`` is open; ``keep_alive`` pins the tensors whose addresses are in the key."""
    if _MEMO is None:
        return make()
    hit = _MEMO.get(key)
    if hit is None:
        hit = (make(), keep_alive)
        _MEMO[key] = hit
    return hit[0]


def _tkey(t: torch.Tensor) -> tuple:
    return (t.data_ptr(), t._version, tuple(t.shape), t.stride())


def fold_bias(weight: torch.Tensor, col_off: int, bias: torch.Tensor, feat: torch.Tensor) -> torch.Tensor:
    """Per-row bias b + W[:, col_off:col_off+E] feat^T for per-ray constant input columns."""
    feat = feat_view(feat.detach())
    weight, bias = weight.detach(), bias.detach()
    key = ("fold", _tkey(weight), col_off, _tkey(bias), _tkey(feat))
    return memoised(key, (weight, bias, feat), lambda: _fold_bias(weight, col_off, bias, feat))


def _fold_bias(weight: torch.Tensor, col_off: int, bias: torch.Tensor, feat: torch.Tensor) -> torch.Tensor:
    import ctypes as C
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
        ops.set_pe(cp, pe, cx, (xyz if xyz is not None else dense).device)
    if xyz is not None:
        cp.xyz = xyz.data_ptr()
        keep.append(xyz)
    return keep

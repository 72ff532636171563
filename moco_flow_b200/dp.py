"""Ray-sharded data parallelism: the only place the path partitions.

One process per GPU.  Every rank renders its own contiguous slice of the step's ray batch with
replicated weights; parameter gradients live in ONE flat fp32 buffer (the ``.grad`` of every
parameter is a view into it), so the exchange is a single NCCL all-reduce (sum, then 1/world) over
NVLink -- 1 321 498 floats = 5.29 MB for the four c2f networks.  The reference wraps its nets in
DistributedDataParallel (trainer/base.py:251-256) but never routes the forward through the wrapper
(SURVEY 2.2); this implements the intended semantics instead of that accident.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def _params_rewritten() -> None:
    """Writes through ``p.data`` advance no version counter: tell the modules to re-pack their bf16 weight images."""
    from . import ops
    ops.bump_param_epoch()


def enable_global_residual_means(group=None) -> None:
    """From now on the fused flow-residual means of ``render_rays(fused_residual_mean=True)`` are taken over the rays of
    ALL ranks: the (masked sum, count) pairs of a call are summed with one small all-reduce before the division, and
    the backward scales by world/count_global, so that the rank-averaged gradients equal the gradients of the
    single-process loss on the concatenated batch (models/rendering.py:306-314,365-373 +
    trainer/trainer_moco_flow.py:319-327 define that global mean).  No-op without an initialised process group."""
    from . import ops
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        ops.RESIDUAL_DP = (group, dist.get_world_size(group))
    else:
        ops.RESIDUAL_DP = None


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of ``n_items`` for ``rank`` (first ranks get the remainder)."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_rays(rays: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    b, e = shard_bounds(rays.shape[0], rank, world)
    return rays[b:e]


class FlatGradients:
    """Flat gradient buffer shared by a list of modules; ``p.grad`` are views, so autograd accumulates in place."""

    def __init__(self, modules: Iterable[torch.nn.Module], fused_accumulate: bool = False,
                 flatten_params: bool = False):
        """``fused_accumulate``: let the fused MLP backward add its weight gradients straight into these ``.grad``
        views (moco_flow_b200.backward_mlp.ACCUMULATE_INTO_GRAD) instead of going through autograd's per-parameter
        accumulation kernels.  Process-wide switch; use only with ``loss.backward()``.

        ``flatten_params``: also move the parameters themselves into one flat buffer (``p.data`` become views, in the
        same order as the gradients), so that ``optim.FusedAdam`` updates all of them with one launch and
        ``broadcast_parameters`` is a single collective.  ``state_dict`` names and shapes are unchanged."""
        if fused_accumulate:
            from . import backward_mlp
            backward_mlp.ACCUMULATE_INTO_GRAD = True
        self.params: List[torch.nn.Parameter] = []
        seen = set()
        for m in modules:
            for p in m.parameters():
                if id(p) not in seen:
                    seen.add(id(p))
                    self.params.append(p)
        if not self.params:
            raise ValueError("no parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        self.buffer = torch.zeros(self.numel, device=dev, dtype=dt)
        self.param_buffer = None
        if flatten_params:
            self.param_buffer = torch.empty(self.numel, device=dev, dtype=dt)
        off = 0
        for p in self.params:
            if flatten_params:
                view = self.param_buffer[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
            p.grad = self.buffer[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self) -> None:
        self.buffer.zero_()

    def broadcast_parameters(self, src: int = 0, group=None) -> None:
        """Rank ``src``'s weights to everyone: one collective when the parameters are flattened."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        if self.param_buffer is not None:
            dist.broadcast(self.param_buffer, src=src, group=group)
        else:
            for p in self.params:
                dist.broadcast(p.data, src=src, group=group)
        _params_rewritten()

    def check_views(self) -> None:
        """Every ``.grad`` must still alias the flat buffer (an ``optimizer.zero_grad(set_to_none=True)`` or a
        ``p.grad = None`` in between would silently exclude that parameter from the all-reduce).  Also the point where
        the weight-gradient side stream (backward_mlp.DW_OVERLAP_SMS) is joined: the buffer is complete after this."""
        if self.buffer.is_cuda:
            from . import backward_mlp
            backward_mlp.join_weight_gradients()
        off = 0
        base = self.buffer.data_ptr()
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != base + 4 * off:
                raise RuntimeError("a parameter's .grad no longer aliases the FlatGradients buffer (zero_grad with "
                                   "set_to_none=True?); use FlatGradients.zero() or FusedAdam.zero_grad()")
            off += p.numel()

    def allreduce_sum(self, group=None) -> float:
        """Sum over ranks; returns the factor (1/world) the caller still has to apply -- ``FusedAdam.step(grad_scale=)``
        folds it into the update instead of spending a pass over the buffer."""
        self.check_views()
        if not (dist.is_available() and dist.is_initialized()):
            return 1.0
        world = dist.get_world_size(group)
        if world == 1:
            return 1.0
        dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM, group=group)
        return 1.0 / world

    def allreduce_mean(self, group=None) -> None:
        """Sum over ranks then divide by the world size (no-op without an initialised process group)."""
        self.check_views()
        if not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size(group)
        if world == 1:
            return
        dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM, group=group)
        self.buffer.mul_(1.0 / world)


def broadcast_parameters(modules: Iterable[torch.nn.Module], src: int = 0, group=None) -> None:
    """Rank ``src``'s weights to everyone (what DDP's constructor does, trainer/base.py:254-256)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for m in modules:
        for p in m.parameters():
            dist.broadcast(p.data, src=src, group=group)
    _params_rewritten()

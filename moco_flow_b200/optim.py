"""Fused Adam for the training step that follows the ray-rendering backward.

``FusedAdam`` is a ``torch.optim.Optimizer`` with the constructor and ``param_groups`` of ``torch.optim.Adam``
as the reference builds it (trainer/base.py:122-133: ``Adam(parameters, lr=..., eps=1e-8, weight_decay=...)``), so the
reference's ``MultiStepLR`` / ``ExponentialLR`` / ... schedulers (trainer/base.py:141-160) and its two-optimizer
arrangement over overlapping parameter sets (trainer/trainer_moco_flow.py:121-139, stepped one after the other at
trainer/base.py:188-197) work unchanged.  Parameters whose storage and ``.grad`` storage are adjacent -- what
``dp.FlatGradients(..., flatten_params=True)`` produces -- are merged into segments and each segment is one
``mcf_adam_step`` launch (csrc/optim.cu); the step count and learning rate are device scalars, so a captured CUDA
graph of the step stays valid while a scheduler keeps changing ``group['lr']`` (call ``sync_lr()`` outside the graph).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import torch

from . import _lib as L
from . import ops


def merge_segments(items: List[Tuple[int, int, int]]) -> List[Tuple[int, int, int, List[int]]]:
    """``items``: (param_ptr, grad_ptr, numel) per parameter.  Returns maximal runs that are contiguous in both the
    parameter and the gradient address space as (param_ptr, grad_ptr, numel, member indices), in address order."""
    order = sorted(range(len(items)), key=lambda i: items[i][0])
    segs: List[Tuple[int, int, int, List[int]]] = []
    for i in order:
        p, g, n = items[i]
        if n == 0:
            continue
        if segs:
            sp, sg, sn, members = segs[-1]
            if sp + 4 * sn == p and sg + 4 * sn == g:
                segs[-1] = (sp, sg, sn + n, members + [i])
                continue
        segs.append((p, g, n, [i]))
    return segs


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._plans: Dict[int, dict] = {}

    # ---- segment plan of one param group (rebuilt when storage pointers change) ----
    def _plan(self, gi: int, group: dict) -> dict:
        ps = [p for p in group["params"] if p.grad is not None]
        key = tuple((p.data_ptr(), p.grad.data_ptr(), p.numel()) for p in ps)
        plan = self._plans.get(gi)
        if plan is not None and plan["key"] == key:
            return plan
        for p in ps:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()
                    and p.grad.dtype == torch.float32):
                raise RuntimeError("FusedAdam needs contiguous CUDA float32 parameters and gradients (no CPU fallback)")
        dev = ps[0].device if ps else torch.device("cuda")
        old = plan
        segs = merge_segments(list(key))
        total = sum(s[2] for s in segs)
        m = torch.zeros(total, device=dev)
        v = torch.zeros(total, device=dev)
        step = old["step"] if old is not None else torch.zeros(1, dtype=torch.int64, device=dev)
        if old is None:   # a loaded state_dict (torch.optim.Adam layout) carries the step count per parameter
            loaded = [int(float(self.state[p]["step"])) for p in ps if "step" in self.state[p]]
            if loaded:
                step.fill_(max(loaded))
        lr_dev = old["lr_dev"] if old is not None else torch.full((1,), float(group["lr"]), device=dev)
        off = 0
        seg_off = []
        for sp, sg, sn, members in segs:
            o = off
            for i in members:
                p = ps[i]
                st = self.state[p]
                for name, buf in (("exp_avg", m), ("exp_avg_sq", v)):
                    view = buf[o:o + p.numel()].view_as(p)
                    if name in st:            # keep moments across a re-plan / load_state_dict
                        view.copy_(st[name])
                    st[name] = view
                st["step"] = step
                o += p.numel()
            seg_off.append(off)
            off += sn
        plan = dict(key=key, segs=segs, seg_off=seg_off, m=m, v=v, step=step, lr_dev=lr_dev, lr_host=None)
        self._plans[gi] = plan
        return plan

    def state_dict(self):
        """Same layout as ``torch.optim.Adam.state_dict()``; the moments are cloned so that a saved file holds one
        tensor per parameter instead of a view of the flat buffer (which ``torch.save`` would store whole, per view)."""
        sd = super().state_dict()
        for st in sd["state"].values():
            for k, v in list(st.items()):
                if torch.is_tensor(v):
                    st[k] = v.detach().clone()
        return sd

    def load_state_dict(self, state_dict) -> None:
        """Accepts the state of a ``torch.optim.Adam`` over the same parameters (trainer/base.py:316-321) or of a
        ``FusedAdam``; the flat moment buffers are rebuilt from it on the next step."""
        super().load_state_dict(state_dict)
        self._plans.clear()

    def zero_grad(self, set_to_none: bool = False) -> None:
        """Zeroes the gradients IN PLACE by default (torch's default drops them).  The reference calls
        ``optimizer.zero_grad()`` every step (trainer/base.py:188-190); with ``dp.FlatGradients`` every ``.grad`` is a
        view of the flat all-reduce buffer, and dropping the views would make the backward allocate fresh tensors that
        the all-reduce never sees.  ``set_to_none=True`` is refused for such views."""
        if set_to_none:
            for group in self.param_groups:
                for p in group["params"]:
                    if p.grad is not None and p.grad._base is not None:
                        raise RuntimeError("FusedAdam.zero_grad(set_to_none=True) would detach gradients that are views "
                                           "of a flat buffer (dp.FlatGradients); use zero_grad() / FlatGradients.zero()")
        super().zero_grad(set_to_none=set_to_none)

    def sync_lr(self) -> None:
        """Copies every group's ``lr`` to its device scalar (call after a scheduler step, outside graph capture)."""
        for gi, group in enumerate(self.param_groups):
            plan = self._plans.get(gi)
            if plan is not None and plan["lr_host"] != float(group["lr"]):
                plan["lr_dev"].fill_(float(group["lr"]))
                plan["lr_host"] = float(group["lr"])

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        """One Adam update of every group.  ``grad_scale`` multiplies the gradients first (1/world after a summing
        all-reduce)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        from . import backward_mlp
        backward_mlp.join_weight_gradients()   # gradients accumulated on the weight-gradient side stream
        for gi, group in enumerate(self.param_groups):
            plan = self._plan(gi, group)
            if not capturing and plan["lr_host"] != float(group["lr"]):
                plan["lr_dev"].fill_(float(group["lr"]))
                plan["lr_host"] = float(group["lr"])
            b1, b2 = group["betas"]
            n_seg = len(plan["segs"])
            for k, ((sp, sg, sn, _), so) in enumerate(zip(plan["segs"], plan["seg_off"])):
                L.check(L.lib().mcf_adam_step(
                    C.c_void_p(sp), C.c_void_p(sg), C.c_void_p(plan["m"].data_ptr() + 4 * so),
                    C.c_void_p(plan["v"].data_ptr() + 4 * so), C.c_longlong(sn), L.ptr(plan["lr_dev"]),
                    L.ptr(plan["step"]), C.c_double(b1), C.c_double(b2), C.c_float(group["eps"]),
                    C.c_float(group["weight_decay"]), C.c_float(grad_scale), C.c_int(1 if k == n_seg - 1 else 0),
                    L.stream_ptr()), "mcf_adam_step")
                if k == n_seg - 1:
                    L.LAUNCHES += 1   # the step-counter tick kernel
        ops.bump_param_epoch()   # raw-pointer writes: tell the modules to re-pack their bf16 weight images
        return loss


def get_optimizer(optimizer_config: dict, parameters):
    """trainer/base.py:122-139: ``{'type': 'adam'|'sgd', 'lr': ..., 'weight_decay': ...[, 'momentum': ...]}``.
    'adam' is the fused kernel above (eps 1e-8 as in the reference); 'sgd' is torch's; the reference's own RAdam /
    Ranger classes (utils/optimizers.py) are outside the hot path and not rebuilt."""
    kind = optimizer_config['type']
    if kind == 'adam':
        return FusedAdam(parameters, lr=optimizer_config['lr'], eps=1e-8, weight_decay=optimizer_config['weight_decay'])
    if kind == 'sgd':
        return torch.optim.SGD(parameters, lr=optimizer_config['lr'], momentum=optimizer_config['momentum'],
                               weight_decay=optimizer_config['weight_decay'])
    raise NotImplementedError(f"Optimizer type {kind} not implemented yet !!!")


def get_scheduler(scheduler_config: dict, optimizer, world_size: int = 1):
    """trainer/base.py:141-160, including the division of the step milestones by the world size (:145-146)."""
    from torch.optim import lr_scheduler
    kind = scheduler_config['type']
    if kind == 'steplr':
        return lr_scheduler.MultiStepLR(optimizer,
                                        milestones=[int(step // world_size) for step in scheduler_config['decay_step']],
                                        gamma=scheduler_config['decay_gamma'])
    if kind == 'explr':
        return lr_scheduler.ExponentialLR(optimizer, scheduler_config['lr_decay'])
    if kind == 'cosine':
        return lr_scheduler.CosineAnnealingLR(optimizer, T_max=scheduler_config['num_epochs'], eta_min=1e-8)
    if kind == 'poly':
        return lr_scheduler.LambdaLR(
            optimizer, lambda epoch: (1 - epoch / scheduler_config['num_epochs']) ** scheduler_config['poly_exp'])
    raise NotImplementedError('Scheduler type {} not implemented yet !!!'.format(kind))

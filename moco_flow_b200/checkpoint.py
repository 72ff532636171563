"""Checkpoint import / export in the reference's layout.

``trainer/base.py:279-327`` writes one ``.pth`` holding ``{'clock': ..., '<net>_net': state_dict, '<name>_optimizer':
state_dict, '<name>_scheduler': state_dict}`` and restores it with ``strict=False``; ``trainer/trainer_moco_flow.py:46-70``
loads single networks out of such a file (for a NeRF only the trunk and the density head: keys containing ``xyz`` or
``sigma``).  The modules of this package keep the reference's parameter names and shapes, so these helpers only restate
that file layout; a checkpoint trained with the reference drives this renderer and vice versa.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch


def _unwrap(net):
    return net.module if isinstance(net, torch.nn.parallel.DistributedDataParallel) else net


def save_ckpt(path: str, nets: Dict[str, torch.nn.Module], optimizers: Optional[Dict[str, object]] = None,
              schedulers: Optional[Dict[str, object]] = None, clock: Optional[dict] = None) -> str:
    """trainer/base.py:279-299.  ``path`` gets ``.pth`` appended when missing; returns the file name written."""
    save_path = path if str(path).endswith('.pth') else f'{path}.pth'
    save_dict = {'clock': clock if clock is not None else {}}
    for key, net in nets.items():
        save_dict[key + '_net'] = {k: v.detach().clone() for k, v in _unwrap(net).state_dict().items()}
    for key, opt in (optimizers or {}).items():
        save_dict[key + '_optimizer'] = opt.state_dict()
    for key, sch in (schedulers or {}).items():
        save_dict[key + '_scheduler'] = sch.state_dict()
    torch.save(save_dict, save_path)
    return save_path


def load_ckpt(path: str, nets: Dict[str, torch.nn.Module], optimizers: Optional[Dict[str, object]] = None,
              schedulers: Optional[Dict[str, object]] = None, map_location=None, restore_optimizer: bool = True):
    """trainer/base.py:301-327: networks with ``strict=False``; optimizer / scheduler entries that are missing from the
    file are skipped.  Returns the stored ``clock`` entry (``None`` if absent)."""
    load_path = path if str(path).endswith('.pth') else f'{path}.pth'
    if not os.path.exists(load_path):
        raise ValueError(f"Checkpoint {load_path} not exists.")
    checkpoint = torch.load(load_path, map_location=map_location)
    for key, net in nets.items():
        _unwrap(net).load_state_dict(checkpoint[key + '_net'], strict=False)
    if restore_optimizer:
        for key, opt in (optimizers or {}).items():
            if key + '_optimizer' in checkpoint:
                opt.load_state_dict(checkpoint[key + '_optimizer'])
        for key, sch in (schedulers or {}).items():
            if key + '_scheduler' in checkpoint:
                sch.load_state_dict(checkpoint[key + '_scheduler'])
    return checkpoint.get('clock')


def load_pretrained_model(net: torch.nn.Module, model_name: str, pretrained_path: str, map_location=None) -> None:
    """trainer/trainer_moco_flow.py:46-58: one network out of a trainer checkpoint; for a NeRF entry only the trunk and
    the density head are taken (keys containing ``xyz`` or ``sigma``), the colour branch keeps its initialisation."""
    try:
        weights = torch.load(pretrained_path, map_location=map_location)[model_name]
    except Exception as e:
        raise ValueError("local model {} error! Please check the model.\n{}".format(pretrained_path, e))
    if 'NeRF' in model_name:
        weights = {k: v for k, v in weights.items() if 'xyz' in k or 'sigma' in k}
    _unwrap(net).load_state_dict(weights, strict=False)

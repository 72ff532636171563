"""moco_flow_b200 -- B200-native drop-in for MoCo-Flow's ray-rendering hot path.

Mirrors the import surface of the reference's ``models`` package for that path
(models/__init__.py:1-39): ``Embedding``, ``NeRF``, ``NoF``, ``get_model``, ``get_loss`` and the
functions of ``models/rendering.py``.
"""
from torch import nn as _nn

from .embedding import Embedding
from .nerf import NeRF
from .nof import NoF
from .losses import MSELoss
from .rendering import Draws, nerf_inference, nof_inference, render_rays, sample_pdf
from .ops import check_device

__all__ = ["Embedding", "NeRF", "NoF", "MSELoss", "Draws", "get_model", "get_loss", "render_rays",
           "nerf_inference", "nof_inference", "sample_pdf", "check_device"]


def get_model(model_config):
    """models/__init__.py:8-29 (same config keys, same positional order)."""
    kind = model_config['type']
    if kind == "Embedding":
        return Embedding(model_config['in_channels'], model_config['N_freqs'], model_config['logscale'])
    if kind == "NeRF":
        return NeRF(model_config['D'], model_config['W'], model_config['in_channels_xyz'], model_config['skips'],
                    model_config['extra_feat_type'], model_config['extra_feat_dim'])
    if kind == "NoF":
        return NoF(model_config['D'], model_config['W'], model_config['in_channels_xyz'], model_config['skips'],
                   model_config['extra_feat_type'], model_config['extra_feat_dim'], model_config['use_quat'])
    raise ValueError('model type: {} not valid'.format(kind))


def get_loss(loss_config):
    """models/__init__.py:31-39."""
    kind = loss_config['type']
    if kind == "MSE":
        return MSELoss()
    if kind == 'L1':
        return _nn.L1Loss()
    if kind == 'BCE':
        return _nn.BCELoss()
    raise ValueError('loss type: {} not support'.format(kind))

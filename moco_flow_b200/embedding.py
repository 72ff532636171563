"""Frequency encoder with the reference's ``Embedding`` interface (models/embedding.py:4-47).

Kept: constructor signature, the attributes callers read or assign (``N_freqs``, ``in_channels``,
``out_channels``, ``freq_bands``, and the per-frequency ``weights`` list the trainer overwrites every
step, trainer/trainer_moco_flow.py:289,301-305) and ``set_weights``.  ``forward`` launches the CUDA
encoder (``mcf_pe_fwd`` / ``mcf_pe_bwd``).  The fused render path never calls ``forward``: it reads
``frequencies()`` / ``multipliers()`` and encodes inside the first MLP layer's operand builder.
"""
from __future__ import annotations

from typing import List, Sequence, Union

import torch
from torch import nn

from . import ops


def _frequency_table(n: int, logscale: bool) -> torch.Tensor:
    steps = torch.linspace(0, n - 1, n) if logscale else torch.linspace(1, 2 ** (n - 1), n)
    return 2 ** steps if logscale else steps


class Embedding(nn.Module):
    """x -> [x, w_k sin(f_k x), w_k cos(f_k x)]_k , channel blocks in that order."""

    def __init__(self, in_channels: int, N_freqs: int, logscale: bool = True):
        super().__init__()
        self.in_channels, self.N_freqs = in_channels, N_freqs
        self.freq_bands = _frequency_table(N_freqs, logscale)
        self.weights: Union[List[float], Sequence[float]] = [1 for _ in range(N_freqs)]

    @property
    def out_channels(self) -> int:
        return (2 * self.N_freqs + 1) * self.in_channels

    def set_weights(self, weights) -> None:
        if isinstance(weights, int):
            weights = [weights for _ in range(self.N_freqs)]
        assert len(weights) == self.N_freqs
        self.weights = weights

    def frequencies(self) -> List[float]:
        return [float(f) for f in self.freq_bands.tolist()]

    def multipliers(self) -> List[float]:
        return [float(w) for w in self.weights]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.pe_forward(x, self.frequencies(), self.multipliers())

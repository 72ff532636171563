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

from . import _lib as L
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

    # -- device copy of (frequencies, weights): what the kernels read ------------------------------------
    def device_table(self, device) -> torch.Tensor:
        """[2*MAX_FREQS] floats {freq[], weight[]} on ``device``, brought up to date with ``self.weights``.

        The kernels read the tables through this pointer, not from launch arguments, so a captured CUDA graph keeps
        following the coarse-to-fine schedule that re-assigns ``weights`` every step
        (trainer/trainer_moco_flow.py:280-305): call ``sync_device()`` (or pass the embeddings to
        ``graph.CudaGraphStep(..., refresh=)``) before each replay.  While a stream is capturing, the table must
        already be current -- a changed ``weights`` raises instead of freezing a stale copy into the graph."""
        device = torch.device(device)
        tables = self.__dict__.setdefault("_mcf_tables", {})
        want = tuple(self.frequencies()) + tuple(self.multipliers())
        ent = tables.get(device)
        if ent is not None and ent[1] == want:
            return ent[0]
        if self.N_freqs > L.MAX_FREQS:
            raise ValueError(f"at most {L.MAX_FREQS} frequencies supported")
        if device.type == "cuda" and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("Embedding.weights changed while a CUDA graph is being captured; call sync_device() first")
        host = torch.zeros(2 * L.MAX_FREQS)
        host[:self.N_freqs] = torch.tensor(self.frequencies())
        host[L.MAX_FREQS:L.MAX_FREQS + self.N_freqs] = torch.tensor(self.multipliers())
        if ent is None:
            ent = (host.to(device), want)
        else:
            ent[0].copy_(host)       # same storage: pointers captured in a graph stay valid
            ent = (ent[0], want)
        tables[device] = ent
        return ent[0]

    def sync_device(self) -> None:
        """Refreshes every device table from ``self.weights`` (cheap no-op when nothing changed)."""
        for device in list(self.__dict__.get("_mcf_tables", {})):
            self.device_table(device)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.pe_forward(x, self.frequencies(), self.multipliers(), self.device_table(x.device) if x.is_cuda else None)

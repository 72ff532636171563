"""CUDA-graph capture of a fixed-shape step (render or train).

A 4096-ray training step issues ~190 C-ABI launches plus the optimizer's; launched one by one from Python the
host cannot keep the GPU fed.  The step has no data-dependent shapes or host synchronisation when
``fused_residual_mean=True`` is used, so the whole of it -- weight re-packing, forward, backward, the NCCL
gradient all-reduce and the Adam update -- is captured once and replayed.

Host-side values that change between steps must reach the replay through device memory, because the Python step
function does not run again: the learning rate (``FusedAdam.sync_lr``) and the per-frequency encoder weights of the
reference's coarse-to-fine schedule (``Embedding.weights``, trainer/trainer_moco_flow.py:280-305; the kernels read
them from ``Embedding.device_table``).  Pass such objects as ``refresh=[...]``: before every replay their
``sync_device()`` / ``sync_lr()`` is called, so assigning ``embedding.weights = [...]`` or stepping an LR scheduler
between replays behaves as in eager mode.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class CudaGraphStep:
    """``fn(*inputs) -> Tensor`` captured into one CUDA graph; call with new inputs of the same shapes."""

    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], warmup: int = 3,
                 refresh: Sequence[object] = ()):
        self.fn = fn
        self.refresh = []
        for obj in refresh:
            hook = getattr(obj, "sync_device", None) or getattr(obj, "sync_lr", None)
            if hook is None:
                raise TypeError(f"{type(obj).__name__} has neither sync_device() nor sync_lr()")
            self.refresh.append(hook)
        self.static_in = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):  # also builds every lazily created table / plan before capture
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: other threads of the process (e.g. the NCCL watchdog polling events) must not invalidate
        # the capture
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        for hook in self.refresh:
            hook()
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out

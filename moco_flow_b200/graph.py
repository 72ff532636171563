"""CUDA-graph capture of a fixed-shape step (render or train).

A 4096-ray training step issues ~190 C-ABI launches plus the optimizer's; launched one by one from Python the
host cannot keep the GPU fed.  The step has no data-dependent shapes or host synchronisation when
``fused_residual_mean=True`` is used, so the whole of it -- weight re-packing, forward, backward, the NCCL
gradient all-reduce and the Adam update -- is captured once and replayed.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class CudaGraphStep:
    """``fn(*inputs) -> Tensor`` captured into one CUDA graph; call with new inputs of the same shapes."""

    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], warmup: int = 3):
        self.fn = fn
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
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out

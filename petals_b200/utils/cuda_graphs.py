"""Generic CUDA-graph capture helper (reference: src/petals/utils/cuda_graphs.py:5-76).

The reference graphs six tiny per-op callables (RoPE, RMSNorm, QKV split ...). The engine instead captures a
whole pipeline stage per decode shape (server/stage_engine.py::_capture); this helper remains for user code
that wants the reference's ``make_inference_graphed_callable`` behaviour for an arbitrary function."""
from __future__ import annotations

from typing import Callable, Sequence

import torch


def make_inference_graphed_callable(callable_: Callable, sample_args: Sequence[torch.Tensor], num_warmup_iters: int = 3) -> Callable:
    """Capture ``callable_(*sample_args)`` and return a function that replays it on new inputs of the same shapes.

    Inputs are copied into static buffers before the replay and the static outputs are returned (clone them if
    they must survive the next call). Inference only (no autograd)."""
    assert not isinstance(callable_, torch.nn.Module) or not callable_.training
    if torch.is_autocast_enabled() and torch.is_autocast_cache_enabled():
        raise RuntimeError("make_inference_graphed_callable does not support autocast caching; disable it")
    static_args = tuple(a.clone() if isinstance(a, torch.Tensor) else a for a in sample_args)
    flat = [a for a in static_args if isinstance(a, torch.Tensor)]
    if not flat or not all(a.is_cuda for a in flat):
        return callable_  # nothing to capture off the GPU: the plain callable has the same semantics
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        for _ in range(num_warmup_iters):
            callable_(*static_args)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph), torch.no_grad():
        static_out = callable_(*static_args)

    def replay(*args):
        if len(args) != len(static_args):
            raise TypeError(f"expected {len(static_args)} arguments, got {len(args)}")
        for dst, src in zip(static_args, args):
            if isinstance(dst, torch.Tensor):
                if dst.shape != src.shape or dst.dtype != src.dtype:
                    raise ValueError("graphed callable was captured for different input shapes/dtypes")
                dst.copy_(src)
        graph.replay()
        return static_out

    return replay

"""Stateless span operations behind rpc_forward / rpc_backward
(reference: src/petals/server/block_functions.py:32-141).

The reference chains per-block tasks and, for backward, *re-runs* the forward through all-but-the-last block
and then once more inside hivemind's ``ModuleBackend.backward`` (3 forwards per step).
Here a request is one span-level task: forward is one pass of the stage engine; backward is one forward
(inputs of each block remembered) + one per-block recompute under autograd."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from petals_b200.utils.misc import is_dummy

# Steps with at most this many tokens run the whole span as one atomic runtime task (reference :26).
MAX_SHORT_INFERENCE_TOKENS = 128
# The reference limits NF4 servers to 1-token merged steps because bitsandbytes' 4-bit GEMM is slow for T>1
# (:25-27); the block-scaled FP8 path has no such cliff, the constant is kept for API compatibility.
MAX_NF4_SHORT_INFERENCE_TOKENS = 1


def _split_prompts(prompts: Optional[torch.Tensor], n_blocks: int, batch: int, hidden: int) -> Optional[List[torch.Tensor]]:
    if prompts is None or is_dummy(prompts):
        return None
    if prompts.dim() != 4 or prompts.shape[0] != n_blocks or prompts.shape[1] not in (1, batch) or prompts.shape[3] != hidden:
        raise ValueError(f"prompts must be [{n_blocks}, {batch} or 1, pre_seq_len, {hidden}], got {tuple(prompts.shape)}")
    return list(prompts.unbind(0))


def _check_activations(name: str, t: torch.Tensor, handler) -> None:
    """Shape / dtype of a tensor that kernels will index by the model's hidden size: never trusted from the wire."""
    hidden_size = handler.stage.spec.hidden_size
    if not isinstance(t, torch.Tensor) or t.dim() != 3 or not t.is_floating_point():
        raise ValueError(f"{name} must be a 3-D floating-point tensor [batch, seq, hidden], got {tuple(getattr(t, 'shape', ()))} {getattr(t, 'dtype', type(t))}")
    if t.shape[2] != hidden_size:
        raise ValueError(f"{name} have hidden size {t.shape[2]}, this model's is {hidden_size}")
    if t.shape[0] < 1 or t.shape[1] < 1:
        raise ValueError(f"{name} must contain at least one token, got {tuple(t.shape)}")


def run_rpc_forward(hidden_states: torch.Tensor, prompts: Optional[torch.Tensor], *, backends: Sequence, handler,
                    active_adapter: Optional[str] = None, points: float = 0.0, take_from: Optional[tuple] = None,
                    push_to: Optional[tuple] = None, stash: Optional[str] = None) -> torch.Tensor:
    """``take_from`` / ``push_to`` / ``stash``: the forward micro-batch hop over the NVLink fabric (server/backend.py:Stage.forward);
    ``hidden_states`` is then only a shape carrier (a meta tensor)."""
    _check_activations("hidden_states", hidden_states, handler)
    B, T, H = hidden_states.shape
    block_prompts = _split_prompts(prompts, len(backends), B, H)
    lo, hi = backends[0].slot, backends[-1].slot + 1
    priority = handler.prioritizer.prioritize(hidden_states, points=points / max(len(backends), 1), type="forward")
    hops = {k: v for k, v in (("take_from", take_from), ("push_to", push_to), ("stash", stash)) if v is not None}
    fut = handler.forward_pool.submit_task(hidden_states, block_prompts, lo, hi, active_adapter, *([hops] if hops else []), priority=priority, size=B * T)
    return fut.result(timeout=handler.request_timeout)


def run_rpc_backward(inputs: torch.Tensor, grad_outputs: torch.Tensor, prompts: Optional[torch.Tensor], *, backends: Sequence,
                     handler, active_adapter: Optional[str] = None, points: float = 0.0, grad_from: Optional[tuple] = None,
                     push_to: Optional[tuple] = None, stash: Optional[str] = None) -> List[torch.Tensor]:
    """``grad_from`` / ``push_to`` / ``stash``: the gradient hop over the NVLink fabric (server/backend.py:Stage.backward); ``inputs`` /
    ``grad_outputs`` are then shape carriers (meta tensors)."""
    _check_activations("inputs", inputs, handler)
    _check_activations("grad_outputs", grad_outputs, handler)
    if inputs.shape != grad_outputs.shape:
        raise ValueError(f"inputs {tuple(inputs.shape)} and grad_outputs {tuple(grad_outputs.shape)} must have the same shape")
    B, T, H = inputs.shape
    block_prompts = _split_prompts(prompts, len(backends), B, H)
    lo, hi = backends[0].slot, backends[-1].slot + 1
    priority = handler.prioritizer.prioritize(inputs, grad_outputs, points=points / max(len(backends), 1), type="backward")
    hops = {k: v for k, v in (("grad_from", grad_from), ("push_to", push_to), ("stash", stash)) if v is not None}
    fut = handler.backward_pool.submit_task(inputs, grad_outputs, block_prompts, lo, hi, active_adapter, *([hops] if hops else []), priority=priority, size=B * T)
    grad_inputs, grad_prompts = fut.result(timeout=handler.request_timeout)
    out = [grad_inputs]
    if block_prompts is not None:
        stacked = [gp if gp is not None else torch.zeros_like(p) for gp, p in zip(grad_prompts, block_prompts)]
        # a broadcast prompt (batch 1) receives the sum over the batch, like autograd would produce
        stacked = [g.sum(0, keepdim=True) if (p.shape[0] == 1 and g.shape[0] != 1) else g for g, p in zip(stacked, block_prompts)]
        out.append(torch.stack(stacked, 0))
    return out

"""Stateless forward / backward calls to one stage (reference: src/petals/client/remote_forward_backward.py:1-149).

The reference serialises tensors in executor threads and picks unary vs. streaming RPCs by payload size. In-box
the stub either *is* the stage's handler (tensors by reference) or a Unix-socket proxy that streams raw bytes, so
both functions reduce to a checked call; signatures are kept for callers that subclass the sequence manager."""
from __future__ import annotations

from typing import Any, Dict, Optional, Sequence, Tuple

import torch

from petals_b200.utils.misc import DUMMY, is_dummy


def run_remote_forward(stub, uids: Sequence[str], inputs: torch.Tensor, prompts: Optional[torch.Tensor] = None, *,
                       metadata: Optional[Dict[str, Any]] = None, timeout: Optional[float] = None) -> Tuple[torch.Tensor, ...]:
    """Forward ``inputs`` through blocks ``uids`` on the stage behind ``stub``; returns ``(outputs,)``."""
    if inputs.dim() != 3:
        raise ValueError(f"inputs must be [batch, seq, hidden], got {tuple(inputs.shape)}")
    p = None if prompts is None or is_dummy(prompts) else prompts
    out = stub.rpc_forward(list(uids), inputs.detach(), p, metadata=metadata or {})
    if out.shape != inputs.shape:
        raise RuntimeError(f"stage returned activations of shape {tuple(out.shape)}, expected {tuple(inputs.shape)}")
    return (out,)


def run_remote_backward(stub, uids: Sequence[str], inputs: torch.Tensor, grad_outputs: torch.Tensor, prompts: Optional[torch.Tensor] = None, *,
                        metadata: Optional[Dict[str, Any]] = None, timeout: Optional[float] = None) -> Sequence[torch.Tensor]:
    """Backward through blocks ``uids``: returns ``(grad_inputs,)`` or ``(grad_inputs, grad_prompts)``."""
    p = None if prompts is None or is_dummy(prompts) else prompts
    grads = stub.rpc_backward(list(uids), inputs.detach(), grad_outputs.detach(), p, metadata=metadata or {})
    if grads[0].shape != inputs.shape:
        raise RuntimeError(f"stage returned grad of shape {tuple(grads[0].shape)}, expected {tuple(inputs.shape)}")
    return tuple(grads)

"""Autograd wrappers that keep the training path on the sm_100a kernels.

Stages are frozen: a backward pass through a block only needs the gradient with respect to the *activations*
(reference: src/petals/server/backend.py:101-110 runs the block under ``torch.autograd`` for the same purpose). For a linear
layer that is ``dX = dY @ W`` — the tcgen05 GEMM with W consumed as an MN-major B operand (``gemm(..., b_mn_major=True)``),
so neither direction needs a transposed copy of the weight and no cuBLAS call is issued."""
from __future__ import annotations

from typing import Optional

import torch

from petals_b200.ops import functional as Fn


def tc_linear_supported(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor]) -> bool:
    """Frozen bf16 weight on a CUDA device, activation needs a gradient, shapes tile-aligned for both GEMM directions."""
    return (x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and w.is_cuda and not w.requires_grad
            and (bias is None or not bias.requires_grad) and w.dim() == 2 and w.shape[0] % 64 == 0 and w.shape[1] % 64 == 0
            and w.is_contiguous() and x.numel() > 0)


class TcLinear(torch.autograd.Function):
    """y = x @ w^T (+ bias) with dgrad only (the weight is frozen)."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
        ctx.save_for_backward(w)
        ctx.x_shape = x.shape
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        y = Fn.gemm(x2, w, bias=bias)
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        (w,) = ctx.saved_tensors
        g2 = grad_out.reshape(-1, grad_out.shape[-1])
        if g2.dtype != torch.bfloat16 or not g2.is_contiguous():
            g2 = g2.to(torch.bfloat16).contiguous()
        gx = Fn.gemm(g2, w, b_mn_major=True)  # [M, N_lin] @ W[N_lin, K_lin]: W is the [K, N] operand of this GEMM, used untransposed
        return gx.view(ctx.x_shape), None, None


def tc_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    return TcLinear.apply(x, w, bias)

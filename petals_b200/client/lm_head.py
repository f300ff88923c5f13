"""Output projection of the client model shell (reference: src/petals/client/lm_head.py:15-82).

On a CUDA client the head runs on the GPU: decode-shape inputs (<= 8 rows) use the weight-streaming
``linear_decode`` kernel (vocab x hidden read once per token, final norm fused as its prologue by the
caller), larger inputs use the tcgen05 GEMM. On CPU, low-precision weights are multiplied chunk-wise in
fp32 (``chunked_forward_step`` vocab rows at a time) exactly like the reference's fallback, so a bf16
checkpoint never needs a full fp32 copy of the embedding matrix."""
from __future__ import annotations

import dataclasses
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclasses.dataclass
class LMHeadConfig:
    use_chunked_forward: object = "auto"  # True | False | "auto" (CPU + 16-bit weights)
    chunked_forward_step: int = 16384


class LMHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        if not getattr(config, "tie_word_embeddings", False):
            self.weight = nn.Parameter(torch.zeros(config.vocab_size, config.hidden_size), requires_grad=False)
        else:
            self.weight = None  # set by the model shell to the embedding matrix (tied)
        self.bias = None
        self.in_features, self.out_features = config.hidden_size, config.vocab_size
        self.use_chunked_forward = config.use_chunked_forward
        self.chunked_forward_step = config.chunked_forward_step

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        w = self.weight
        if w.is_cuda and w.dtype == torch.bfloat16 and hidden_states.dtype == torch.bfloat16 and not torch.is_grad_enabled():
            from petals_b200.ops import functional as Fn

            x = hidden_states.contiguous()
            rows = x.numel() // x.shape[-1]
            if rows <= 8 and w.shape[0] % 2 == 0:
                return Fn.linear_decode(x, w)
            if w.shape[0] % 8 == 0 and w.shape[1] % 8 == 0:
                return Fn.gemm(x, w)
        chunked = self.use_chunked_forward
        if chunked == "auto":
            chunked = w.device.type == "cpu" and w.dtype in (torch.float16, torch.bfloat16)
        if chunked:
            return self.chunked_forward(hidden_states)
        return F.linear(hidden_states, w.to(hidden_states.dtype), self.bias)

    def chunked_forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        if self.chunked_forward_step <= 0:
            raise ValueError("chunked_forward_step must be positive")
        w = self.weight
        out = torch.empty(*hidden_states.shape[:-1], self.out_features, dtype=hidden_states.dtype, device=hidden_states.device)
        xf = hidden_states.float()
        for i in range(0, self.out_features, self.chunked_forward_step):
            chunk = w[i: i + self.chunked_forward_step].float()
            out[..., i: i + self.chunked_forward_step] = torch.matmul(xf, chunk.T).to(out.dtype)
        return out

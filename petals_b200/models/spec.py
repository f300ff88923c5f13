"""Architecture-neutral description of one transformer block.

The reference keeps four hand-specialised block wrappers around Hugging Face modules
(src/petals/models/{bloom,llama,falcon,mixtral}/block.py). Here every family is reduced to a
``BlockSpec`` + a weight-name mapping, and both executors — the PyTorch oracle
(``models/block_oracle.py``) and the sm_100a engine (``server/stage_engine.py``) — are written once
against the spec. Canonical parameter names (all stored ``[out, in]`` like ``nn.Linear``):

``ln1_w ln1_b ln2_w ln2_b wqkv bqkv wo bo w_gate w_up w_down b_up b_down router we_gate we_up we_down``
"""
from __future__ import annotations

import dataclasses
import math
from typing import Optional

import torch


@dataclasses.dataclass
class BlockSpec:
    family: str
    hidden_size: int
    num_heads: int
    num_kv_heads: int
    head_dim: int
    intermediate_size: int
    norm: str = "rms"  # "rms" | "layer"
    norm_eps: float = 1e-6
    rotary: bool = True
    rope_theta: float = 10000.0
    rope_scaling: Optional[dict] = None
    max_position: int = 8192
    alibi: bool = False
    qkv_interleaved: bool = False  # fused projection laid out per kv group [G q heads, k, v]
    qkv_bias: bool = False
    out_bias: bool = False
    mlp: str = "swiglu"  # "swiglu" | "gelu" | "moe"
    gelu_tanh: bool = False
    mlp_bias: bool = False
    parallel_attn: bool = False  # Falcon: attention and MLP read the same residual and are summed
    dual_ln: bool = False  # Falcon new decoder architecture: separate ln_attn / ln_mlp
    post_ln_residual: bool = False  # BLOOM apply_residual_connection_post_layernorm
    num_experts: int = 0
    top_k: int = 0
    sliding_window: int = 0
    block_prefix: str = "model.layers"
    # a tensor-parallel shard holds a slice of the heads: ALiBi slopes are those of the *global* head indices
    alibi_total_heads: int = 0  # 0 = num_heads (not sharded)
    alibi_head_offset: int = 0

    @property
    def group_size(self) -> int:
        return self.num_heads // self.num_kv_heads

    @property
    def qkv_dim(self) -> int:
        return (self.num_heads + 2 * self.num_kv_heads) * self.head_dim

    @property
    def attn_scale(self) -> float:
        return 1.0 / math.sqrt(self.head_dim)

    def param_shapes(self) -> dict:
        H, I, Dq = self.hidden_size, self.intermediate_size, self.num_heads * self.head_dim
        shapes = {"ln1_w": (H,), "wqkv": (self.qkv_dim, H), "wo": (H, Dq)}
        if self.norm == "layer":
            shapes["ln1_b"] = (H,)
        if not (self.parallel_attn and not self.dual_ln):
            shapes["ln2_w"] = (H,)
            if self.norm == "layer":
                shapes["ln2_b"] = (H,)
        if self.qkv_bias:
            shapes["bqkv"] = (self.qkv_dim,)
        if self.out_bias:
            shapes["bo"] = (H,)
        if self.mlp == "swiglu":
            shapes.update(w_gate=(I, H), w_up=(I, H), w_down=(H, I))
        elif self.mlp == "gelu":
            shapes.update(w_up=(I, H), w_down=(H, I))
            if self.mlp_bias:
                shapes.update(b_up=(I,), b_down=(H,))
        elif self.mlp == "moe":
            E = self.num_experts
            shapes.update(router=(E, H), we_gate=(E, I, H), we_up=(E, I, H), we_down=(E, H, I))
        return shapes

    def num_params(self) -> int:
        return sum(math.prod(s) for s in self.param_shapes().values())

    def active_params(self) -> int:
        """Parameters ONE token reads (the bytes a decode step must stream): a sparse-MoE block touches the router and only
        ``top_k`` of its ``num_experts`` expert FFNs."""
        total = 0
        for name, shape in self.param_shapes().items():
            n = math.prod(shape)
            if self.mlp == "moe" and name in ("we_gate", "we_up", "we_down"):
                n = n // self.num_experts * self.top_k
            total += n
        return total

    def kv_bytes_per_token(self, dtype: torch.dtype = torch.bfloat16) -> int:
        """K and V for one token of one block (reference: src/petals/server/backend.py:88-99)."""
        return 2 * self.num_kv_heads * self.head_dim * torch.finfo(dtype).bits // 8


def alibi_slopes(num_heads: int) -> torch.Tensor:
    """ALiBi head slopes (Press et al.), the closed form used by BLOOM / Falcon-RW checkpoints."""
    closest = 2 ** math.floor(math.log2(num_heads))
    base = 2.0 ** (-(2.0 ** -(math.log2(closest) - 3)))
    slopes = [base ** (i + 1) for i in range(closest)]
    if closest != num_heads:
        extra_base = 2.0 ** (-(2.0 ** -(math.log2(2 * closest) - 3)))
        n_extra = min(closest, num_heads - closest)
        slopes += [extra_base ** (2 * i + 1) for i in range(n_extra)]
    return torch.tensor(slopes, dtype=torch.float32)

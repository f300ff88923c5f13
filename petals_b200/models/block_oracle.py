"""Plain-PyTorch transformer block driven by a :class:`BlockSpec`.

This is the *oracle*: the readable definition of the math of every supported family (Llama, Mixtral,
BLOOM, Falcon). It runs on CPU (plumbing tests, BASELINE config #1), provides autograd for the
prompt-tuning path where no hand-written backward kernel exists yet, and is the numerical reference
the sm_100a engine is tested against — the same role the unmodified HF blocks play in the reference's
tests (tests/test_optimized_layers.py:187-224, tests/test_block_exact_match.py:12-43).

KV caches use the engine's own layout ``[B, L, Hkv, D]`` (token-major), not the BLOOM layout the
reference forces on every model.
"""
from __future__ import annotations

import os

from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from petals_b200.models.spec import BlockSpec, alibi_slopes
from petals_b200.ops.functional import rope_tables


def _tc_ok(x, w, b) -> bool:
    if not x.is_cuda:
        return False
    from petals_b200.ops.autograd import tc_linear_supported

    return tc_linear_supported(x, w, b)


def _tc_linear(x, w, b):
    from petals_b200.ops.autograd import tc_linear

    return tc_linear(x, w, b)


def _rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    dt = x.dtype
    xf = x.float()
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return w * xf.to(dt)


class GenericBlock(nn.Module):
    """One decoder block. Parameters use the canonical names of ``BlockSpec.param_shapes()``."""

    def __init__(self, spec: BlockSpec, dtype: torch.dtype = torch.float32, device="cpu", init_std: Optional[float] = None):
        super().__init__()
        self.spec = spec
        for name, shape in spec.param_shapes().items():
            if init_std is None:
                t = torch.empty(shape, dtype=dtype, device=device)
            elif name.startswith("ln") and name.endswith("_w"):
                t = torch.ones(shape, dtype=dtype, device=device)
            elif name.startswith("ln") or name.startswith("b"):
                t = torch.zeros(shape, dtype=dtype, device=device)
            else:
                t = (torch.randn(shape, dtype=torch.float32, device=device) * init_std).to(dtype)
            self.register_parameter(name, nn.Parameter(t, requires_grad=False))
        self._rope: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
        self._slopes: Optional[torch.Tensor] = None
        # opt-in (PETALS_B200_TC_BACKWARD=1): frozen bf16 blocks on CUDA run their linears (forward + dgrad) on the tcgen05 GEMM when autograd
        # is recording. Measured on Llama-3-8B prompt tuning: 6.7k vs 7.8k backward tokens/s for cuBLAS at these small-M shapes, so off by default
        self.tc_backward = os.environ.get("PETALS_B200_TC_BACKWARD", "0") != "0"
        # opt-in (PETALS_B200_SDPA_BACKWARD=1): the cache-less recompute pass of rpc_backward runs its attention through the fused
        # scaled_dot_product_attention (one flash kernel each way instead of ~20 eager ones over a [B, H, T, T] fp32 logits tensor).
        # A library call, so only a stop-gap until the flash-attention backward kernel exists; unmeasured, hence off.
        self.sdpa_backward = os.environ.get("PETALS_B200_SDPA_BACKWARD", "0") != "0"
        self.lora: dict = {}  # target param name -> list[(A [r,in], B [out,r], scale)] for the active adapter

    # ---- helpers -----------------------------------------------------------------------------------
    def _p(self, name: str) -> Optional[torch.Tensor]:
        return getattr(self, name, None)

    def _norm(self, x: torch.Tensor, which: str) -> torch.Tensor:
        w, b = self._p(f"{which}_w"), self._p(f"{which}_b")
        if self.spec.norm == "rms":
            return _rms_norm(x, w, self.spec.norm_eps)
        return F.layer_norm(x, (x.shape[-1],), w, b, self.spec.norm_eps)

    def _linear(self, x: torch.Tensor, wname: str, bname: Optional[str] = None, rows: Optional[slice] = None) -> torch.Tensor:
        w = self._p(wname)
        b = self._p(bname) if bname else None
        if rows is not None:
            w = w[rows]
            b = b[rows] if b is not None else None
        if self.tc_backward and torch.is_grad_enabled() and x.requires_grad and _tc_ok(x, w, b):
            y = _tc_linear(x, w, b)  # training path of a frozen stage: forward and dgrad on the tcgen05 GEMM (ops/autograd.py)
        else:
            y = F.linear(x, w, b)
        for (A, Bm, scale, target_rows) in self.lora.get(wname, ()):  # LoRA: y += scale * (x A^T) B^T
            delta = F.linear(F.linear(x, A.to(x.dtype)), Bm.to(x.dtype)) * scale
            if target_rows is None:
                y = y + delta
            else:
                y = y.clone()
                y[..., target_rows] = y[..., target_rows] + delta
        return y

    def rope_cache(self, device) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._rope is None or self._rope[0].device != torch.device(device):
            s = self.spec
            self._rope = rope_tables(s.head_dim, s.max_position, s.rope_theta, s.rope_scaling, device=device)
        return self._rope

    def _split_qkv(self, qkv: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        s = self.spec
        B, T, _ = qkv.shape
        if s.qkv_interleaved:
            g = qkv.view(B, T, s.num_kv_heads, s.group_size + 2, s.head_dim)
            q = g[:, :, :, : s.group_size].reshape(B, T, s.num_heads, s.head_dim)
            return q, g[:, :, :, s.group_size], g[:, :, :, s.group_size + 1]
        h = qkv.view(B, T, s.num_heads + 2 * s.num_kv_heads, s.head_dim)
        return h[:, :, : s.num_heads], h[:, :, s.num_heads: s.num_heads + s.num_kv_heads], h[:, :, s.num_heads + s.num_kv_heads:]

    def _rope_apply(self, x: torch.Tensor, pos: int) -> torch.Tensor:
        cos, sin = self.rope_cache(x.device)
        T = x.shape[1]
        c = torch.cat([cos[pos:pos + T]] * 2, -1).to(x.dtype)[None, :, None, :]
        s_ = torch.cat([sin[pos:pos + T]] * 2, -1).to(x.dtype)[None, :, None, :]
        half = x.shape[-1] // 2
        rot = torch.cat([-x[..., half:], x[..., :half]], -1)
        return x * c + rot * s_

    # ---- sub-layers ----------------------------------------------------------------------------------
    def attention(self, x: torch.Tensor, k_cache: Optional[torch.Tensor], v_cache: Optional[torch.Tensor], pos: int):
        """x: normed input [B,T,H]. Caches [B, Lmax, Hkv, D] are updated in place when given."""
        s = self.spec
        B, T, _ = x.shape
        q, k, v = self._split_qkv(self._linear(x, "wqkv", "bqkv"))
        if s.rotary:
            q, k = self._rope_apply(q, pos), self._rope_apply(k, pos)
        if k_cache is not None:
            k_cache[:, pos:pos + T] = k
            v_cache[:, pos:pos + T] = v
            k, v = k_cache[:, :pos + T], v_cache[:, :pos + T]
        L = k.shape[1]
        G = s.group_size
        if (self.sdpa_backward and k_cache is None and not s.alibi and not s.sliding_window and torch.is_grad_enabled() and x.requires_grad):
            ctx = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True, scale=s.attn_scale,
                                                 enable_gqa=G > 1)
            return self._linear(ctx.transpose(1, 2).reshape(B, T, s.num_heads * s.head_dim), "wo", "bo")
        kf = k.repeat_interleave(G, dim=2) if G > 1 else k
        vf = v.repeat_interleave(G, dim=2) if G > 1 else v
        scores = torch.einsum("bthd,blhd->bhtl", q, kf).float() * s.attn_scale
        qpos = (L - T) + torch.arange(T, device=x.device)[:, None]
        kpos = torch.arange(L, device=x.device)[None, :]
        if s.alibi:
            if self._slopes is None or self._slopes.device != x.device:
                first = s.alibi_head_offset
                self._slopes = alibi_slopes(s.alibi_total_heads or s.num_heads)[first: first + s.num_heads].to(x.device)
            scores = scores + self._slopes.view(1, -1, 1, 1) * (kpos - qpos).float()[None, None]
        ok = kpos <= qpos
        if s.sliding_window:
            ok = ok & (kpos > qpos - s.sliding_window)
        scores = scores.masked_fill(~ok[None, None], float("-inf"))
        probs = torch.softmax(scores, dim=-1).to(x.dtype)
        ctx = torch.einsum("bhtl,blhd->bthd", probs, vf).reshape(B, T, s.num_heads * s.head_dim)
        return self._linear(ctx, "wo", "bo")

    def mlp(self, x: torch.Tensor) -> torch.Tensor:
        s = self.spec
        if s.mlp == "swiglu":
            return self._linear(F.silu(self._linear(x, "w_gate")) * self._linear(x, "w_up"), "w_down")
        if s.mlp == "gelu":
            h = self._linear(x, "w_up", "b_up")
            h = F.gelu(h, approximate="tanh") if s.gelu_tanh else F.gelu(h)
            return self._linear(h, "w_down", "b_down")
        # sparse MoE: softmax router (fp32), top-k, renormalise, SwiGLU experts
        B, T, H = x.shape
        flat = x.reshape(-1, H)
        logits = F.linear(flat, self.router)
        weights = torch.softmax(logits.float(), dim=-1)
        topw, topi = torch.topk(weights, s.top_k, dim=-1)
        topw = (topw / topw.sum(-1, keepdim=True)).to(x.dtype)
        out = torch.zeros_like(flat)
        for e in range(s.num_experts):
            tok, slot = torch.where(topi == e)
            if tok.numel() == 0:
                continue
            xe = flat[tok]
            he = F.silu(F.linear(xe, self.we_gate[e])) * F.linear(xe, self.we_up[e])
            out.index_add_(0, tok, F.linear(he, self.we_down[e]) * topw[tok, slot, None])
        return out.view(B, T, H)

    # ---- block -----------------------------------------------------------------------------------------
    def forward_cached(self, hidden: torch.Tensor, k_cache: Optional[torch.Tensor], v_cache: Optional[torch.Tensor],
                       pos: int = 0) -> torch.Tensor:
        s = self.spec
        if s.parallel_attn:
            a_in = self._norm(hidden, "ln1")
            m_in = self._norm(hidden, "ln2") if s.dual_ln else a_in
            return hidden + self.attention(a_in, k_cache, v_cache, pos) + self.mlp(m_in)
        ln1 = self._norm(hidden, "ln1")
        res = ln1 if s.post_ln_residual else hidden
        h = res + self.attention(ln1, k_cache, v_cache, pos)
        ln2 = self._norm(h, "ln2")
        res = ln2 if s.post_ln_residual else h
        return res + self.mlp(ln2)

    def forward(self, hidden_states: torch.Tensor, layer_past: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                use_cache: bool = False, **_):
        """HF-style entry point: ``(hidden, (k, v))`` with k/v laid out ``[B, L, Hkv, D]``."""
        B, T, _ = hidden_states.shape
        s = self.spec
        if layer_past is None and not use_cache:
            return (self.forward_cached(hidden_states, None, None, 0),)
        P = 0 if layer_past is None else layer_past[0].shape[1]
        k = torch.zeros(B, P + T, s.num_kv_heads, s.head_dim, dtype=hidden_states.dtype, device=hidden_states.device)
        v = torch.zeros_like(k)
        if layer_past is not None:
            k[:, :P], v[:, :P] = layer_past[0], layer_past[1]
        out = self.forward_cached(hidden_states, k, v, P)
        return (out, (k, v)) if use_cache else (out,)

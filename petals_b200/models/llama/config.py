"""Llama family config (reference: src/petals/models/llama/config.py:16-47)."""
from __future__ import annotations

from typing import Dict

import torch

from petals_b200.models.base import DistributedConfig
from petals_b200.models.spec import BlockSpec


class DistributedLlamaConfig(DistributedConfig):
    model_type = "llama"
    block_prefix = "model.layers"
    defaults = dict(vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                    num_attention_heads=32, num_key_value_heads=None, hidden_act="silu", max_position_embeddings=2048,
                    rms_norm_eps=1e-6, rope_theta=10000.0, rope_scaling=None, attention_bias=False, mlp_bias=False,
                    tie_word_embeddings=False, head_dim=None, torch_dtype=None, pretraining_tp=1)
    client_weight_names = {"embed": "model.embed_tokens.weight", "norm_w": "model.norm.weight", "head": "lm_head.weight"}

    def block_spec(self) -> BlockSpec:
        # options of the Hugging Face Llama family that would change the math must not be dropped silently
        if getattr(self, "mlp_bias", False):
            raise NotImplementedError("Llama checkpoints with mlp_bias=true are not supported (the SwiGLU path has no bias tensors)")
        if getattr(self, "hidden_act", "silu") not in ("silu", "swish"):
            raise NotImplementedError(f"hidden_act={self.hidden_act!r} is not supported for Llama blocks (SwiGLU uses SiLU)")
        kv = self.num_key_value_heads or self.num_attention_heads
        rope = getattr(self, "rope_parameters", None) or {}
        theta = rope.get("rope_theta", self.rope_theta) if isinstance(rope, dict) else self.rope_theta
        scaling = self.rope_scaling or (rope if isinstance(rope, dict) and rope.get("rope_type", "default") != "default" else None)
        return BlockSpec(
            family="llama", hidden_size=self.hidden_size, num_heads=self.num_attention_heads, num_kv_heads=kv,
            head_dim=self.head_dim or self.hidden_size // self.num_attention_heads,
            intermediate_size=self.intermediate_size, norm="rms", norm_eps=self.rms_norm_eps, rotary=True,
            rope_theta=theta, rope_scaling=scaling, max_position=max(self.max_position_embeddings, 2048),
            qkv_bias=bool(self.attention_bias), out_bias=bool(self.attention_bias), mlp="swiglu",
            block_prefix=self.block_prefix)

    @classmethod
    def convert_block_weights(cls, hf: Dict[str, torch.Tensor], spec: BlockSpec) -> Dict[str, torch.Tensor]:
        out = {
            "ln1_w": hf["input_layernorm.weight"], "ln2_w": hf["post_attention_layernorm.weight"],
            "wqkv": torch.cat([hf["self_attn.q_proj.weight"], hf["self_attn.k_proj.weight"], hf["self_attn.v_proj.weight"]], 0),
            "wo": hf["self_attn.o_proj.weight"], "w_gate": hf["mlp.gate_proj.weight"], "w_up": hf["mlp.up_proj.weight"],
            "w_down": hf["mlp.down_proj.weight"],
        }
        if spec.qkv_bias:
            out["bqkv"] = torch.cat([hf["self_attn.q_proj.bias"], hf["self_attn.k_proj.bias"], hf["self_attn.v_proj.bias"]], 0)
            out["bo"] = hf["self_attn.o_proj.bias"]
        return out

    @classmethod
    def export_block_weights(cls, canon: Dict[str, torch.Tensor], spec: BlockSpec) -> Dict[str, torch.Tensor]:
        q, k, v = canon["wqkv"].split([spec.num_heads * spec.head_dim, spec.num_kv_heads * spec.head_dim, spec.num_kv_heads * spec.head_dim], 0)
        out = {
            "input_layernorm.weight": canon["ln1_w"], "post_attention_layernorm.weight": canon["ln2_w"],
            "self_attn.q_proj.weight": q, "self_attn.k_proj.weight": k, "self_attn.v_proj.weight": v,
            "self_attn.o_proj.weight": canon["wo"], "mlp.gate_proj.weight": canon["w_gate"],
            "mlp.up_proj.weight": canon["w_up"], "mlp.down_proj.weight": canon["w_down"],
        }
        if spec.qkv_bias:
            bq, bk, bv = canon["bqkv"].split([spec.num_heads * spec.head_dim, spec.num_kv_heads * spec.head_dim, spec.num_kv_heads * spec.head_dim], 0)
            out.update({"self_attn.q_proj.bias": bq, "self_attn.k_proj.bias": bk, "self_attn.v_proj.bias": bv, "self_attn.o_proj.bias": canon["bo"]})
        return out

"""Mixtral (sparse MoE) config (reference: src/petals/models/mixtral/config.py:16-36).
Unlike the reference ``num_key_value_groups`` reflects the real GQA ratio."""
from __future__ import annotations

from typing import Dict

import torch

from petals_b200.models.base import DistributedConfig
from petals_b200.models.spec import BlockSpec


class DistributedMixtralConfig(DistributedConfig):
    model_type = "mixtral"
    block_prefix = "model.layers"
    defaults = dict(vocab_size=32000, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                    num_attention_heads=32, num_key_value_heads=8, max_position_embeddings=4096 * 32, rms_norm_eps=1e-5,
                    rope_theta=1e6, sliding_window=None, num_local_experts=8, num_experts_per_tok=2,
                    tie_word_embeddings=False, head_dim=None, torch_dtype=None)
    client_weight_names = {"embed": "model.embed_tokens.weight", "norm_w": "model.norm.weight", "head": "lm_head.weight"}

    def block_spec(self) -> BlockSpec:
        if getattr(self, "hidden_act", "silu") not in ("silu", "swish"):
            raise NotImplementedError(f"hidden_act={self.hidden_act!r} is not supported for Mixtral experts (SwiGLU uses SiLU)")
        return BlockSpec(
            family="mixtral", hidden_size=self.hidden_size, num_heads=self.num_attention_heads,
            num_kv_heads=self.num_key_value_heads or self.num_attention_heads,
            head_dim=self.head_dim or self.hidden_size // self.num_attention_heads,
            intermediate_size=self.intermediate_size, norm="rms", norm_eps=self.rms_norm_eps, rotary=True,
            rope_theta=self.rope_theta, max_position=min(max(self.max_position_embeddings, 2048), 32768),
            mlp="moe", num_experts=self.num_local_experts, top_k=self.num_experts_per_tok,
            sliding_window=int(self.sliding_window or 0), block_prefix=self.block_prefix)

    @classmethod
    def convert_block_weights(cls, hf: Dict[str, torch.Tensor], spec: BlockSpec) -> Dict[str, torch.Tensor]:
        E = spec.num_experts
        ex = lambda w: torch.stack([hf[f"block_sparse_moe.experts.{e}.{w}.weight"] for e in range(E)], 0)
        return {
            "ln1_w": hf["input_layernorm.weight"], "ln2_w": hf["post_attention_layernorm.weight"],
            "wqkv": torch.cat([hf["self_attn.q_proj.weight"], hf["self_attn.k_proj.weight"], hf["self_attn.v_proj.weight"]], 0),
            "wo": hf["self_attn.o_proj.weight"], "router": hf["block_sparse_moe.gate.weight"],
            "we_gate": ex("w1"), "we_up": ex("w3"), "we_down": ex("w2"),
        }

    @classmethod
    def export_block_weights(cls, canon: Dict[str, torch.Tensor], spec: BlockSpec) -> Dict[str, torch.Tensor]:
        q, k, v = canon["wqkv"].split([spec.num_heads * spec.head_dim, spec.num_kv_heads * spec.head_dim, spec.num_kv_heads * spec.head_dim], 0)
        out = {"input_layernorm.weight": canon["ln1_w"], "post_attention_layernorm.weight": canon["ln2_w"],
               "self_attn.q_proj.weight": q, "self_attn.k_proj.weight": k, "self_attn.v_proj.weight": v,
               "self_attn.o_proj.weight": canon["wo"], "block_sparse_moe.gate.weight": canon["router"]}
        for e in range(spec.num_experts):
            out[f"block_sparse_moe.experts.{e}.w1.weight"] = canon["we_gate"][e]
            out[f"block_sparse_moe.experts.{e}.w3.weight"] = canon["we_up"][e]
            out[f"block_sparse_moe.experts.{e}.w2.weight"] = canon["we_down"][e]
        return out

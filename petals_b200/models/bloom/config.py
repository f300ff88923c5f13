"""BLOOM config (reference: src/petals/models/bloom/config.py:16-35). ALiBi, LayerNorm, tanh-GELU, fused
per-head-interleaved QKV, embedding LayerNorm, tied LM head."""
from __future__ import annotations

from typing import Dict

import torch

from petals_b200.models.base import DistributedConfig
from petals_b200.models.spec import BlockSpec


class DistributedBloomConfig(DistributedConfig):
    model_type = "bloom"
    block_prefix = "h"
    attribute_map = {"num_hidden_layers": "n_layer", "num_attention_heads": "n_head"}
    defaults = dict(vocab_size=250880, hidden_size=64, n_layer=2, n_head=8, layer_norm_epsilon=1e-5,
                    apply_residual_connection_post_layernorm=False, tie_word_embeddings=True, torch_dtype=None)
    client_weight_names = {"embed": "word_embeddings.weight", "embed_ln_w": "word_embeddings_layernorm.weight",
                           "embed_ln_b": "word_embeddings_layernorm.bias", "norm_w": "ln_f.weight", "norm_b": "ln_f.bias"}

    @classmethod
    def default_dht_prefix(cls, name_or_path: str) -> str:
        return super().default_dht_prefix(name_or_path).removesuffix("-hf") + "-petals"

    def block_spec(self) -> BlockSpec:
        H, nh = self.hidden_size, self.n_head
        return BlockSpec(
            family="bloom", hidden_size=H, num_heads=nh, num_kv_heads=nh, head_dim=H // nh, intermediate_size=4 * H,
            norm="layer", norm_eps=self.layer_norm_epsilon, rotary=False, alibi=True, qkv_interleaved=True,
            qkv_bias=True, out_bias=True, mlp="gelu", gelu_tanh=True, mlp_bias=True,
            post_ln_residual=bool(self.apply_residual_connection_post_layernorm), block_prefix=self.block_prefix)

    _MAP = {"ln1_w": "input_layernorm.weight", "ln1_b": "input_layernorm.bias", "ln2_w": "post_attention_layernorm.weight",
            "ln2_b": "post_attention_layernorm.bias", "wqkv": "self_attention.query_key_value.weight",
            "bqkv": "self_attention.query_key_value.bias", "wo": "self_attention.dense.weight", "bo": "self_attention.dense.bias",
            "w_up": "mlp.dense_h_to_4h.weight", "b_up": "mlp.dense_h_to_4h.bias", "w_down": "mlp.dense_4h_to_h.weight",
            "b_down": "mlp.dense_4h_to_h.bias"}

    @classmethod
    def convert_block_weights(cls, hf: Dict[str, torch.Tensor], spec: BlockSpec) -> Dict[str, torch.Tensor]:
        return {c: hf[h] for c, h in cls._MAP.items()}

    @classmethod
    def export_block_weights(cls, canon: Dict[str, torch.Tensor], spec: BlockSpec) -> Dict[str, torch.Tensor]:
        return {h: canon[c] for c, h in cls._MAP.items()}

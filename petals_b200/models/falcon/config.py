"""Falcon config (reference: src/petals/models/falcon/config.py:16-48): 7B (multi-query, parallel attention,
single LayerNorm), 40B/180B (new decoder architecture: grouped KV, two parallel LayerNorms) and the RW
variants (ALiBi, sequential)."""
from __future__ import annotations

from typing import Dict

import torch

from petals_b200.models.base import DistributedConfig
from petals_b200.models.spec import BlockSpec


class DistributedFalconConfig(DistributedConfig):
    model_type = "falcon"
    block_prefix = "transformer.h"
    defaults = dict(vocab_size=65024, hidden_size=4544, num_hidden_layers=32, num_attention_heads=71, num_kv_heads=None,
                    layer_norm_epsilon=1e-5, alibi=False, new_decoder_architecture=False, multi_query=True,
                    parallel_attn=True, bias=False, rope_theta=10000.0, max_position_embeddings=2048,
                    tie_word_embeddings=True, torch_dtype=None, ffn_hidden_size=None)
    client_weight_names = {"embed": "transformer.word_embeddings.weight", "norm_w": "transformer.ln_f.weight",
                           "norm_b": "transformer.ln_f.bias", "head": "lm_head.weight"}

    def block_spec(self) -> BlockSpec:
        nh = self.num_attention_heads
        if self.new_decoder_architecture:
            kv = self.num_kv_heads or nh
        elif self.multi_query:
            kv = 1
        else:
            kv = nh
        H = self.hidden_size
        # fused QKV layout: new arch -> [kv, G+2, D]; old multi-query -> [q heads | k | v]; old MHA -> per-head [q,k,v]
        interleaved = bool(self.new_decoder_architecture) or not self.multi_query
        return BlockSpec(
            family="falcon", hidden_size=H, num_heads=nh, num_kv_heads=kv, head_dim=H // nh,
            intermediate_size=self.ffn_hidden_size or 4 * H, norm="layer", norm_eps=self.layer_norm_epsilon,
            rotary=not self.alibi, rope_theta=self.rope_theta, max_position=max(self.max_position_embeddings, 8192),
            alibi=bool(self.alibi), qkv_interleaved=interleaved, qkv_bias=bool(self.bias), out_bias=bool(self.bias),
            mlp="gelu", gelu_tanh=False, mlp_bias=bool(self.bias),
            parallel_attn=bool(self.parallel_attn or self.new_decoder_architecture),
            dual_ln=bool(self.new_decoder_architecture), block_prefix=self.block_prefix)

    @classmethod
    def _map(cls, spec: BlockSpec) -> Dict[str, str]:
        m = {"wqkv": "self_attention.query_key_value.weight", "wo": "self_attention.dense.weight",
             "w_up": "mlp.dense_h_to_4h.weight", "w_down": "mlp.dense_4h_to_h.weight"}
        if spec.dual_ln:
            m.update(ln1_w="ln_attn.weight", ln1_b="ln_attn.bias", ln2_w="ln_mlp.weight", ln2_b="ln_mlp.bias")
        else:
            m.update(ln1_w="input_layernorm.weight", ln1_b="input_layernorm.bias")
            if not spec.parallel_attn:
                m.update(ln2_w="post_attention_layernorm.weight", ln2_b="post_attention_layernorm.bias")
        if spec.qkv_bias:
            m.update(bqkv="self_attention.query_key_value.bias", bo="self_attention.dense.bias",
                     b_up="mlp.dense_h_to_4h.bias", b_down="mlp.dense_4h_to_h.bias")
        return m

    @classmethod
    def convert_block_weights(cls, hf: Dict[str, torch.Tensor], spec: BlockSpec) -> Dict[str, torch.Tensor]:
        return {c: hf[h] for c, h in cls._map(spec).items()}

    @classmethod
    def export_block_weights(cls, canon: Dict[str, torch.Tensor], spec: BlockSpec) -> Dict[str, torch.Tensor]:
        return {h: canon[c] for c, h in cls._map(spec).items()}

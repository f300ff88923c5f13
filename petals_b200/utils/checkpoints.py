"""Synthetic Hugging Face-layout checkpoints (random init) for tests and benchmarks.

There is no network on the build/GPU boxes, so models are materialised locally in exactly the layout the
reference loads from the hub: ``config.json`` + sharded ``*.safetensors`` + ``model.safetensors.index.json``
with one shard per group of blocks (so per-block loading really does skip unrelated files, like
src/petals/server/from_pretrained.py:93-109)."""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import torch

from petals_b200.models.block_oracle import GenericBlock
from petals_b200.utils.safetensors_io import save_file

TINY = {
    "llama": dict(vocab_size=512, hidden_size=128, intermediate_size=256, num_hidden_layers=4, num_attention_heads=4,
                  num_key_value_heads=2, max_position_embeddings=512, rms_norm_eps=1e-5, rope_theta=10000.0),
    "mixtral": dict(vocab_size=512, hidden_size=128, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4,
                    num_key_value_heads=2, max_position_embeddings=512, num_local_experts=4, num_experts_per_tok=2,
                    sliding_window=None, rms_norm_eps=1e-5),
    "bloom": dict(vocab_size=512, hidden_size=128, n_layer=4, n_head=4, layer_norm_epsilon=1e-5),
    "falcon": dict(vocab_size=512, hidden_size=128, num_hidden_layers=4, num_attention_heads=4, num_kv_heads=2,
                   new_decoder_architecture=True, multi_query=True, parallel_attn=True, bias=False, alibi=False),
}


def config_class(model_type: str):
    from petals_b200.utils.auto_config import get_model_classes

    return get_model_classes(model_type)["config"]


def make_random_checkpoint(path: str, model_type: str = "llama", *, dtype: torch.dtype = torch.float32, seed: int = 0,
                           blocks_per_shard: int = 2, init_std: float = 0.05, **overrides) -> str:
    """Write a random-init checkpoint of `model_type` to `path` and return `path`."""
    os.makedirs(path, exist_ok=True)
    cfg_cls = config_class(model_type)
    values = dict(TINY.get(model_type, {}))
    values.update(overrides)
    config = cfg_cls(**values)
    config.torch_dtype = dtype
    config.architectures = [{"llama": "LlamaForCausalLM", "mixtral": "MixtralForCausalLM", "bloom": "BloomForCausalLM",
                             "falcon": "FalconForCausalLM"}[model_type]]
    config.save_pretrained(path)
    spec = config.block_spec()
    gen = torch.Generator().manual_seed(seed)
    weight_map: Dict[str, str] = {}
    H, V = config.hidden_size, config.vocab_size

    def rnd(*shape, std=init_std):
        return (torch.randn(*shape, generator=gen) * std).to(dtype)

    client: Dict[str, torch.Tensor] = {}
    names = cfg_cls.client_weight_names
    client[names["embed"]] = rnd(V, H, std=0.5)
    for key in ("norm_w", "embed_ln_w"):
        if key in names:
            client[names[key]] = (1.0 + 0.1 * torch.randn(H, generator=gen)).to(dtype)
    for key in ("norm_b", "embed_ln_b"):
        if key in names:
            client[names[key]] = (0.1 * torch.randn(H, generator=gen)).to(dtype)
    if "head" in names and not getattr(config, "tie_word_embeddings", False):
        client[names["head"]] = rnd(V, H, std=0.1)
    n_layers = config.num_hidden_layers
    n_shards = 1 + (n_layers + blocks_per_shard - 1) // blocks_per_shard
    shard_name = lambda i: f"model-{i + 1:05d}-of-{n_shards:05d}.safetensors"
    save_file(client, os.path.join(path, shard_name(0)), metadata={"format": "pt"})
    weight_map.update({k: shard_name(0) for k in client})
    total = sum(t.numel() * t.element_size() for t in client.values())
    for s in range(1, n_shards):
        shard: Dict[str, torch.Tensor] = {}
        for layer in range((s - 1) * blocks_per_shard, min(n_layers, s * blocks_per_shard)):
            torch.manual_seed(seed * 1000003 + layer)
            block = GenericBlock(spec, dtype=torch.float32, init_std=init_std)
            canon = {}
            for k, v in block.state_dict().items():
                if k.startswith("ln") and k.endswith("_w"):
                    v = 1.0 + 0.1 * torch.randn(v.shape)
                elif k.startswith("ln") or k.startswith("b"):
                    v = 0.05 * torch.randn(v.shape)
                canon[k] = v.to(dtype)
            for hf_name, t in cfg_cls.export_block_weights(canon, spec).items():
                shard[f"{config.block_prefix}.{layer}.{hf_name}"] = t.contiguous()
        save_file(shard, os.path.join(path, shard_name(s)), metadata={"format": "pt"})
        weight_map.update({k: shard_name(s) for k in shard})
        total += sum(t.numel() * t.element_size() for t in shard.values())
    with open(os.path.join(path, "model.safetensors.index.json"), "w") as f:
        json.dump({"metadata": {"total_size": total}, "weight_map": weight_map}, f, indent=1)
    return path

"""Random-init models materialised directly on the device (benchmarks, smoke tests).

BASELINE.json names Llama-3-70B-class configs; there is no network and a 140 GB checkpoint round trip through the
disk would dominate every run, so stages and client shells can be built from a config with weights drawn on the GPU
(same tensor shapes, dtypes and memory layout as a loaded checkpoint — the serving path is identical)."""
from __future__ import annotations

import json
import os
import tempfile
from typing import List, Optional, Sequence

import torch

from petals_b200.data_structures import ModelInfo, ServerInfo, ServerState
from petals_b200.models.block_oracle import GenericBlock
from petals_b200.parallel.swarm import Swarm
from petals_b200.utils.convert_block import QuantType

MODEL_PRESETS = {
    "llama-3-70b": dict(model_type="llama", vocab_size=128256, hidden_size=8192, intermediate_size=28672, num_hidden_layers=80,
                        num_attention_heads=64, num_key_value_heads=8, max_position_embeddings=8192, rms_norm_eps=1e-5, rope_theta=500000.0),
    "llama-3-8b": dict(model_type="llama", vocab_size=128256, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                       num_attention_heads=32, num_key_value_heads=8, max_position_embeddings=8192, rms_norm_eps=1e-5, rope_theta=500000.0),
    "llama-tiny": dict(model_type="llama", vocab_size=4096, hidden_size=1024, intermediate_size=2816, num_hidden_layers=4,
                       num_attention_heads=8, num_key_value_heads=2, max_position_embeddings=2048, rms_norm_eps=1e-5, rope_theta=10000.0),
    "mixtral-tiny": dict(model_type="mixtral", vocab_size=4096, hidden_size=1024, intermediate_size=1408, num_hidden_layers=3,
                         num_attention_heads=8, num_key_value_heads=2, num_local_experts=8, num_experts_per_tok=2, rms_norm_eps=1e-5,
                         rope_theta=1e6, max_position_embeddings=2048, sliding_window=None),
    "mixtral-8x7b": dict(model_type="mixtral", vocab_size=32000, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                         num_attention_heads=32, num_key_value_heads=8, num_local_experts=8, num_experts_per_tok=2, rms_norm_eps=1e-5,
                         rope_theta=1e6, max_position_embeddings=32768),
    # Falcon layouts (reference GPU test: tests/test_optimized_layers.py:187-224): 7B-style = multi-query + parallel attention + one
    # LayerNorm; 40B-style = new decoder architecture (grouped KV, two parallel LayerNorms, interleaved fused QKV); RW = ALiBi, sequential
    "falcon-tiny-7b": dict(model_type="falcon", vocab_size=4096, hidden_size=1024, num_hidden_layers=3, num_attention_heads=16, multi_query=True,
                           parallel_attn=True, new_decoder_architecture=False, alibi=False, bias=False, layer_norm_epsilon=1e-5),
    "falcon-tiny-40b": dict(model_type="falcon", vocab_size=4096, hidden_size=1024, num_hidden_layers=3, num_attention_heads=8, num_kv_heads=2,
                            new_decoder_architecture=True, parallel_attn=True, alibi=False, bias=False, layer_norm_epsilon=1e-5),
    "falcon-tiny-rw": dict(model_type="falcon", vocab_size=4096, hidden_size=1024, num_hidden_layers=3, num_attention_heads=16, multi_query=False,
                           parallel_attn=False, new_decoder_architecture=False, alibi=True, bias=True, layer_norm_epsilon=1e-5),
    "bloom-560m": dict(model_type="bloom", vocab_size=250880, hidden_size=1024, n_layer=24, n_head=16, layer_norm_epsilon=1e-5),
}


def write_config_only(name: str, overrides: Optional[dict] = None, path: Optional[str] = None) -> str:
    """A checkpoint directory containing only config.json (weights are random-initialised on the device)."""
    cfg = dict(MODEL_PRESETS[name])
    cfg.update(overrides or {})
    cfg.setdefault("torch_dtype", "bfloat16")
    if path is None:
        # deterministic across processes: every rank of a job must derive the same dht_prefix from the directory name
        import hashlib

        tag = hashlib.sha1(json.dumps(cfg, sort_keys=True).encode()).hexdigest()[:8]
        path = os.path.join(tempfile.gettempdir(), f"petals_b200_{name}_{tag}")
    os.makedirs(path, exist_ok=True)
    tmp = os.path.join(path, f"config.json.{os.getpid()}")
    with open(tmp, "w") as f:
        json.dump(cfg, f)
    os.replace(tmp, os.path.join(path, "config.json"))
    return path


def random_blocks(config, block_indices: Sequence[int], device, dtype=torch.bfloat16, seed: int = 0, init_std: float = 0.02) -> List[GenericBlock]:
    spec = config.block_spec()
    blocks = []
    for i in block_indices:
        torch.manual_seed(seed * 1000003 + i)
        blocks.append(GenericBlock(spec, dtype=dtype, device=device, init_std=init_std))
    return blocks


def launch_random_stage(model_path: str, block_indices: Sequence[int], swarm: Swarm, device, *, dtype=torch.bfloat16, seed: int = 0,
                        attn_cache_tokens: int = 4096, inference_max_length: int = 4096, max_batch_size: int = 65536, peer_id: Optional[str] = None,
                        use_cuda_graphs: bool = True, force_oracle: bool = False, max_chunk_size_bytes: int = 256 * 1024 * 1024,
                        quant_type: QuantType = QuantType.NONE):
    """Start serving ``block_indices`` with random weights; returns the ModuleContainer (call ``.shutdown()``)."""
    from petals_b200.server.server import ModuleContainer
    from petals_b200.utils.auto_config import AutoDistributedConfig
    import petals_b200

    config = AutoDistributedConfig.from_pretrained(model_path)
    device = torch.device(device)
    blocks = random_blocks(config, block_indices, device, dtype, seed)
    info = ServerInfo(state=ServerState.JOINING, throughput=1.0, version=petals_b200.__version__, torch_dtype=str(dtype).replace("torch.", ""),
                      quant_type=quant_type.name.lower(), using_relay=False)
    return ModuleContainer.create(
        dht=swarm, dht_prefix=config.dht_prefix, converted_model_name_or_path=model_path, block_config=config,
        attn_cache_tokens=attn_cache_tokens, server_info=info, model_info=ModelInfo(num_blocks=config.num_hidden_layers, repository=model_path),
        block_indices=list(block_indices), min_batch_size=1, max_batch_size=max_batch_size, max_chunk_size_bytes=max_chunk_size_bytes,
        max_alloc_timeout=600, inference_max_length=inference_max_length, torch_dtype=dtype, device=device, quant_type=quant_type,
        tensor_parallel_devices=(device,), adapters=(), update_period=30, expiration=3600, request_timeout=180, session_timeout=1800,
        step_timeout=300, stats_report_interval=None, peer_id=peer_id or f"{device.type}{device.index or 0}-stage{block_indices[0]}",
        use_cuda_graphs=use_cuda_graphs, force_oracle=force_oracle, prebuilt_blocks=blocks)


def random_client_model(model_path: str, swarm: Swarm, device, *, dtype=torch.bfloat16, seed: int = 1, model_class: str = "model_for_causal_lm", **kwargs):
    """A client shell with random embeddings / head (tied configs share them) attached to ``swarm``."""
    from petals_b200.utils.auto_config import detect_model_type, get_model_classes

    classes = get_model_classes(detect_model_type(model_path))
    kwargs.setdefault("max_retries", 3)  # synthetic-model runs should fail fast instead of retrying forever
    config = classes["config"].from_pretrained(model_path, initial_peers=[swarm.address], **kwargs)
    torch.manual_seed(seed)
    with torch.device(device):
        model = classes[model_class](config, dht=swarm)
    model = model.to(dtype)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() > 1 and "prompt" not in name:
                p.normal_(0.0, 0.02)
    model.float_trainable_()
    return model.eval()

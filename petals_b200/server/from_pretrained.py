"""Per-block weight loading from a Hugging Face checkpoint layout
(reference: src/petals/server/from_pretrained.py:35-224).

Only the shards that contain ``"<block_prefix>.<idx>."`` tensors are opened (via the index json), tensors
are filtered by key inside each shard, the prefix is stripped and HF names are converted to the canonical
fused layout of the family (QKV concatenated once at load time). Reading uses the native mmap reader
(csrc/runtime/safetensors_reader.cpp); there is no hub download / disk-space eviction loop because the box
has no network — ``resolve_model_path`` finds a local directory or an existing hub cache snapshot."""
from __future__ import annotations

import json
import os
from typing import Dict, Optional, Union

import torch

from petals_b200.utils.paths import resolve_model_path
from petals_b200.server.block_utils import get_model_block, resolve_block_dtype
from petals_b200.utils.auto_config import AutoDistributedConfig
from petals_b200.utils.logging import get_logger
from petals_b200.utils.safetensors_io import SafetensorsFile

logger = get_logger(__name__)
INDEX_NAMES = ("model.safetensors.index.json",)
SINGLE_NAMES = ("model.safetensors",)


def _shards_for_prefix(path: str, prefix: str) -> list:
    for name in INDEX_NAMES:
        index_path = os.path.join(path, name)
        if os.path.exists(index_path):
            with open(index_path) as f:
                weight_map = json.load(f)["weight_map"]
            files = sorted({fn for key, fn in weight_map.items() if key.startswith(prefix)})
            return [os.path.join(path, fn) for fn in files]
    for name in SINGLE_NAMES:
        if os.path.exists(os.path.join(path, name)):
            return [os.path.join(path, name)]
    raise FileNotFoundError(f"no safetensors weights found in {path} (only safetensors checkpoints are supported)")


def load_block_state(model_name: str, block_index: int, *, config=None, pinned: bool = False) -> Dict[str, torch.Tensor]:
    """HF tensors of one block with the ``<block_prefix>.<idx>.`` prefix stripped (checkpoint dtype)."""
    config = config or AutoDistributedConfig.from_pretrained(model_name)
    path = resolve_model_path(model_name)
    prefixes = [f"{config.block_prefix}.{block_index}."]
    if config.model_type == "bloom":
        prefixes.append(f"transformer.{config.block_prefix}.{block_index}.")  # some BLOOM exports keep "transformer."
    state: Dict[str, torch.Tensor] = {}
    for prefix in prefixes:
        try:
            shards = _shards_for_prefix(path, prefix)
        except FileNotFoundError:
            raise
        for shard in shards:
            with SafetensorsFile(shard) as f:
                for key in f.keys():
                    if key.startswith(prefix):
                        state[key[len(prefix):]] = f.get_tensor(key, pinned=pinned)
        if state:
            break
    if not state:
        raise KeyError(f"checkpoint {model_name} has no tensors for block {block_index}")
    return state


def load_canonical_block(model_name: str, block_index: int, *, config=None, torch_dtype="auto", pinned: bool = False) -> Dict[str, torch.Tensor]:
    """Canonical (fused) tensors of one block, cast to the serving dtype."""
    config = config or AutoDistributedConfig.from_pretrained(model_name)
    dtype = resolve_block_dtype(config, torch_dtype)
    hf = load_block_state(model_name, block_index, config=config, pinned=pinned)
    canon = type(config).convert_block_weights(hf, config.block_spec())
    return {k: (v.to(dtype) if v.is_floating_point() else v).contiguous() for k, v in canon.items()}


def load_pretrained_block(model_name: str, block_index: int, *, config=None, torch_dtype: Union[torch.dtype, str] = "auto",
                          revision: Optional[str] = None, token=None, cache_dir: Optional[str] = None,
                          max_disk_space: Optional[int] = None, device="cpu"):
    """A ready-to-run block module (oracle executor) holding block ``block_index`` of ``model_name``."""
    config = config or AutoDistributedConfig.from_pretrained(model_name)
    dtype = resolve_block_dtype(config, torch_dtype)
    block = get_model_block(config, layer_idx=block_index, dtype=dtype, device="cpu")
    canon = load_canonical_block(model_name, block_index, config=config, torch_dtype=dtype)
    expected = set(block.state_dict().keys())
    missing, unexpected = expected - set(canon), set(canon) - expected
    if missing or unexpected:
        raise RuntimeError(f"block {block_index}: missing {sorted(missing)}, unexpected {sorted(unexpected)}")
    # the kernels size their launches from the config (BlockSpec): a checkpoint whose tensors disagree with its own config.json must be
    # refused here, not discovered as an out-of-bounds read on the device
    wrong = {name: (tuple(tensor.shape), tuple(getattr(block, name).shape)) for name, tensor in canon.items()
             if tuple(tensor.shape) != tuple(getattr(block, name).shape)}
    if wrong:
        details = ", ".join(f"{n}: checkpoint {got} vs config {want}" for n, (got, want) in sorted(wrong.items()))
        raise RuntimeError(f"block {block_index} of {model_name}: tensor shapes do not match config.json ({details})")
    for name, tensor in canon.items():
        getattr(block, name).data = tensor
    block.requires_grad_(False)
    logger.debug(f"loaded {model_name} block {block_index} ({dtype})")
    return block.to(device)

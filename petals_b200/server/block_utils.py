"""Dtype / size helpers for blocks (reference: src/petals/server/block_utils.py:1-65)."""
from __future__ import annotations

from typing import Optional, Union

import torch

from petals_b200.utils.convert_block import QuantType


def resolve_block_dtype(config, dtype: Union[str, torch.dtype, None]) -> torch.dtype:
    """"auto" -> the checkpoint dtype, except fp32 checkpoints are served in bf16 on GPUs."""
    if dtype not in ("auto", None):
        return dtype
    cfg_dtype = getattr(config, "torch_dtype", None)
    if cfg_dtype in ("auto", None, torch.float32):
        return torch.bfloat16
    return cfg_dtype


def get_block_size(config, location: str, *, dtype: Optional[Union[str, torch.dtype]] = None,
                   quant_type: QuantType = QuantType.NONE, eps: float = 0.01) -> int:
    """Bytes one block occupies in ``location`` ("memory" = device footprint, "disk" = checkpoint)."""
    if location not in ("memory", "disk"):
        raise ValueError('location must be "memory" or "disk"')
    n_params = config.block_spec().num_params()
    if location == "memory":
        if quant_type == QuantType.NONE:
            dtype = resolve_block_dtype(config, dtype)
            bytes_per_value = torch.finfo(dtype).bits / 8
        elif quant_type in (QuantType.FP8, QuantType.INT8):
            bytes_per_value = 1 + 1 / 32  # e4m3 payload + one UE8M0 scale per 32 values
        elif quant_type == QuantType.NF4:
            bytes_per_value = 4.25 / 8
        else:
            raise ValueError(f"unsupported quant_type {quant_type}")
    else:
        cfg_dtype = getattr(config, "torch_dtype", None)
        bytes_per_value = torch.finfo(cfg_dtype if isinstance(cfg_dtype, torch.dtype) else torch.float32).bits / 8
    return round(n_params * bytes_per_value * (1 + eps))


def get_model_block(config, layer_idx: int = 0, dtype: torch.dtype = torch.float32, device="cpu"):
    """An uninitialised block module of the config's family."""
    from petals_b200.utils.auto_config import get_model_classes

    block_cls = get_model_classes(config.model_type)["block"]
    return block_cls(config, layer_idx=layer_idx, dtype=dtype, device=device)

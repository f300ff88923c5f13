"""Post-load block conversion: freeze -> (tensor-parallel plan) -> quantise -> place -> adapters
(reference: src/petals/utils/convert_block.py:25-156).

Differences by design: quantisation is block-scaled FP8 (``--quant_type fp8``; the legacy names ``int8`` /
``nf4`` are accepted and mapped to it with a warning because bitsandbytes has no Blackwell kernels), and
tensor parallelism is not a module wrapper with per-device threads but a sharding plan consumed by the
stage engine, one process per GPU (parallel/tensor_parallel.py)."""
from __future__ import annotations

from enum import Enum
from typing import Optional, Sequence

import torch
import torch.nn as nn

from petals_b200.utils.logging import get_logger

logger = get_logger(__name__)


class QuantType(Enum):
    NONE = 0
    INT8 = 1  # accepted for CLI compatibility -> served as FP8
    NF4 = 2  # accepted for CLI compatibility -> served as FP8
    FP8 = 3  # block-scaled MXFP8 (E4M3 + UE8M0 per 32)


QUANTIZABLE = ("wqkv", "wo", "w_gate", "w_up", "w_down", "we_gate", "we_up", "we_down")


def resolve_quant_type(quant_type) -> QuantType:
    if isinstance(quant_type, str):
        if quant_type.upper() not in QuantType.__members__:
            raise ValueError(f"quant_type must be one of {[q.name.lower() for q in QuantType]}, got {quant_type!r}")
        quant_type = QuantType[quant_type.upper()]
    if quant_type in (QuantType.INT8, QuantType.NF4):
        logger.warning(f"quant_type={quant_type.name.lower()} has no sm_100 kernels (bitsandbytes); using block-scaled FP8 instead")
        return QuantType.FP8
    return quant_type or QuantType.NONE


def quantize_module(block: nn.Module, *, quant_type: QuantType) -> nn.Module:
    """Oracle-side quantisation: weights are replaced by their MXFP8 round trip (engine keeps the 1-byte form)."""
    from petals_b200.ops.quant import fake_quantize_mxfp8

    quant_type = resolve_quant_type(quant_type)
    if quant_type == QuantType.NONE:
        return block
    for name in QUANTIZABLE:
        p = getattr(block, name, None)
        if p is not None and p.shape[-1] % 32 == 0:
            p.data = fake_quantize_mxfp8(p.data)
    block.quant_type = quant_type
    return block


def convert_block(block: nn.Module, block_index: int, config, tensor_parallel_devices: Sequence[torch.device],
                  output_device: torch.device, quant_type: QuantType, freeze: bool = True,
                  adapters: Optional[Sequence[str]] = None, **kwargs) -> nn.Module:
    """Prepare a freshly loaded block for serving. Returns the same module (moved / quantised / with adapters)."""
    if freeze:
        block.requires_grad_(False)
    devices = tuple(torch.device(d) for d in tensor_parallel_devices)
    if len(devices) > 1:
        # same order as the reference (freeze -> make_tensor_parallel -> quantize -> adapters, convert_block.py:25-73); the generic
        # split runs on the oracle blocks, so it is combined with neither weight quantisation nor LoRA
        if quant_type != QuantType.NONE or adapters:
            raise ValueError("tensor-parallel blocks serve unquantised checkpoints without adapters; use pipeline stages for quantised / LoRA serving")
        from petals_b200.parallel.tp_generic import make_tensor_parallel

        block = make_tensor_parallel(block, config.block_spec(), devices)
        block.tensor_parallel_devices = devices
        return block
    block = quantize_module(block, quant_type=quant_type)
    block.tensor_parallel_devices = devices
    block = block.to(output_device)
    if adapters:
        from petals_b200.utils.peft import add_adapter_to_block, create_lora_adapter, load_peft

        create_lora_adapter(block)
        for adapter_name in adapters:
            adapter_config, adapter_state = load_peft(adapter_name, block_idx=block_index, **kwargs)
            add_adapter_to_block(block, block_index, adapter_name, adapter_config, adapter_state)
    return block


def check_device_balance(devices: Sequence[torch.device]) -> None:
    if not all(d.type == "cuda" for d in devices):
        return
    unique = sorted(set(d.index for d in devices))
    mem = [torch.cuda.get_device_properties(i).total_memory for i in unique]
    if min(mem) < 0.9 * max(mem):
        logger.warning("tensor-parallel devices have uneven memory; shards are equal-sized so the smallest GPU limits the span")

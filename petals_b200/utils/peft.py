"""LoRA adapters served next to the frozen blocks (reference: src/petals/utils/peft.py:31-283).

Kept behaviour: adapters are read from PEFT-format directories, **safetensors only** (arbitrary pickles are
refused), only the tensors of the requested block are loaded, bias terms and dropout are not supported,
``scale = lora_alpha / r``. Changed by design: the active adapter is a per-request property of a stage
(``Stage.use_adapter``), not a process-global class attribute, and LoRA factors attach
to the canonical fused projections — an adapter on ``q_proj`` updates the first ``Hq*D`` output rows of the
fused QKV weight.
"""
from __future__ import annotations

import contextlib
import json
import os
import re
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from petals_b200.utils.paths import resolve_model_path
from petals_b200.utils.logging import get_logger
from petals_b200.utils.safetensors_io import SafetensorsFile

logger = get_logger(__name__)
ADAPTER_CONFIG, ADAPTER_WEIGHTS = "adapter_config.json", "adapter_model.safetensors"


def check_peft_repository(path: str) -> bool:
    """True iff the adapter ships safetensors weights (the only format that is ever loaded)."""
    try:
        p = resolve_model_path(path)
    except FileNotFoundError:
        return False
    return os.path.exists(os.path.join(p, ADAPTER_WEIGHTS))


def load_specific_module(block_idx: int, filepath: str, framework: str = "pt", device: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Tensors of one block from an adapter file, selected by the ``.<idx>.`` component of their names."""
    out = {}
    pattern = re.compile(rf"\.{block_idx}\.")
    with SafetensorsFile(filepath) as f:
        for key in f.keys():
            if pattern.search(key):
                out[key] = f.get_tensor(key)
    if not out:
        logger.warning(f"adapter file {filepath} has no tensors for block {block_idx}")
    return out


def load_peft(repo_id: str, block_idx: Optional[int] = None, device: Optional[int] = None, **_) -> Tuple[dict, Dict[str, torch.Tensor]]:
    if not check_peft_repository(repo_id):
        raise ValueError(f"adapter {repo_id!r} not found or has no {ADAPTER_WEIGHTS}: only safetensors adapters are supported")
    path = resolve_model_path(repo_id)
    with open(os.path.join(path, ADAPTER_CONFIG)) as f:
        config = json.load(f)
    if config.get("peft_type", "LORA") != "LORA":
        raise NotImplementedError(f"peft_type={config.get('peft_type')} is not supported (LoRA only)")
    if config.get("bias", "none") != "none":
        raise NotImplementedError("LoRA adapters with bias are not supported")
    weights = os.path.join(path, ADAPTER_WEIGHTS)
    if block_idx is None:
        with SafetensorsFile(weights) as f:
            state = {k: f.get_tensor(k) for k in f.keys()}
    else:
        state = load_specific_module(block_idx, weights)
    return config, state


def _targets(spec) -> Dict[str, Tuple[str, Optional[slice]]]:
    qd, kd = spec.num_heads * spec.head_dim, spec.num_kv_heads * spec.head_dim
    if spec.family in ("llama", "mixtral"):
        return {"q_proj": ("wqkv", slice(0, qd)), "k_proj": ("wqkv", slice(qd, qd + kd)), "v_proj": ("wqkv", slice(qd + kd, qd + 2 * kd)),
                "o_proj": ("wo", None), "gate_proj": ("w_gate", None), "up_proj": ("w_up", None), "down_proj": ("w_down", None)}
    return {"query_key_value": ("wqkv", None), "dense": ("wo", None), "dense_h_to_4h": ("w_up", None), "dense_4h_to_h": ("w_down", None)}


def create_lora_adapter(block) -> None:
    if not hasattr(block, "lora_adapters"):
        block.lora_adapters = {}
        block.lora = {}


def add_adapter_to_block(block, block_index: int, adapter_name: str, peft_config: dict, peft_state_dict: Dict[str, torch.Tensor]) -> None:
    create_lora_adapter(block)
    r, alpha = peft_config["r"], peft_config.get("lora_alpha", peft_config["r"])
    scale = alpha / r
    targets = _targets(block.spec)
    pairs: Dict[str, Dict[str, torch.Tensor]] = {}
    for key, tensor in peft_state_dict.items():
        m = re.search(rf"\.{block_index}\.(?:[\w]+\.)*?(\w+)\.lora_([AB])(?:\.\w+)?\.weight$", key)
        if m is None:
            continue
        module, which = m.group(1), m.group(2)
        if module not in targets:
            logger.warning(f"adapter {adapter_name}: module {module} has no LoRA target in {block.spec.family} blocks; skipped")
            continue
        pairs.setdefault(module, {})[which] = tensor
    entry: Dict[str, list] = {}
    dev, dt = next(block.parameters()).device, next(block.parameters()).dtype
    for module, ab in pairs.items():
        if "A" not in ab or "B" not in ab:
            raise ValueError(f"adapter {adapter_name}: incomplete LoRA pair for {module} in block {block_index}")
        name, rows = targets[module]
        weight = getattr(block, name)
        out_rows = weight.shape[0] if rows is None else len(range(*rows.indices(weight.shape[0])))
        A, B = ab["A"], ab["B"]
        if A.dim() != 2 or B.dim() != 2 or A.shape[1] != weight.shape[1] or B.shape[0] != out_rows or A.shape[0] != B.shape[1]:
            raise ValueError(f"adapter {adapter_name}: LoRA pair for {module} in block {block_index} has shapes A {tuple(A.shape)}, B {tuple(B.shape)}; "
                             f"the projection maps {weight.shape[1]} -> {out_rows} features (an adapter trained for another model?)")
        entry.setdefault(name, []).append((A.to(dev, dt), B.to(dev, dt), scale, rows))
    block.lora_adapters[adapter_name] = entry
    logger.debug(f"block {block_index}: loaded adapter {adapter_name} ({sum(len(v) for v in entry.values())} LoRA pairs)")


def merge_lora(weight: torch.Tensor, entries) -> torch.Tensor:
    """``W + sum(scale * B @ A)`` as a new tensor of ``weight``'s dtype. ``entries`` is one value of a block's adapter table:
    ``[(A [r, in], B [rows, r], scale, rows-or-None)]`` where ``rows`` selects the output rows of a fused projection the pair applies to
    (e.g. the q rows of ``wqkv``).  Accumulated in fp32 so that several pairs on one matrix round once."""
    merged = weight.detach().float().clone()
    for A, Bm, scale, rows in entries:
        delta = (Bm.float() @ A.float()) * float(scale)
        if rows is None:
            merged += delta
        else:
            merged[rows] += delta
    return merged.to(weight.dtype)


class MergedAdapterBlock:
    """A read-only view of a block in which the projections an adapter targets are replaced by merged copies ``W + scale * B A``;
    every other attribute (norms, untouched projections, spec, helpers) is the base block's.  The kernels then serve a LoRA request at
    full speed, at the price of one extra copy of the targeted matrices per adapter."""

    def __init__(self, base, adapter_name: str):
        create_lora_adapter(base)
        if adapter_name not in base.lora_adapters:
            raise KeyError(f"Adapter {adapter_name!r} is not loaded on this server (available: {sorted(base.lora_adapters)})")
        object.__setattr__(self, "_base", base)
        object.__setattr__(self, "_merged", {name: merge_lora(getattr(base, name), entries)
                                             for name, entries in base.lora_adapters[adapter_name].items()})
        object.__setattr__(self, "lora", {})  # already folded into the weights

    def __getattr__(self, name: str):
        merged = object.__getattribute__(self, "_merged")
        if name in merged:
            return merged[name]
        return getattr(object.__getattribute__(self, "_base"), name)

    def __setattr__(self, name, value):
        raise AttributeError("MergedAdapterBlock is read-only")

    def _p(self, name: str):
        return getattr(self, name, None)

    @property
    def merged_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in object.__getattribute__(self, "_merged").values())


def set_active_adapter(block, adapter_name: Optional[str]) -> None:
    create_lora_adapter(block)
    if adapter_name in (None, ""):
        block.lora = {}
    elif adapter_name not in block.lora_adapters:
        raise KeyError(f"Adapter {adapter_name!r} is not loaded on this server (available: {sorted(block.lora_adapters)})")
    else:
        block.lora = block.lora_adapters[adapter_name]


@contextlib.contextmanager
def using_adapter(block, adapter_name: Optional[str]):
    prev = getattr(block, "lora", {})
    set_active_adapter(block, adapter_name)
    try:
        yield
    finally:
        block.lora = prev


def estimate_adapter_memory_per_block(block_config, torch_dtype: Optional[torch.dtype], adapters: Sequence[str], **load_peft_kwargs) -> int:
    """Bytes of LoRA parameters one block needs for all requested adapters (reference peft.py:263-283)."""
    total = 0
    for adapter in adapters:
        _, state = load_peft(adapter, block_idx=0, **load_peft_kwargs)
        for t in state.values():
            itemsize = torch.finfo(torch_dtype).bits // 8 if torch_dtype is not None else t.element_size()
            total += t.numel() * itemsize
    return total

"""Model-type registry and the ``AutoDistributed*`` factories (reference: src/petals/utils/auto_config.py:1-99)."""
from __future__ import annotations

import json
import os
from typing import Dict, Optional, Type

from petals_b200.utils.paths import resolve_model_path

_REGISTRY: Dict[str, Dict[str, type]] = {}


def register_model_classes(*, config: type, model: Optional[type] = None, model_for_causal_lm: Optional[type] = None,
                           model_for_speculative: Optional[type] = None,
                           model_for_sequence_classification: Optional[type] = None, block: Optional[type] = None) -> None:
    model_type = config.model_type
    if not model_type:
        raise ValueError("config class must define model_type")
    if model_type in _REGISTRY and _REGISTRY[model_type].get("config") is not config:
        raise ValueError(f"model type {model_type!r} is already registered")
    entry = _REGISTRY.setdefault(model_type, {})
    for key, cls in dict(config=config, model=model, model_for_causal_lm=model_for_causal_lm,
                         model_for_speculative=model_for_speculative,
                         model_for_sequence_classification=model_for_sequence_classification, block=block).items():
        if cls is not None:
            entry[key] = cls


def register_family(package: str, family: str, *, speculative: bool = False) -> Dict[str, type]:
    """Register the classes of ``petals_b200.models.<family>`` by their naming convention (``Distributed<Family>Config``,
    ``...Model``, ``...ForCausalLM``, ``...ForSequenceClassification``, ``Wrapped<Family>Block`` and, for families that have one,
    ``...ForSpeculativeGeneration``) and return them so the package can re-export the names."""
    import importlib

    stem = f"Distributed{family}"
    config_mod, model_mod, block_mod = (importlib.import_module(f"{package}.{name}") for name in ("config", "model", "block"))
    found = {
        "config": getattr(config_mod, f"{stem}Config"),
        "model": getattr(model_mod, f"{stem}Model"),
        "model_for_causal_lm": getattr(model_mod, f"{stem}ForCausalLM"),
        "model_for_sequence_classification": getattr(model_mod, f"{stem}ForSequenceClassification"),
        "block": getattr(block_mod, f"Wrapped{family}Block"),
    }
    if speculative:
        found["model_for_speculative"] = getattr(importlib.import_module(f"{package}.speculative_model"), f"{stem}ForSpeculativeGeneration")
    register_model_classes(**found)
    return {cls.__name__: cls for cls in found.values()}


def get_model_classes(model_type: str) -> Dict[str, type]:
    import petals_b200.models  # noqa: F401  (registers the built-in families)

    if model_type not in _REGISTRY:
        raise ValueError(f"Petals-B200 does not support model type {model_type!r} (known: {sorted(_REGISTRY)})")
    return _REGISTRY[model_type]


def detect_model_type(model_name_or_path: str) -> str:
    path = resolve_model_path(str(model_name_or_path))
    with open(os.path.join(path, "config.json")) as f:
        return json.load(f)["model_type"]


class _AutoDistributedBase:
    _mapping_field: str = ""

    @classmethod
    def from_pretrained(cls, model_name_or_path, *args, **kwargs):
        model_type = detect_model_type(model_name_or_path)
        classes = get_model_classes(model_type)
        if cls._mapping_field not in classes:
            raise ValueError(f"{cls.__name__} is not available for model type {model_type!r}")
        return classes[cls._mapping_field].from_pretrained(model_name_or_path, *args, **kwargs)


class AutoDistributedConfig(_AutoDistributedBase):
    _mapping_field = "config"


class AutoDistributedModel(_AutoDistributedBase):
    _mapping_field = "model"


class AutoDistributedModelForCausalLM(_AutoDistributedBase):
    _mapping_field = "model_for_causal_lm"


class AutoDistributedSpeculativeModel(_AutoDistributedBase):
    _mapping_field = "model_for_speculative"


class AutoDistributedModelForSequenceClassification(_AutoDistributedBase):
    _mapping_field = "model_for_sequence_classification"

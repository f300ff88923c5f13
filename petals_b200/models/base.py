"""Common base of the distributed model configs.

The reference multiply-inherits Hugging Face config classes with its client mix-ins
(src/petals/models/llama/config.py:16). HF internals drifted (installed transformers is 5.x; the
reference pins 4.43), so configs here are self-contained: ``config.json`` of an HF
checkpoint is read directly, every key becomes an attribute, and the client / prompt-tuning / LM-head
options are ordinary attributes with the reference's names and defaults."""
from __future__ import annotations

import copy
import dataclasses
import glob
import json
import os
from typing import Any, Dict, Optional

import torch

from petals_b200.client.config import ClientConfig
from petals_b200.client.lm_head import LMHeadConfig
from petals_b200.client.ptune import PTuneConfig
from petals_b200.constants import DTYPE_MAP
from petals_b200.models.spec import BlockSpec
from petals_b200.utils.paths import resolve_model_path  # noqa: F401  (re-exported)

_EXTRA_DEFAULTS: Dict[str, Any] = {}
for _cls in (ClientConfig, PTuneConfig, LMHeadConfig):
    for _f in dataclasses.fields(_cls):
        _EXTRA_DEFAULTS[_f.name] = _f.default


class DistributedConfig:
    """Attribute bag over an HF ``config.json`` + client options. Subclasses set ``model_type`` & co."""

    model_type: str = ""
    block_prefix: str = ""
    attribute_map: Dict[str, str] = {}
    defaults: Dict[str, Any] = {}

    def __init__(self, **kwargs):
        values = dict(_EXTRA_DEFAULTS)
        values.update(copy.deepcopy(self.defaults))
        values.update(kwargs)
        for k, v in values.items():
            object.__setattr__(self, self.attribute_map.get(k, k), v)
        if isinstance(getattr(self, "torch_dtype", None), str) and self.torch_dtype != "auto":
            self.torch_dtype = DTYPE_MAP.get(self.torch_dtype.replace("torch.", ""), self.torch_dtype)
        if getattr(self, "dtype", None) is not None and getattr(self, "torch_dtype", None) is None:
            self.torch_dtype = DTYPE_MAP.get(str(self.dtype).replace("torch.", ""), None)

    def __getattr__(self, name):  # only called when normal lookup fails
        mapped = type(self).attribute_map.get(name)
        if mapped is not None and mapped in self.__dict__:
            return self.__dict__[mapped]
        raise AttributeError(f"{type(self).__name__} has no attribute {name!r}")

    # ---- HF-compatible surface ----------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, model_name_or_path: str, *args, dht_prefix: Optional[str] = None, **kwargs):
        path = resolve_model_path(str(model_name_or_path))
        with open(os.path.join(path, "config.json")) as f:
            raw = json.load(f)
        kwargs.pop("use_auth_token", None), kwargs.pop("token", None), kwargs.pop("revision", None)
        return_unused = kwargs.pop("return_unused_kwargs", False)
        raw.update(kwargs)
        config = cls(**raw)
        config.name_or_path = str(model_name_or_path)
        config._local_path = path
        if dht_prefix is None:
            dht_prefix = raw.get("dht_prefix") or cls.default_dht_prefix(str(model_name_or_path))
        config.dht_prefix = dht_prefix
        return (config, {}) if return_unused else config

    @classmethod
    def default_dht_prefix(cls, name_or_path: str) -> str:
        base = os.path.basename(os.path.normpath(name_or_path)) if os.path.isdir(name_or_path) else name_or_path.split("/")[-1]
        return base.replace(".", "-") + "-hf"

    def to_dict(self) -> dict:
        out = {}
        for k, v in self.__dict__.items():
            if k.startswith("_"):
                continue
            if isinstance(v, torch.dtype):
                v = str(v).replace("torch.", "")
            try:
                json.dumps(v)
            except TypeError:
                continue
            out[k] = v
        out["model_type"] = self.model_type
        return out

    def save_pretrained(self, path: str) -> None:
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump(self.to_dict(), f, indent=2, sort_keys=True)

    def __repr__(self) -> str:
        return f"{type(self).__name__}({json.dumps(self.to_dict(), sort_keys=True)})"

    # ---- what the engine needs ----------------------------------------------------------------------
    def block_spec(self) -> BlockSpec:
        raise NotImplementedError

    @property
    def num_key_value_groups(self) -> int:
        spec = self.block_spec()
        return spec.num_heads // spec.num_kv_heads

    # name mapping: HF tensor name (without the "<block_prefix>.<i>." part) -> canonical tensor(s)
    @classmethod
    def convert_block_weights(cls, hf: Dict[str, torch.Tensor], spec: BlockSpec) -> Dict[str, torch.Tensor]:
        raise NotImplementedError

    @classmethod
    def export_block_weights(cls, canon: Dict[str, torch.Tensor], spec: BlockSpec) -> Dict[str, torch.Tensor]:
        """Inverse of convert_block_weights (used by the synthetic checkpoint generator)."""
        raise NotImplementedError

    # client-side (non-block) tensors: canonical name -> HF name
    client_weight_names: Dict[str, str] = {}

"""Flatten nested (args, kwargs) into a tensor list + a msgpack-able skeleton, and back
(reference behaviour: src/petals/utils/packaging.py:1-49). The skeleton is what travels on the control
channel; tensors travel by address (same box) or as raw bytes (multi-process CPU transport)."""
from __future__ import annotations

from typing import Any, Dict, List, Tuple

import torch

import re

_TENSOR_TAG = "__T"
_PLACEHOLDER = re.compile(rb"^__T(\d+)$")
_ESCAPE = b"__E"  # user bytes that look like a placeholder (or like an escape) travel with this prefix


def _is_placeholder(x: Any) -> bool:
    """Only *bytes* of the exact form ``__T<index>`` stand for a tensor; text arguments are never reinterpreted."""
    return isinstance(x, bytes) and _PLACEHOLDER.match(x) is not None


def _flatten(obj: Any, tensors: List[torch.Tensor]) -> Any:
    if isinstance(obj, torch.Tensor):
        tensors.append(obj)
        return f"{_TENSOR_TAG}{len(tensors) - 1}".encode()
    if isinstance(obj, (list, tuple)):
        return [_flatten(o, tensors) for o in obj]
    if isinstance(obj, dict):
        return {k: _flatten(v, tensors) for k, v in obj.items()}
    if isinstance(obj, bytes) and (obj.startswith(_ESCAPE) or _PLACEHOLDER.match(obj)):
        return _ESCAPE + obj
    return obj


def _restore(obj: Any, tensors: List[torch.Tensor]) -> Any:
    if _is_placeholder(obj):
        return tensors[int(_PLACEHOLDER.match(obj).group(1))]
    if isinstance(obj, bytes) and obj.startswith(_ESCAPE):
        return obj[len(_ESCAPE):]
    if isinstance(obj, (list, tuple)):
        return [_restore(o, tensors) for o in obj]
    if isinstance(obj, dict):
        return {k: _restore(v, tensors) for k, v in obj.items()}
    return obj


def pack_args_kwargs(*args, **kwargs) -> Tuple[List[torch.Tensor], Any]:
    """Returns (flat tensors, structure) such that unpack_args_kwargs inverts it."""
    tensors: List[torch.Tensor] = []
    structure = (_flatten(list(args), tensors), _flatten(dict(kwargs), tensors))
    return tensors, structure


def unpack_args_kwargs(flat_tensors: List[torch.Tensor], args_structure: Any) -> Tuple[List[Any], Dict[str, Any]]:
    args_s, kwargs_s = args_structure
    return _restore(args_s, flat_tensors), _restore(kwargs_s, flat_tensors)

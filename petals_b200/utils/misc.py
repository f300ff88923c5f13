"""Sentinels of the request schemas and two tiny helpers (reference: src/petals/utils/misc.py:1-29).

Every RPC has a fixed tensor signature (hidden states, prompts, hypo_ids); an argument that is not used still has to occupy
its position, so "absent" is spelled as an empty tensor — ``DUMMY`` — on both sides of the wire."""
from typing import Any, Callable

import torch

DUMMY = torch.empty(0)  # an absent floating-point argument (prompts)
DUMMY_INT64 = torch.empty(0, dtype=torch.int64)  # an absent index argument (hypo_ids)
DUMMY_KEY_PAST = torch.empty((0, 0, 0))  # legacy placeholder of a key cache in BLOOM layout


def is_dummy(value: Any) -> bool:
    """True for ``None`` and for tensors without elements (whatever their shape / dtype)."""
    if value is None:
        return True
    return torch.is_tensor(value) and value.numel() == 0


def get_size_in_bytes(dtype: torch.dtype) -> int:
    """Bytes of one element of ``dtype`` (floats, ints, bool, fp8 and complex alike)."""
    return torch.empty((), dtype=dtype).element_size()


def docstring_from(source: Any) -> Callable[[Any], Any]:
    """Decorator: reuse ``source``'s docstring."""

    def copy_doc(target: Any) -> Any:
        target.__doc__ = getattr(source, "__doc__", None)
        return target

    return copy_doc

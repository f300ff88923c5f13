"""Sentinel tensors and tiny helpers (reference: src/petals/utils/misc.py:1-29).

An *empty* tensor means "argument absent" in every request schema (prompts, hypo_ids, ...)."""
import torch

DUMMY = torch.empty(0)  # absent float tensor
DUMMY_INT64 = torch.empty(0, dtype=torch.int64)  # absent index tensor
DUMMY_KEY_PAST = torch.empty((0, 0, 0))


def is_dummy(tensor) -> bool:
    return tensor is None or (isinstance(tensor, torch.Tensor) and tensor.numel() == 0)


_BITS = {torch.bool: 8, torch.float8_e4m3fn: 8, torch.float8_e5m2: 8}


def get_size_in_bytes(dtype: torch.dtype) -> int:
    if dtype in _BITS:
        return _BITS[dtype] // 8
    info = torch.finfo(dtype) if dtype.is_floating_point else torch.iinfo(dtype)
    return info.bits * (1 + dtype.is_complex) // 8


def docstring_from(source):
    def deco(dest):
        dest.__doc__ = source.__doc__
        return dest

    return deco

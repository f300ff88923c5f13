"""safetensors read/write without the Rust package on the hot path.

Reading goes through the native mmap reader (csrc/runtime/safetensors_reader.cpp): tensors are exposed
zero-copy as torch views over the mapping, or copied with several threads into pinned staging memory
for the H2D upload. Writing (only needed by the synthetic checkpoint generator and tests) is a few lines
of Python since the format is an 8-byte length + JSON header + raw little-endian data."""
from __future__ import annotations

import ctypes as C
import json
import struct
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch

from petals_b200.ops import native

_DTYPES = {
    "F64": torch.float64, "F32": torch.float32, "F16": torch.float16, "BF16": torch.bfloat16,
    "I64": torch.int64, "I32": torch.int32, "I16": torch.int16, "I8": torch.int8, "U8": torch.uint8, "BOOL": torch.bool,
    "U16": torch.uint16, "U32": torch.uint32, "U64": torch.uint64,
    "F8_E4M3": torch.float8_e4m3fn, "F8_E5M2": torch.float8_e5m2,
}
_NAMES = {v: k for k, v in _DTYPES.items()}


class SafetensorsFile:
    """Read-only view of one ``.safetensors`` file. Use as a context manager."""

    def __init__(self, path: str):
        self._rt = native.rt()
        self._h = self._rt.pb_st_open(str(path).encode())
        if not self._h:
            raise IOError(f"cannot read {path}: {self._rt.pb_st_error().decode()}")
        self.path = str(path)
        self._index: Dict[str, Tuple[int, str, Tuple[int, ...], int, int]] = {}
        name, dtype = C.create_string_buffer(1024), C.create_string_buffer(32)
        shape = (C.c_int64 * 8)()
        off, nbytes = C.c_int64(), C.c_int64()
        for i in range(self._rt.pb_st_num_tensors(self._h)):
            nd = self._rt.pb_st_tensor_info(self._h, i, name, 1024, dtype, 32, shape, C.byref(off), C.byref(nbytes))
            key, kind, dims = name.value.decode(), dtype.value.decode(), tuple(shape[:max(nd, 0)])
            # the header is untrusted input: a record whose shape does not account for exactly its byte range would make the copy in
            # get_tensor() overrun the destination (or leave most of it uninitialised)
            if nd < 0:
                self.close()
                raise IOError(f"cannot read {path}: malformed record for tensor {key!r}")
            if kind not in _DTYPES:  # some other tool's dtype: the file stays usable, only this tensor is refused when asked for
                self._index[key] = (i, kind, dims, off.value, nbytes.value)
                continue
            numel = 1
            for d in dims:
                numel *= d
            expected = numel * torch.empty((), dtype=_DTYPES[kind]).element_size()
            if any(d < 0 for d in dims) or expected != nbytes.value:
                self.close()
                raise IOError(f"cannot read {path}: tensor {key!r} of shape {dims} ({kind}) needs {expected} bytes but its data range has {nbytes.value}")
            self._index[key] = (i, kind, dims, off.value, nbytes.value)
        self._base = self._rt.pb_st_data(self._h)

    def keys(self) -> List[str]:
        return list(self._index)

    def __contains__(self, name: str) -> bool:
        return name in self._index

    def _dtype_of(self, name: str, kind: str) -> torch.dtype:
        if kind not in _DTYPES:
            raise IOError(f"tensor {name!r} in {self.path} has an unsupported dtype {kind!r}")
        return _DTYPES[kind]

    def info(self, name: str) -> Tuple[torch.dtype, Tuple[int, ...]]:
        _, dt, shape, _, _ = self._index[name]
        return self._dtype_of(name, dt), shape

    def get_tensor(self, name: str, pinned: bool = False, threads: int = 8) -> torch.Tensor:
        """A private copy of the tensor (optionally in pinned memory, filled with a multi-threaded memcpy)."""
        idx, dt, shape, _, nbytes = self._index[name]
        dtype = self._dtype_of(name, dt)
        out = torch.empty(shape, dtype=dtype, pin_memory=pinned and torch.cuda.is_available())
        if nbytes:
            rc = self._rt.pb_st_read(self._h, idx, out.data_ptr(), nbytes, threads)
            if rc != 0:
                raise IOError(f"read of {name} failed ({rc})")
        return out

    def close(self) -> None:
        if self._h:
            self._rt.pb_st_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def load_file(path: str, keys: Optional[Iterable[str]] = None) -> Dict[str, torch.Tensor]:
    with SafetensorsFile(path) as f:
        wanted = f.keys() if keys is None else [k for k in keys if k in f]
        return {k: f.get_tensor(k) for k in wanted}


def save_file(tensors: Dict[str, torch.Tensor], path: str, metadata: Optional[Dict[str, str]] = None) -> None:
    header: Dict[str, object] = {}
    if metadata:
        header["__metadata__"] = metadata
    offset = 0
    blobs = []
    for name in sorted(tensors):
        t = tensors[name].detach().cpu().contiguous()
        raw = t.view(torch.uint8).numpy().tobytes() if t.numel() else b""
        header[name] = {"dtype": _NAMES[t.dtype], "shape": list(t.shape), "data_offsets": [offset, offset + len(raw)]}
        offset += len(raw)
        blobs.append(raw)
    js = json.dumps(header, separators=(",", ":")).encode()
    js += b" " * ((8 - len(js) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(js)))
        f.write(js)
        for raw in blobs:
            f.write(raw)

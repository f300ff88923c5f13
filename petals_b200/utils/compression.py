"""Activation codecs for the socket transport.

The reference lets a server pick a wire codec for hidden states (``--compression``, reference
src/petals/cli/run_server.py:62,196-197) and lets a client ask for one per request
(``metadata["output_compression"]``, reference src/petals/server/handler.py:421-432); the codecs themselves live
in hivemind (NONE / FLOAT16 / MEANSTD_16BIT / QUANTILE_8BIT / UNIFORM_8BIT / BLOCKWISE_8BIT).

Here the GPU data plane never serialises anything (stage hops are NVLink stores in bf16), so the codecs only
matter for the unix-socket control transport (``parallel/transport.py``) that carries tensors between processes
that do not share a GPU fabric — exactly the case the reference built them for (slow links).  They are written
from the definitions, with torch ops only:

* ``FLOAT16``        – clamp to the fp16 range, cast; 2 bytes/value.
* ``MEANSTD_16BIT``  – per-row (last dim) standardisation, payload in fp16, fp32 mean and std per row.
* ``UNIFORM_8BIT``   – 256 equal-width buckets over mean ± 6σ; codebook = mean of the values in each bucket.
* ``QUANTILE_8BIT``  – 256 buckets with (sampled) quantile borders; codebook = bucket means.
* ``BLOCKWISE_8BIT`` – blocks of 4096 values scaled by their absmax, signed 8-bit code on a quadratic grid
                        (dense near zero, where hidden states live).
* ``MXFP8``          – not a hivemind codec: the block-scaled FP8 format of ``ops/quant.py`` (E4M3 payload, one UE8M0 power-of-two
                        scale per 32 values along the last dimension) — the format Blackwell tensor cores consume and the fp8 stages
                        quantise their activations to anyway; 1.03 bytes/value, no code book, error bounded per 32-value group.

``encode`` returns ``(meta, [flat uint8 tensors])`` and ``decode`` inverts it; both are exact inverses for ``NONE`` and
lossy within the bounds tested in tests/test_compression.py otherwise.
"""
from __future__ import annotations

from enum import IntEnum
from typing import Any, Dict, List, Sequence, Tuple, Union

import torch


class CompressionType(IntEnum):
    # numbering follows the wire enum the reference's clients send in ``output_compression``
    NONE = 0
    MEANSTD_16BIT = 1
    FLOAT16 = 2
    QUANTILE_8BIT = 3
    UNIFORM_8BIT = 4
    BLOCKWISE_8BIT = 5
    MXFP8 = 6  # beyond the reference's enum: block-scaled FP8 (ops/quant.py)


CodecSpec = Union[None, str, int, CompressionType]

_DTYPE_NAMES = {torch.float32: "f32", torch.float16: "f16", torch.bfloat16: "bf16", torch.float64: "f64"}
_DTYPES = {v: k for k, v in _DTYPE_NAMES.items()}
_FP16_MAX = 65504.0
_BLOCK = 4096
_N_BUCKETS = 256
_QUANTILE_SAMPLE = 1 << 17
_UNIFORM_SIGMAS = 6.0


def parse_compression(spec: CodecSpec) -> CompressionType:
    if spec is None:
        return CompressionType.NONE
    if isinstance(spec, CompressionType):
        return spec
    if isinstance(spec, int):
        return CompressionType(spec)
    try:
        return CompressionType[str(spec).upper()]
    except KeyError:
        raise ValueError(f"unknown compression {spec!r}; choose from {[c.name for c in CompressionType]}") from None


_EMPTY = torch.empty(0, dtype=torch.uint8)


def _raw(t: torch.Tensor) -> torch.Tensor:
    """The tensor's bytes as a flat uint8 CPU view (no copy for contiguous CPU tensors): the transport hands the storage
    pointer to a scatter-gather ``sendmsg`` instead of building a ``bytes`` object."""
    if t.numel() == 0:
        return _EMPTY
    return t.detach().to("cpu").contiguous().reshape(-1).view(torch.uint8)


class WireFormatError(ValueError):
    """A received tensor descriptor does not match its payload (the shape comes from an untrusted header)."""


def _from_raw(raw, dtype: torch.dtype, shape: Sequence[int]) -> torch.Tensor:
    """Reinterpret received bytes as ``dtype[shape]``. The byte count must match the announced shape EXACTLY: a header that
    announces a shape with no (or too few) bytes must never yield uninitialised memory or a huge allocation."""
    if isinstance(raw, (bytes, bytearray, memoryview)):
        raw = torch.frombuffer(bytearray(raw), dtype=torch.uint8) if len(raw) else _EMPTY
    shape = [int(s) for s in shape]
    if any(s < 0 for s in shape):
        raise WireFormatError(f"negative dimension in {shape}")
    numel = 1
    for s in shape:
        numel *= s
    itemsize = torch.empty(0, dtype=dtype).element_size()
    if raw.numel() != numel * itemsize:
        raise WireFormatError(f"payload of {raw.numel()} bytes does not match {dtype} {shape} ({numel * itemsize} bytes)")
    if numel == 0:
        return torch.empty(shape, dtype=dtype)
    return raw.view(dtype).reshape(shape)


def _codebook_from_buckets(x: torch.Tensor, idx: torch.Tensor, fallback: torch.Tensor) -> torch.Tensor:
    """Mean of the values that fell into each bucket; empty buckets keep ``fallback`` (their centre)."""
    sums = torch.zeros(_N_BUCKETS, dtype=torch.float32).scatter_add_(0, idx, x)
    counts = torch.zeros(_N_BUCKETS, dtype=torch.float32).scatter_add_(0, idx, torch.ones_like(x))
    return torch.where(counts > 0, sums / counts.clamp(min=1), fallback)


def _bucket_encode(x: torch.Tensor, borders: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """``borders`` are the 255 inner edges; returns (uint8 indices, fp32 codebook[256])."""
    idx = torch.bucketize(x, borders).clamp_(0, _N_BUCKETS - 1)
    edges = torch.cat([borders[:1], borders, borders[-1:]])
    centres = 0.5 * (edges[:-1] + edges[1:])
    return idx.to(torch.uint8), _codebook_from_buckets(x, idx, centres)


# signed quadratic grid for BLOCKWISE_8BIT: code c in [-127, 127] -> sign(c) * (c / 127)^2
_BLOCK_GRID = None


def _block_grid() -> torch.Tensor:
    global _BLOCK_GRID
    if _BLOCK_GRID is None:
        c = torch.arange(-127, 128, dtype=torch.float32) / 127.0
        _BLOCK_GRID = torch.sign(c) * c * c
    return _BLOCK_GRID


def encode(t: torch.Tensor, compression: CodecSpec = None) -> Tuple[Dict[str, Any], List[torch.Tensor]]:
    """-> (meta, blobs); every blob is a flat uint8 CPU tensor. Non-float and empty tensors always travel as ``NONE``."""
    codec = parse_compression(compression)
    t = t.detach()
    if codec != CompressionType.NONE and (t.dtype not in _DTYPE_NAMES or t.numel() == 0):
        codec = CompressionType.NONE
    if codec == CompressionType.NONE:
        return {"codec": "NONE"}, [_raw(t)]

    meta: Dict[str, Any] = {"codec": codec.name, "orig_dtype": _DTYPE_NAMES[t.dtype]}
    x = t.to("cpu", torch.float32).contiguous()
    if codec == CompressionType.FLOAT16:
        return meta, [_raw(x.clamp(-_FP16_MAX, _FP16_MAX).to(torch.float16))]

    if codec == CompressionType.MEANSTD_16BIT:
        rows = x.reshape(-1, x.shape[-1]) if x.dim() else x.reshape(1, 1)
        mean = rows.mean(dim=-1, keepdim=True)
        std = rows.std(dim=-1, keepdim=True, unbiased=False).clamp_(min=1e-6) if rows.shape[-1] > 1 else torch.ones_like(mean)
        payload = ((rows - mean) / std).clamp_(-_FP16_MAX, _FP16_MAX).to(torch.float16)
        meta["rows"] = rows.shape[0]
        return meta, [_raw(payload), _raw(mean), _raw(std)]

    flat = x.reshape(-1)
    if codec == CompressionType.UNIFORM_8BIT:
        centre = flat.mean()
        half = _UNIFORM_SIGMAS * (flat.std(unbiased=False) if flat.numel() > 1 else flat.abs().max()).clamp(min=1e-12)
        borders = torch.linspace(float(centre - half), float(centre + half), _N_BUCKETS + 1)[1:-1]
        idx, book = _bucket_encode(flat, borders)
        return meta, [_raw(idx), _raw(book)]

    if codec == CompressionType.QUANTILE_8BIT:
        sample = flat
        if flat.numel() > _QUANTILE_SAMPLE:  # deterministic strided sample keeps the codec reproducible
            step = flat.numel() // _QUANTILE_SAMPLE
            sample = flat[::step][:_QUANTILE_SAMPLE]
        qs = torch.linspace(0, 1, _N_BUCKETS + 1)[1:-1]
        borders = torch.quantile(sample, qs)
        idx, book = _bucket_encode(flat, borders)
        return meta, [_raw(idx), _raw(book)]

    if codec == CompressionType.MXFP8:
        from petals_b200.ops.quant import BLOCK, quantize_mxfp8

        n = flat.numel()
        rows = torch.nn.functional.pad(flat, (0, (-n) % BLOCK)).reshape(-1, BLOCK)  # groups of 32 consecutive values (rows of hidden states are multiples of 32)
        q, e = quantize_mxfp8(rows)
        meta["n"] = n
        return meta, [_raw(q.view(torch.uint8).reshape(-1)[:n]), _raw(e.reshape(-1))]

    if codec == CompressionType.BLOCKWISE_8BIT:
        n = flat.numel()
        pad = (-n) % _BLOCK
        blocks = torch.nn.functional.pad(flat, (0, pad)).reshape(-1, _BLOCK)
        absmax = blocks.abs().amax(dim=-1, keepdim=True).clamp_(min=1e-30)
        y = blocks / absmax  # in [-1, 1]
        code = torch.round(torch.sign(y) * torch.sqrt(y.abs()) * 127.0).to(torch.int16) + 127  # 0..254
        meta["n"] = n
        return meta, [_raw(code.to(torch.uint8).reshape(-1)[:n]), _raw(absmax.reshape(-1))]

    raise AssertionError(codec)


def decode(meta: Dict[str, Any], blobs: Sequence[torch.Tensor], dtype: torch.dtype, shape: Sequence[int]) -> torch.Tensor:
    codec = parse_compression(meta.get("codec", "NONE"))
    if codec == CompressionType.NONE:
        return _from_raw(blobs[0], dtype, shape)
    numel = 1
    for s in shape:
        numel *= int(s)

    if codec == CompressionType.FLOAT16:
        return _from_raw(blobs[0], torch.float16, shape).to(dtype)

    if codec == CompressionType.MEANSTD_16BIT:
        rows = int(meta["rows"])
        if rows <= 0 or numel % rows:
            raise WireFormatError(f"MEANSTD_16BIT: {rows} rows do not divide {numel} elements")
        payload = _from_raw(blobs[0], torch.float16, [rows, numel // rows]).float()
        mean = _from_raw(blobs[1], torch.float32, [rows, 1])
        std = _from_raw(blobs[2], torch.float32, [rows, 1])
        return (payload * std + mean).reshape(list(shape)).to(dtype)

    if codec in (CompressionType.UNIFORM_8BIT, CompressionType.QUANTILE_8BIT):
        idx = _from_raw(blobs[0], torch.uint8, [numel]).long()
        book = _from_raw(blobs[1], torch.float32, [_N_BUCKETS])
        return book[idx].reshape(list(shape)).to(dtype)

    if codec == CompressionType.MXFP8:
        from petals_b200.ops.quant import BLOCK, dequantize_mxfp8

        n = int(meta["n"])
        if n != numel:
            raise WireFormatError(f"MXFP8: {n} values announced for {numel} elements")
        groups = (n + BLOCK - 1) // BLOCK
        q = torch.zeros(groups * BLOCK, dtype=torch.uint8)
        q[:n] = _from_raw(blobs[0], torch.uint8, [n])
        e = _from_raw(blobs[1], torch.uint8, [groups, 1])
        vals = dequantize_mxfp8(q.view(torch.float8_e4m3fn).view(groups, BLOCK), e, torch.float32).reshape(-1)[:n]
        return vals.reshape(list(shape)).to(dtype)

    if codec == CompressionType.BLOCKWISE_8BIT:
        n = int(meta["n"])
        if n != numel:
            raise WireFormatError(f"BLOCKWISE_8BIT: {n} codes announced for {numel} elements")
        code = _from_raw(blobs[0], torch.uint8, [n]).long()
        absmax = _from_raw(blobs[1], torch.float32, [(n + _BLOCK - 1) // _BLOCK])
        vals = _block_grid()[code] * absmax.repeat_interleave(_BLOCK)[:n]
        return vals.reshape(list(shape)).to(dtype)

    raise AssertionError(codec)


def roundtrip(t: torch.Tensor, compression: CodecSpec) -> torch.Tensor:
    """encode + decode in one call (what the peer would see)."""
    meta, blobs = encode(t, compression)
    return decode(meta, blobs, t.dtype, t.shape)


def compressed_nbytes(t: torch.Tensor, compression: CodecSpec) -> int:
    return sum(b.numel() for b in encode(t, compression)[1])


def normalize_output_compression(spec: Any, n_outputs: int, default: CodecSpec = None) -> List[CompressionType]:
    """``metadata["output_compression"]``: one codec per returned tensor (reference handler.py:421-427 insists on
    a list/tuple of valid enum values with the right length)."""
    if spec is None:
        return [parse_compression(default)] * n_outputs
    if not isinstance(spec, (list, tuple)):
        raise ValueError("output_compression must be a list or a tuple")
    if len(spec) != n_outputs:
        raise ValueError(f"output_compression should have {n_outputs} elements, got {len(spec)}")
    return [parse_compression(s) for s in spec]

"""Block-scaled FP8 (MXFP8: E4M3 payload, one power-of-two UE8M0 scale per 32 values along K).

This replaces the reference's bitsandbytes INT8 / NF4 weight quantisation
(src/petals/utils/convert_block.py:76-115), which has no sm_100 path: Blackwell tensor
cores consume exactly this format (``tcgen05.mma kind::mxf8f6f4``). The functions below are the format
definition (quantise / dequantise in PyTorch); the decode-path kernel that streams the 1-byte weights is
in csrc/linear_decode_fp8.cu."""
from __future__ import annotations

from typing import Tuple

import torch

BLOCK = 32
E4M3_MAX = 448.0


def quantize_mxfp8(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """w [N, K] (K % 32 == 0) -> (payload float8_e4m3fn [N, K], scale exponents uint8 [N, K/32])."""
    N, K = w.shape
    if K % BLOCK:
        raise ValueError(f"K={K} must be a multiple of {BLOCK}")
    wf = w.float().view(N, K // BLOCK, BLOCK)
    amax = wf.abs().amax(-1).clamp_min(1e-30)
    # smallest power of two such that amax / 2^e <= E4M3_MAX
    e = torch.ceil(torch.log2(amax / E4M3_MAX)).clamp(-127, 127)
    scale = torch.exp2(e)
    q = (wf / scale[..., None]).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn)
    return q.view(N, K), (e + 127).to(torch.uint8)


def dequantize_mxfp8(q: torch.Tensor, e: torch.Tensor, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    N, K = q.shape
    scale = torch.exp2(e.float() - 127.0)
    return (q.float().view(N, K // BLOCK, BLOCK) * scale[..., None]).view(N, K).to(dtype)


def pack_scales(e: torch.Tensor) -> torch.Tensor:
    """Scale exponents [rows, K/32] -> the block layout the tensor cores read (csrc/gemm_mxfp8.cu): ``[K/128][ceil(rows/128)][512]``
    bytes, the scale of (row r, K slice c) of a block at ``(r % 32) * 16 + (r // 32) * 4 + c``. Rows are padded with exponent 0."""
    rows, kb32 = e.shape
    if kb32 % 4:
        raise ValueError(f"K={kb32 * BLOCK} must be a multiple of 128")
    blocks = (rows + 127) // 128
    padded = torch.zeros(blocks * 128, kb32, dtype=torch.uint8, device=e.device)
    padded[:rows] = e
    # [block, r1 (4), r0 (32), kb128, c (4)] -> [kb128, block, r0, r1, c]
    return padded.view(blocks, 4, 32, kb32 // 4, 4).permute(3, 0, 2, 1, 4).contiguous().view(-1)


def unpack_scales(packed: torch.Tensor, rows: int, K: int) -> torch.Tensor:
    """Inverse of :func:`pack_scales` (tests)."""
    blocks, kb128 = (rows + 127) // 128, K // 128
    e = packed.view(kb128, blocks, 32, 4, 4).permute(1, 3, 2, 0, 4).contiguous().view(blocks * 128, kb128 * 4)
    return e[:rows]


def fake_quantize_mxfp8(w: torch.Tensor) -> torch.Tensor:
    """Round-trip through MXFP8 (what an oracle block must hold to match the fp8 engine bit-for-bit in weights)."""
    if w.dim() == 3:
        return torch.stack([fake_quantize_mxfp8(x) for x in w], 0)
    q, e = quantize_mxfp8(w)
    return dequantize_mxfp8(q, e, w.dtype)

"""Tensor parallelism for *every* block family, written against :class:`BlockSpec` (the generic path).

The reference shards any block with the ``tensor_parallel`` library: ``--tensor_parallel_devices d0 d1 ...`` wraps each
loaded block so that every device holds a slice of the heads and of the FFN and an all-reduce follows the two
row-parallel linears (src/petals/utils/convert_block.py:118-135, src/petals/server/backend.py:67-98; its CI runs it on
``cpu cpu`` with BLOOM, .github/workflows/run-tests.yaml:81-83).  This repo has two implementations of the same split:

* ``parallel/tensor_parallel.py`` — the NVLink engine (one process per GPU, all-reduce fused into the GEMV epilogue /
  next prologue) for the layouts its kernels cover (``tp_supported``: Llama-style blocks);
* this module — the same Megatron split for **all** families (BLOOM's and Falcon's fused interleaved QKV, ALiBi head
  slopes, multi-query attention, parallel-attention blocks, biases, Mixtral experts), executed by the PyTorch oracle
  blocks on whatever devices are named.  It is what ``--tensor_parallel_devices`` uses on CPU devices and for the
  families the engine does not shard yet, and it is the numerical reference the engine's sharding is tested against.

Split rules (``world`` ranks; ``hq``/``hkv`` = local query / kv heads):
  column-parallel: ``wqkv``/``bqkv`` by kv group (query heads follow their kv head; multi-query attention replicates the single
  kv head and splits the query heads), ``w_gate``/``w_up``/``b_up`` and every expert's gate/up by FFN rows;
  row-parallel: ``wo``, ``w_down`` and every expert's down by input columns; their biases live on rank 0 only so that the
  sum of the partial outputs adds them once; norms and the MoE router are replicated.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from petals_b200.models.block_oracle import GenericBlock
from petals_b200.models.spec import BlockSpec


def tp_shardable(spec: BlockSpec, world: int) -> bool:
    """Can this block be split ``world`` ways by heads and FFN columns?"""
    if world < 1:
        return False
    kv_ok = spec.num_kv_heads % world == 0 or (spec.num_kv_heads == 1 and not spec.qkv_interleaved)
    return kv_ok and spec.num_heads % world == 0 and spec.intermediate_size % world == 0


def shard_spec(spec: BlockSpec, rank: int, world: int) -> BlockSpec:
    """What one rank executes: its share of the heads (with the global ALiBi slope indices) and FFN columns."""
    hq = spec.num_heads // world
    hkv = spec.num_kv_heads // world if spec.num_kv_heads % world == 0 else spec.num_kv_heads  # multi-query: replicated
    return dataclasses.replace(spec, num_heads=hq, num_kv_heads=hkv, intermediate_size=spec.intermediate_size // world,
                               alibi_total_heads=spec.alibi_total_heads or spec.num_heads,
                               alibi_head_offset=spec.alibi_head_offset + rank * hq)


def _qkv_rows(spec: BlockSpec, rank: int, world: int) -> torch.Tensor:
    """Row indices of the fused QKV projection that belong to ``rank``, in the order of the shard's own layout."""
    D, nh, nkv = spec.head_dim, spec.num_heads, spec.num_kv_heads
    hq = nh // world
    rows = torch.arange(spec.qkv_dim)
    if spec.qkv_interleaved:  # [kv group][G query heads, k, v][D]: whole groups move together
        hkv = nkv // world
        per_group = (spec.group_size + 2) * D
        return rows[rank * hkv * per_group: (rank + 1) * hkv * per_group]
    q = rows[rank * hq * D: (rank + 1) * hq * D]
    if nkv % world == 0:
        hkv = nkv // world
        k = rows[nh * D + rank * hkv * D: nh * D + (rank + 1) * hkv * D]
        v = rows[(nh + nkv) * D + rank * hkv * D: (nh + nkv) * D + (rank + 1) * hkv * D]
    else:  # multi-query: every rank keeps the single k / v head
        k = rows[nh * D: (nh + nkv) * D]
        v = rows[(nh + nkv) * D:]
    return torch.cat([q, k, v])


def shard_tensors(block: GenericBlock, spec: BlockSpec, rank: int, world: int) -> Dict[str, torch.Tensor]:
    """canonical name -> this rank's slice (views / copies on the block's device)."""
    if not tp_shardable(spec, world):
        raise ValueError(f"cannot split {spec.num_heads} heads / {spec.num_kv_heads} kv heads / FFN {spec.intermediate_size} {world} ways")
    D, I = spec.head_dim, spec.intermediate_size // world
    hq = spec.num_heads // world
    qkv_rows = _qkv_rows(spec, rank, world).to(block.wqkv.device)
    cols = slice(rank * I, (rank + 1) * I)
    out: Dict[str, torch.Tensor] = {}
    for name in spec.param_shapes():
        t = getattr(block, name)
        if name == "wqkv" or name == "bqkv":
            out[name] = t.index_select(0, qkv_rows)
        elif name == "wo":
            out[name] = t[:, rank * hq * D: (rank + 1) * hq * D]
        elif name in ("w_gate", "w_up", "b_up"):
            out[name] = t[cols]
        elif name == "w_down":
            out[name] = t[:, cols]
        elif name in ("we_gate", "we_up"):
            out[name] = t[:, cols]
        elif name == "we_down":
            out[name] = t[:, :, cols]
        elif name in ("bo", "b_down"):  # added once by the sum over ranks
            out[name] = t if rank == 0 else torch.zeros_like(t)
        else:  # norms, router: replicated
            out[name] = t
    return out


def make_shards(block: GenericBlock, spec: BlockSpec, devices: Sequence) -> List[GenericBlock]:
    """One oracle block per device holding exactly its slice of ``block`` (frozen like the source)."""
    world = len(devices)
    shards = []
    for rank, device in enumerate(devices):
        b = GenericBlock(shard_spec(spec, rank, world), dtype=block.wqkv.dtype, device=device)
        with torch.no_grad():
            for name, t in shard_tensors(block, spec, rank, world).items():
                getattr(b, name).copy_(t)
        b.requires_grad_(False)
        shards.append(b)
    return shards


class TensorParallelBlock(nn.Module):
    """A block split over ``devices``; same call surface as :class:`GenericBlock` (``forward_cached`` / ``forward``).

    The residual stream, the norms and (for MoE) the routing decision are replicated; each shard contributes a partial
    attention output and a partial MLP output, summed on the primary device where an engine would all-reduce.  The KV
    cache keeps the dense layout ``[B, L, Hkv, D]`` of the primary device: a shard that lives there reads and writes
    its head slice in place; a shard on another device works on a copy of its slice and returns the new positions."""

    def __init__(self, block: GenericBlock, spec: BlockSpec, devices: Sequence):
        super().__init__()
        self.spec = spec
        self.devices = tuple(torch.device(d) for d in devices)
        self.shards = nn.ModuleList(make_shards(block, spec, self.devices))
        self.kv_split = spec.num_kv_heads % len(self.devices) == 0
        self.lora: dict = {}  # adapters are not sharded (the server refuses --adapters together with tensor parallelism)

    @property
    def world(self) -> int:
        return len(self.devices)

    def _kv_slice(self, cache: Optional[torch.Tensor], rank: int) -> Optional[torch.Tensor]:
        if cache is None or not self.kv_split:
            return cache
        hkv = self.spec.num_kv_heads // self.world
        return cache[:, :, rank * hkv: (rank + 1) * hkv]

    def _attention(self, x: torch.Tensor, k_cache, v_cache, pos: int) -> torch.Tensor:
        total = None
        T = x.shape[1]
        for rank, shard in enumerate(self.shards):
            dev = self.devices[rank]
            k, v = self._kv_slice(k_cache, rank), self._kv_slice(v_cache, rank)
            if k is not None and k.device != dev:  # remote shard: work on a copy, bring the new positions back
                kr, vr = k.to(dev), v.to(dev)
                part = shard.attention(x.to(dev), kr, vr, pos)
                k[:, pos:pos + T], v[:, pos:pos + T] = kr[:, pos:pos + T].to(k.device), vr[:, pos:pos + T].to(v.device)
            else:
                part = shard.attention(x.to(dev), k, v, pos)
            part = part.to(x.device)
            total = part if total is None else total + part
        return total

    def _mlp(self, x: torch.Tensor) -> torch.Tensor:
        total = None
        for rank, shard in enumerate(self.shards):
            part = shard.mlp(x.to(self.devices[rank])).to(x.device)
            total = part if total is None else total + part
        return total

    def forward_cached(self, hidden: torch.Tensor, k_cache: Optional[torch.Tensor], v_cache: Optional[torch.Tensor], pos: int = 0) -> torch.Tensor:
        s, first = self.spec, self.shards[0]
        h0 = hidden.to(self.devices[0])
        if s.parallel_attn:
            a_in = first._norm(h0, "ln1")
            m_in = first._norm(h0, "ln2") if s.dual_ln else a_in
            out = h0 + self._attention(a_in, k_cache, v_cache, pos) + self._mlp(m_in)
        else:
            ln1 = first._norm(h0, "ln1")
            res = ln1 if s.post_ln_residual else h0
            h = res + self._attention(ln1, k_cache, v_cache, pos)
            ln2 = first._norm(h, "ln2")
            res = ln2 if s.post_ln_residual else h
            out = res + self._mlp(ln2)
        return out.to(hidden.device)

    def forward(self, hidden_states: torch.Tensor, layer_past=None, use_cache: bool = False, **_):
        B, T, _ = hidden_states.shape
        s = self.spec
        if layer_past is None and not use_cache:
            return (self.forward_cached(hidden_states, None, None, 0),)
        P = 0 if layer_past is None else layer_past[0].shape[1]
        k = torch.zeros(B, P + T, s.num_kv_heads, s.head_dim, dtype=hidden_states.dtype, device=hidden_states.device)
        v = torch.zeros_like(k)
        if layer_past is not None:
            k[:, :P], v[:, :P] = layer_past[0], layer_past[1]
        out = self.forward_cached(hidden_states, k, v, P)
        return (out, (k, v)) if use_cache else (out,)

    def extra_repr(self) -> str:
        return f"family={self.spec.family}, world={self.world}, devices={[str(d) for d in self.devices]}"


def make_tensor_parallel(block: GenericBlock, spec: BlockSpec, devices: Sequence) -> nn.Module:
    """The reference's ``make_tensor_parallel`` (convert_block.py:118-135): identity for one device."""
    devices = list(devices)
    if len(devices) <= 1:
        return block
    return TensorParallelBlock(block, spec, devices)


# =====================================================================================================================
# One shard per process: the same split with collectives (what the ranks of a tensor-parallel worker group execute for
# the paths that have no fused kernels — training forward/backward through a TP stage).
# =====================================================================================================================
import torch.distributed as dist  # noqa: E402


class _CopyToTP(torch.autograd.Function):
    """Megatron's ``f``: identity going in, all-reduce of the gradient coming back (the replicated activation feeds
    every rank's sharded sub-layer, so its gradient is the sum of the ranks' contributions)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g, group=ctx.group)
        return g, None


class _ReduceFromTP(torch.autograd.Function):
    """Megatron's ``g``: all-reduce of the partial outputs going out, identity for the gradient."""

    @staticmethod
    def forward(ctx, x, group):
        y = x.contiguous().clone()
        dist.all_reduce(y, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


class ShardedBlock(nn.Module):
    """This rank's shard of a block + the process group it all-reduces over; call surface of :class:`GenericBlock`.

    Every rank must make the same calls with the same (replicated) inputs.  The KV caches passed in hold this rank's kv
    heads only (``shard.spec.num_kv_heads``)."""

    def __init__(self, shard: GenericBlock, spec: BlockSpec, group=None):
        super().__init__()
        self.shard, self.spec, self.group = shard, spec, group
        self.lora: dict = {}

    @classmethod
    def from_block(cls, block: GenericBlock, spec: BlockSpec, group=None, device=None) -> "ShardedBlock":
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        shard = GenericBlock(shard_spec(spec, rank, world), dtype=block.wqkv.dtype, device=device or block.wqkv.device)
        with torch.no_grad():
            for name, t in shard_tensors(block, spec, rank, world).items():
                getattr(shard, name).copy_(t)
        shard.requires_grad_(False)
        return cls(shard, spec, group)

    @classmethod
    def from_tensors(cls, tensors: Dict[str, torch.Tensor], spec: BlockSpec, rank: int, world: int, group=None) -> "ShardedBlock":
        """Wrap shard tensors that already live on this rank's device (no copy: the parameters alias them)."""
        ls = shard_spec(spec, rank, world)
        any_t = next(iter(tensors.values()))
        shard = GenericBlock(ls, dtype=any_t.dtype, device="meta")
        for name in ls.param_shapes():
            t = tensors.get(name)
            if t is None:
                raise KeyError(f"shard tensor {name!r} is missing")
            setattr(shard, name, nn.Parameter(t, requires_grad=False))
        return cls(shard, spec, group)

    def _sub(self, fn, x: torch.Tensor, *args) -> torch.Tensor:
        return _ReduceFromTP.apply(fn(_CopyToTP.apply(x, self.group), *args), self.group)

    def forward_cached(self, hidden: torch.Tensor, k_cache: Optional[torch.Tensor], v_cache: Optional[torch.Tensor], pos: int = 0) -> torch.Tensor:
        s, b = self.spec, self.shard
        if s.parallel_attn:
            a_in = b._norm(hidden, "ln1")
            m_in = b._norm(hidden, "ln2") if s.dual_ln else a_in
            return hidden + self._sub(b.attention, a_in, k_cache, v_cache, pos) + self._sub(b.mlp, m_in)
        ln1 = b._norm(hidden, "ln1")
        res = ln1 if s.post_ln_residual else hidden
        h = res + self._sub(b.attention, ln1, k_cache, v_cache, pos)
        ln2 = b._norm(h, "ln2")
        res = ln2 if s.post_ln_residual else h
        return res + self._sub(b.mlp, ln2)

    def forward(self, hidden_states: torch.Tensor, **_):
        return (self.forward_cached(hidden_states, None, None, 0),)


def add_deep_prompt(hidden: torch.Tensor, prompt: Optional[torch.Tensor]) -> torch.Tensor:
    """``hidden[:, :pre] += prompt`` out of place (a batch-1 prompt broadcasts), the per-block deep-prompt rule."""
    if prompt is None or prompt.numel() == 0:
        return hidden
    pre = prompt.shape[1]
    return torch.cat([hidden[:, :pre] + prompt, hidden[:, pre:]], dim=1)


def span_forward(blocks: Sequence[nn.Module], hidden: torch.Tensor, prompts: Optional[Sequence[Optional[torch.Tensor]]] = None) -> torch.Tensor:
    h = hidden
    with torch.no_grad():
        for i, block in enumerate(blocks):
            h = block.forward_cached(add_deep_prompt(h, prompts[i] if prompts is not None else None), None, None, 0)
    return h


def span_backward(blocks: Sequence[nn.Module], hidden: torch.Tensor, grad_out: torch.Tensor,
                  prompts: Optional[Sequence[Optional[torch.Tensor]]] = None):
    """Gradient of a span of frozen blocks wrt its input and its deep prompts: one no-grad forward that keeps every block's
    input, then per-block recompute under autograd from the last block to the first (2 forwards per step in total).
    With :class:`ShardedBlock` s this is a collective: every rank of the group calls it with the same arguments and gets
    the same result."""
    n = len(blocks)
    prompts = list(prompts) if prompts is not None else [None] * n
    inputs, h = [], hidden
    with torch.no_grad():
        for i, block in enumerate(blocks):
            inputs.append(h)
            if i + 1 < n:
                h = block.forward_cached(add_deep_prompt(h, prompts[i]), None, None, 0)
    grad = grad_out
    grad_prompts: List[Optional[torch.Tensor]] = [None] * n
    for i in reversed(range(n)):
        x = inputs[i].detach().requires_grad_(True)
        p = prompts[i]
        p = p.detach().requires_grad_(True) if p is not None and p.numel() else None
        with torch.enable_grad():
            y = blocks[i].forward_cached(add_deep_prompt(x, p), None, None, 0)
        grads = torch.autograd.grad(y, [x] + ([p] if p is not None else []), grad)
        grad = grads[0]
        if p is not None:
            grad_prompts[i] = grads[1]
    return grad, grad_prompts

"""Serving back-ends: ``Stage`` executes a span of blocks, ``TransformerBackend`` is the per-block facade
with the reference's name and methods (reference: src/petals/server/backend.py:24-235).

``Stage`` has two executors behind one interface:
  * CUDA + supported family -> :class:`petals_b200.server.stage_engine.StageEngine` (fused sm_100a kernels,
    paged KV, CUDA graphs) for inference and forward;
  * otherwise (CPU plumbing tests, exotic configs) -> the oracle blocks with dense per-session KV tensors.
Backward (prompt-tuning) recomputes each block once with autograd enabled (activation checkpointing per
block: 2 forwards per step instead of the reference's 3) and returns the gradient
w.r.t. the span input and the deep prompts; weights are frozen (reference backend.py:48-51).
"""
from __future__ import annotations

import os
import threading
import time
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from petals_b200.models.block_oracle import GenericBlock
from petals_b200.server.memory_cache import MemoryCache, SessionCache
from petals_b200.server.task_pool import PrioritizedTaskPool, Runtime
from petals_b200.utils.logging import get_logger
from petals_b200.utils.tracing import nvtx_range
from petals_b200.utils.misc import is_dummy

logger = get_logger(__name__)


class Stage:
    """A contiguous span ``[start_block, end_block)`` of one model living on one device."""

    def __init__(self, config, blocks: Sequence[GenericBlock], start_block: int, *, device, memory_cache: MemoryCache,
                 torch_dtype: torch.dtype, max_chunk_size_bytes: int = 256 * 1024 * 1024, use_cuda_graphs: bool = True,
                 force_oracle: bool = False, engine=None, fp8: bool = False):
        self.config, self.blocks = config, list(blocks)
        self.start_block, self.end_block = start_block, start_block + len(blocks)
        self.device, self.dtype = torch.device(device), torch_dtype
        self.memory_cache = memory_cache
        self.spec = config.block_spec()
        self.max_chunk_size_bytes = max_chunk_size_bytes
        self.engine = engine  # e.g. a tensor-parallel leader engine (parallel/tp_worker.py)
        if engine is None and memory_cache.paged and not force_oracle:
            from petals_b200.server.stage_engine import StageEngine

            chunk_tokens = max(256, min(8192, max_chunk_size_bytes // max(1, 2 * self.spec.intermediate_size)))
            self.engine = StageEngine(self.spec, self.blocks, memory_cache, device=self.device, max_chunk_tokens=chunk_tokens,
                                      use_cuda_graphs=use_cuda_graphs, fp8=fp8)
        self.fp8 = fp8 and self.engine is not None
        self.active_adapter: Optional[str] = None
        # training over the NVLink fabric: the span input of a forward micro-batch stays here until its backward arrives (the client
        # never sees intermediate activations on that path, so it cannot send them back like the reference's client does)
        self._stash: "OrderedDict[str, Tuple[torch.Tensor, float]]" = OrderedDict()
        self._stash_lock = threading.Lock()

    def __len__(self) -> int:
        return len(self.blocks)

    # ---- adapters -------------------------------------------------------------------------------------
    def use_adapter(self, name: Optional[str]) -> None:
        """Activate a LoRA adapter for the *next call only* (per request, not process-global: Q10)."""
        from petals_b200.utils.peft import set_active_adapter

        if name != self.active_adapter:
            for block in self.blocks:
                set_active_adapter(block, name)
            if getattr(self.engine, "lora_on_engine", False):
                self.engine.use_adapter(name)  # merged-weight views + per-adapter graphs (server/stage_engine.py)
            self.active_adapter = name

    # ---- oracle helpers ----------------------------------------------------------------------------------
    @staticmethod
    def _add_prompt(hidden: torch.Tensor, prompt: Optional[torch.Tensor]) -> torch.Tensor:
        if prompt is None or is_dummy(prompt):
            return hidden
        P = min(prompt.shape[1], hidden.shape[1])
        if P == 0:
            return hidden
        return torch.cat([hidden[:, :P] + prompt[:, :P].to(hidden.dtype), hidden[:, P:]], dim=1)

    def _oracle_chunk_len(self, batch: int, prefix: int, T: int) -> int:
        """Chunk length such that attention logits stay below max_chunk_size_bytes (reference backend.py:146-152)."""
        heads = self.spec.num_heads
        per_token = max(1, heads * batch * 4 * max(prefix + T, 1))
        return max(1, self.max_chunk_size_bytes // per_token)

    # ---- public span ops -----------------------------------------------------------------------------------
    STASH_ENTRIES, STASH_TTL = 64, 600.0

    def stash_put(self, key: str, hidden: torch.Tensor) -> None:
        now = time.monotonic()
        with self._stash_lock:
            self._stash[key] = (hidden, now)
            self._stash.move_to_end(key)
            while len(self._stash) > self.STASH_ENTRIES or (self._stash and now - next(iter(self._stash.values()))[1] > self.STASH_TTL):
                self._stash.popitem(last=False)

    def stash_pop(self, key: str) -> torch.Tensor:
        with self._stash_lock:
            item = self._stash.pop(key, None)
        if item is None:
            raise KeyError(f"no activations are stashed under {key!r} (expired, evicted, or this stage never ran that forward)")
        return item[0]

    def forward(self, hidden: torch.Tensor, prompts: Optional[Sequence[torch.Tensor]] = None, lo: int = 0, hi: Optional[int] = None,
                take_from: Optional[tuple] = None, push_to: Optional[tuple] = None, stash: Optional[str] = None) -> torch.Tensor:
        """``take_from = (fabric, src_rank, B, T[, slot])``: the input sits in this rank's "x_in" landing slot; ``push_to = (fabric, kind,
        rank[, slot])``: the output goes to that landing slot (fused into the span's last GEMM when it can be) and the return value is
        empty; ``stash``: keep the span input under that key for the matching :meth:`backward` (forward/backward hops of training,
        parallel/fabric.py)."""
        hi = len(self.blocks) if hi is None else hi
        if take_from is not None:
            fabric, src_rank, B, T = take_from[:4]
            hidden = fabric.recv(B * T, "x_in", src_rank, *take_from[4:5]).view(B, T, -1)  # a fresh tensor: it doubles as the stash copy
        hidden = hidden.to(self.device)
        if stash is not None:
            self.stash_put(stash, hidden if take_from is not None else hidden.clone())
        prompts = None if prompts is None else [None if is_dummy(p) else p.to(self.device).contiguous() for p in prompts]
        if self.engine is not None and self._lora_free():
            with nvtx_range(f"stage[{self.start_block + lo}:{self.start_block + hi}].forward"):
                if push_to is None:
                    return self.engine.forward(hidden, prompts, (lo, hi))
                if not getattr(self.engine, "whole_span_only", False):
                    return self.engine.forward(hidden, prompts, (lo, hi), hop=push_to)[:, :0]
                out = self.engine.forward(hidden, prompts, (lo, hi))  # a tensor-parallel leader has no fused hop: host-issued peer copy
                push_to[0].send(out.reshape(-1, out.shape[-1]), push_to[2], push_to[1], *push_to[3:4])
                return out[:, :0]
        h = hidden.to(self.dtype)
        with torch.no_grad():
            for i in range(lo, hi):
                h = self._add_prompt(h, prompts[i - lo] if prompts is not None else None)
                h = self.blocks[i].forward_cached(h, None, None, 0)
        if push_to is not None:
            push_to[0].send(h.reshape(-1, h.shape[-1]), push_to[2], push_to[1], *push_to[3:4])
            return h[:, :0]
        return h

    def _materialize(self, slot: int) -> None:
        """FP8 stages keep no bf16 projections; the oracle paths (backward) get them dequantised block by block."""
        if not getattr(self, "fp8", False):
            return
        from petals_b200.ops import functional as Fn

        for name, (q, e) in self.engine.fp8[slot].items():
            p = getattr(self.blocks[slot], name)
            if p.numel() == 0:
                p.data = Fn.dequant_mxfp8(q, e)

    def _dematerialize(self, slot: int) -> None:
        if getattr(self, "fp8", False):
            for name in self.engine.fp8[slot]:
                p = getattr(self.blocks[slot], name)
                p.data = torch.empty(0, dtype=p.dtype, device=p.device)

    def backward(self, hidden: Optional[torch.Tensor], grad_out: Optional[torch.Tensor], prompts: Optional[Sequence[torch.Tensor]] = None,
                 lo: int = 0, hi: Optional[int] = None, stash: Optional[str] = None, grad_from: Optional[tuple] = None,
                 push_to: Optional[tuple] = None) -> Tuple[torch.Tensor, List[Optional[torch.Tensor]]]:
        """Returns (grad wrt span input, [grad wrt each block's prompt or None]). Over the fabric: ``stash`` names the span input kept by
        :meth:`forward`; ``grad_from = (fabric, src_rank, B, T[, slot])`` says dL/d(output) sits in this rank's "g_in" landing slot (read
        in place by the kernels); with ``push_to = (fabric, kind, rank[, slot])`` the gradient of the span input is stored into that
        landing slot by the last backward kernel and the returned gradient is empty."""
        hi = len(self.blocks) if hi is None else hi
        grad_ack = None
        if grad_from is not None:  # first of all: the ring protocol counts transfers, a request that fails later must still consume its slot
            fabric, src_rank, B, T = grad_from[:4]
            slot = grad_from[4] if len(grad_from) > 4 else 0
            grad_out = fabric.landing(B * T, "g_in", slot).view(B, T, -1)
            grad_ack = (fabric, "g_in", src_rank, slot)
        if stash is not None:
            try:
                hidden = self.stash_pop(stash)
            except KeyError:
                if grad_ack is not None:
                    grad_ack[0].acknowledge(*grad_ack[1:])
                raise
        if tuple(hidden.shape) != tuple(grad_out.shape):
            raise ValueError(f"inputs {tuple(hidden.shape)} and grad_outputs {tuple(grad_out.shape)} must have the same shape")
        engine_hops = (self.engine is not None and self._lora_free() and not getattr(self.engine, "whole_span_only", False)
                       and getattr(self.engine, "backward_supported", lambda: False)() and os.environ.get("PETALS_B200_ENGINE_BACKWARD", "1") != "0")
        if (grad_ack is not None or push_to is not None) and not engine_hops:
            # executors without the fused hops honour the protocol with host-issued copies around the plain call
            if grad_ack is not None:
                grad_out = grad_out.clone()
                fabric.acknowledge(*grad_ack[1:])
            g, gp = self.backward(hidden, grad_out, prompts, lo, hi)
            if push_to is not None:
                push_to[0].send(g.reshape(-1, g.shape[-1]), push_to[2], push_to[1], *push_to[3:4])
                g = g[:, :0]
            return g, gp
        if getattr(self.engine, "whole_span_only", False):
            # a tensor-parallel stage: the weights exist only as per-rank shards, so the whole worker group runs the recompute
            if not hasattr(self.engine, "backward"):
                raise NotImplementedError("backward through this tensor-parallel stage is not implemented; serve prompt-tuning from pipeline stages")
            ps = None if prompts is None else [None if is_dummy(p) else p.to(self.device, self.dtype).contiguous() for p in prompts]
            return self.engine.backward(hidden.to(self.device, self.dtype), grad_out.to(self.device, self.dtype), ps, (lo, hi))
        hidden, grad = hidden.to(self.device, self.dtype), grad_out.to(self.device, self.dtype)
        prompts = [None] * (hi - lo) if prompts is None else [None if is_dummy(p) else p.to(self.device, self.dtype).contiguous() for p in prompts]
        if (self.engine is not None and self._lora_free() and getattr(self.engine, "backward_supported", lambda: False)()
                and os.environ.get("PETALS_B200_ENGINE_BACKWARD", "1") != "0"):
            # the whole backward on the kernels: dgrad GEMMs (tcgen05, untransposed weights), flash-attention backward, norm / SwiGLU /
            # RoPE backward kernels (server/stage_engine.py:backward)
            with torch.no_grad():
                g, gp = self.engine.backward(hidden, grad, prompts, (lo, hi), grad_hop=push_to, grad_ack=grad_ack)
                return (g if push_to is None else hidden[:, :0]), gp
        # pass 1 (no grad): remember every block's input
        inputs = []
        h = hidden
        use_engine = self.engine is not None and self._lora_free()  # pass 1 is a plain forward: run it on the kernels
        with torch.no_grad():
            for i in range(lo, hi):
                inputs.append(h)
                if i + 1 < hi:
                    if use_engine:
                        h = self.engine.forward(h, [prompts[i - lo]] if prompts[i - lo] is not None else None, (i, i + 1))
                    else:
                        self._materialize(i)
                        h = self.blocks[i].forward_cached(self._add_prompt(h, prompts[i - lo]), None, None, 0)
                        self._dematerialize(i)
        # pass 2: per-block recompute with autograd, last block first
        grad_prompts: List[Optional[torch.Tensor]] = [None] * (hi - lo)
        for i in reversed(range(lo, hi)):
            x = inputs[i - lo].detach().requires_grad_(True)
            p = prompts[i - lo]
            p = p.detach().requires_grad_(True) if p is not None else None
            self._materialize(i)
            with torch.enable_grad():
                y = self.blocks[i].forward_cached(self._add_prompt(x, p), None, None, 0)
            targets = [x] + ([p] if p is not None else [])
            grads = torch.autograd.grad(y, targets, grad)
            self._dematerialize(i)
            grad = grads[0]
            if p is not None:
                grad_prompts[i - lo] = grads[1]
        return grad, grad_prompts

    def inference_step(self, session: SessionCache, hidden: torch.Tensor, prompts: Optional[Sequence[torch.Tensor]] = None,
                       hypo_ids: Optional[torch.Tensor] = None, lo: int = 0, hi: Optional[int] = None,
                       take_from: Optional[tuple] = None, push_to: Optional[tuple] = None) -> torch.Tensor:
        hi = len(self.blocks) if hi is None else hi
        hidden = hidden.to(self.device)
        prompts = None if prompts is None else [None if is_dummy(p) else p.to(self.device).contiguous() for p in prompts]
        if self.engine is not None and self._lora_free():
            with nvtx_range(f"stage[{self.start_block + lo}:{self.start_block + hi}].inference_step"):
                fused_hops = not getattr(self.engine, "whole_span_only", False)  # a tensor-parallel leader has no fused hop: host-issued copies
                if (take_from is not None or push_to is not None) and fused_hops:
                    return self.engine.inference_step(session, hidden, prompts, hypo_ids, (lo, hi), take_from=take_from, push_to=push_to)
                if take_from is not None:
                    fabric, src_rank, B, T = take_from[:4]
                    hidden = fabric.recv(B * T, "x_in", src_rank, *take_from[4:5]).view(B, T, -1)
                out = self.engine.inference_step(session, hidden, prompts, hypo_ids, (lo, hi))
                if push_to is not None:
                    push_to[0].send(out.reshape(-1, out.shape[-1]), push_to[2], push_to[1], *push_to[3:4])
                    return out[:, :0]
                return out
        if take_from is not None:  # executors without fused hops still honour the fabric protocol (host-issued copies)
            fabric, src_rank, B, T = take_from[:4]
            hidden = fabric.recv(B * T, "x_in", src_rank, *take_from[4:5]).view(B, T, -1)
        out = self._oracle_inference(session, hidden, prompts, hypo_ids, lo, hi)
        if push_to is not None:
            push_to[0].send(out.reshape(-1, out.shape[-1]), push_to[2], push_to[1], *push_to[3:4])
            return out[:, :0]
        return out

    def _lora_free(self) -> bool:
        """May this call run on the engine? Yes without an adapter, or when the engine serves adapters from merged weights."""
        return self.active_adapter is None or getattr(self.engine, "lora_on_engine", False)

    @torch.no_grad()
    def _oracle_inference(self, session: SessionCache, hidden, prompts, hypo_ids, lo: int, hi: int) -> torch.Tensor:
        B, T, _ = hidden.shape
        if hypo_ids is not None and not is_dummy(hypo_ids):
            session.reorder(hypo_ids)
        if T == 0:
            return hidden
        session.prepare_write(T)
        pos = session.position
        h = hidden.to(self.dtype)
        if prompts is not None:
            # deep prompts are added to the first positions of the step input, per block (backend.py:231-233)
            pass
        out = torch.empty_like(h)
        chunk = self._oracle_chunk_len(B, pos, T)
        for t0 in range(0, T, chunk):
            t1 = min(T, t0 + chunk)
            c = h[:, t0:t1]
            for i in range(lo, hi):
                p = prompts[i - lo] if prompts is not None else None
                if p is not None and t0 < p.shape[1]:
                    c = self._add_prompt(c, p[:, t0:t1])
                k, v = session.dense_kv(i, self.spec, self.dtype, self.device)
                c = self.blocks[i].forward_cached(c, k, v, pos + t0)
            out[:, t0:t1] = c
        session.set_position(pos + T)
        return out


class TransformerBackend:
    """One served block: uid, module, schemas and task pools. Execution is delegated to the owning ``Stage``."""

    def __init__(self, name: str, module: GenericBlock, *, stage: Stage, slot: int, max_batch_size: int, runtime: Optional[Runtime] = None):
        self.name, self.module, self.stage, self.slot = name, module, stage, slot
        self.config, self.dtype = stage.config, stage.dtype
        self.max_batch_size = max_batch_size
        self.inference_pool = PrioritizedTaskPool(self.inference_step, max_batch_size, f"{name}_inference", runtime)
        self.forward_pool = PrioritizedTaskPool(self.forward, max_batch_size, f"{name}_forward", runtime)
        self.backward_pool = PrioritizedTaskPool(self.backward, max_batch_size, f"{name}_backward", runtime, in_caller_thread=True)
        for p in module.parameters():
            p.requires_grad_(False)

    @property
    def memory_cache(self) -> MemoryCache:
        return self.stage.memory_cache

    def forward(self, hidden: torch.Tensor, prompt: Optional[torch.Tensor] = None, active_adapter: Optional[str] = None) -> torch.Tensor:
        self.stage.use_adapter(active_adapter)
        return self.stage.forward(hidden, [prompt], self.slot, self.slot + 1)

    def backward(self, hidden: torch.Tensor, grad_out: torch.Tensor, prompt: Optional[torch.Tensor] = None,
                 active_adapter: Optional[str] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        self.stage.use_adapter(active_adapter)
        g, gp = self.stage.backward(hidden, grad_out, [prompt], self.slot, self.slot + 1)
        return g, gp[0]

    def inference_step(self, hidden: torch.Tensor, hypo_ids: Optional[torch.Tensor], session: SessionCache,
                       prompt: Optional[torch.Tensor] = None, active_adapter: Optional[str] = None) -> torch.Tensor:
        self.stage.use_adapter(active_adapter)
        return self.stage.inference_step(session, hidden, [prompt], hypo_ids, self.slot, self.slot + 1)

    def get_inference_cache_descriptors(self, batch_size: int, max_length: int) -> Dict[str, object]:
        """Shape/size of the KV reservation a session of this block needs (reference backend.py:88-99)."""
        spec = self.stage.spec
        return dict(batch_size=batch_size, max_length=max_length, num_kv_heads=spec.num_kv_heads, head_dim=spec.head_dim,
                    dtype=str(self.dtype), bytes=batch_size * max_length * spec.kv_bytes_per_token(self.dtype))

    def get_pools(self) -> Sequence[PrioritizedTaskPool]:
        return self.forward_pool, self.backward_pool, self.inference_pool

    def get_info(self) -> Dict[str, object]:
        return dict(name=self.name, dtype=str(self.dtype), max_batch_size=self.max_batch_size)

    def shutdown(self) -> None:
        for p in self.module.parameters():
            p.data = torch.empty(0, dtype=p.dtype)  # release device memory promptly (reference backend.py:190-198)


class _MergedInferenceStep:
    """Runs all requested blocks of a span back to back as ONE task (reference backend.py:201-235)."""

    def __init__(self, stage: Stage):
        self.stage = stage

    def __call__(self, hidden: torch.Tensor, hypo_ids: Optional[torch.Tensor], session: SessionCache, lo: int, hi: int,
                 prompts: Optional[Sequence[torch.Tensor]], active_adapter: Optional[str], take_from: Optional[tuple] = None,
                 push_to: Optional[tuple] = None) -> torch.Tensor:
        self.stage.use_adapter(active_adapter)
        return self.stage.inference_step(session, hidden, prompts, hypo_ids, lo, hi, take_from=take_from, push_to=push_to)


def merge_inference_pools_inplace(backends: Dict[str, TransformerBackend], stage: Stage, runtime: Optional[Runtime], max_batch_size: int) -> PrioritizedTaskPool:
    """All blocks of a stage share one inference pool whose task executes the whole requested sub-span."""
    pool = PrioritizedTaskPool(_MergedInferenceStep(stage), max_batch_size, "span_inference", runtime)
    for backend in backends.values():
        backend.inference_pool = pool
    return pool

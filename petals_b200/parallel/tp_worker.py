"""A tensor-parallel worker group behaving as ONE stage of the swarm.

Rank 0 of the group is the *leader*: it owns the swarm endpoint (handler, task pools, sessions) exactly like a
single-GPU stage, but its engine is :class:`TPLeaderEngine`, which for every inference step (1) tells the followers
what to run through the shared-memory command ring, (2) pushes the step's input rows into every rank's symmetric
buffer over NVLink and (3) replays the same CUDA graph as the followers. All ranks keep mirror KV sessions (their
own pages, identical positions). This is the B200 counterpart of ``--tensor_parallel_devices``
(reference: src/petals/cli/run_server.py:154-157, src/petals/utils/convert_block.py:118-135)."""
from __future__ import annotations

import itertools
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from petals_b200.ops.functional import PAGE
from petals_b200.parallel.control import CommandRing
from petals_b200.parallel.symmetric import SymmetricHeap
from petals_b200.parallel.tensor_parallel import MAX_ROWS, TPDecodeEngine, local_spec
from petals_b200.server.memory_cache import MemoryCache, SessionCache
from petals_b200.utils.logging import get_logger
from petals_b200.utils.misc import is_dummy

logger = get_logger(__name__)


class TPLeaderEngine:
    """Drop-in for ``StageEngine`` on the leader rank (inference sessions only)."""

    def __init__(self, engine: TPDecodeEngine, ring: CommandRing):
        self.engine, self.ring = engine, ring
        self.n_blocks = engine.n_blocks
        self._sids: Dict[int, int] = {}
        self._next_sid = itertools.count(1)

    def _sid(self, session: SessionCache) -> int:
        key = id(session)
        if key not in self._sids:
            sid = next(self._next_sid)
            self._sids[key] = sid
            self.ring.send({"op": "open", "sid": sid, "B": session.batch_size, "max_length": session.max_length})

            def closed(s, sid=sid, key=key):
                self._sids.pop(key, None)
                self.ring.send({"op": "close", "sid": sid})

            session.on_close.append(closed)
        return self._sids[key]

    def inference_step(self, session: SessionCache, hidden: torch.Tensor, prompts=None, hypo_ids: Optional[torch.Tensor] = None,
                       block_range: Optional[Tuple[int, int]] = None) -> torch.Tensor:
        lo, hi = block_range or (0, self.n_blocks)
        if (lo, hi) != (0, self.n_blocks):
            raise NotImplementedError("a tensor-parallel stage serves its whole span per request")
        if prompts is not None and any(p is not None and not is_dummy(p) for p in prompts):
            raise NotImplementedError("deep prompts are not supported by the tensor-parallel engine yet")
        B, T, H = hidden.shape
        sid = self._sid(session)
        hypo = None if hypo_ids is None or is_dummy(hypo_ids) else [int(i) for i in hypo_ids.tolist()]
        if hypo is not None:
            session.reorder(hypo_ids)
        if T == 0:
            return hidden
        if B > MAX_ROWS:
            raise ValueError(f"batch of {B} sequences exceeds the tensor-parallel decode engine's {MAX_ROWS} rows")
        hidden = hidden.to(torch.bfloat16)
        out = torch.empty(B, T, H, dtype=torch.bfloat16, device=hidden.device)
        step_t = max(1, MAX_ROWS // B)  # longer inputs (prompt ingestion) go through in row-limited micro-steps
        for t0 in range(0, T, step_t):
            t1 = min(T, t0 + step_t)
            cmd = {"op": "step", "sid": sid, "B": B, "T": t1 - t0, "pos": session.position}
            if hypo is not None and t0 == 0:
                cmd["hypo"] = hypo
            self.ring.send(cmd)
            self.engine.push_inputs(hidden[:, t0:t1].reshape(B * (t1 - t0), H))
            y = self.engine.run_step(session, B, t1 - t0)
            out[:, t0:t1] = y.view(B, t1 - t0, H)
        return out

    def forward(self, hidden, prompts=None, block_range=None):
        raise NotImplementedError("cache-less forward/backward through a tensor-parallel stage is not implemented yet; "
                                  "serve training traffic from pipeline stages")

    def check_errors(self) -> None:
        self.engine.check_errors()

    def shutdown(self) -> None:
        self.ring.send({"op": "stop"})


def follower_loop(engine: TPDecodeEngine, cache: MemoryCache, ring: CommandRing, consumer: int, idle_timeout: Optional[float] = None) -> None:
    """Followers: mirror the leader's sessions and replay the same step graphs until told to stop."""
    sessions: Dict[int, SessionCache] = {}
    while True:
        cmd = ring.recv(consumer, timeout=idle_timeout)
        op = cmd["op"]
        if op == "step":
            s = sessions[cmd["sid"]]
            if cmd["pos"] != s.position:
                s.set_position(cmd["pos"])  # rollback (speculative decoding) decided on the leader
            if "hypo" in cmd:
                s.reorder(torch.tensor(cmd["hypo"], dtype=torch.int64))
            engine.run_step(s, cmd["B"], cmd["T"])
        elif op == "open":
            sessions[cmd["sid"]] = cache.open_session(cmd["B"], cmd["max_length"], timeout=None)
        elif op == "close":
            s = sessions.pop(cmd["sid"], None)
            if s is not None:
                s.close()
        elif op == "stop":
            break
        else:
            logger.warning(f"unknown command {cmd}")
    for s in sessions.values():
        s.close()
    torch.cuda.synchronize(engine.device)
    engine.check_errors()


def build_tp_engine(config, n_blocks: int, *, group=None, attn_cache_tokens: int = 4096, inference_max_length: int = 4096, blocks=None,
                    seed: int = 0, heap_bytes: Optional[int] = None, use_cuda_graphs: bool = True):
    """Collective: every rank of ``group`` builds its shard of an ``n_blocks`` span. ``blocks`` (full GenericBlocks, same on
    all ranks) are sharded if given, otherwise shards are random-initialised in place. Returns (engine, cache, heap)."""
    from petals_b200.parallel.tensor_parallel import random_shard, shard_block, tp_supported

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    device = torch.device("cuda", torch.cuda.current_device())
    spec = config.block_spec()
    if not tp_supported(spec, world):
        raise ValueError(f"{spec.family} blocks cannot be tensor-parallelised over {world} ranks by this engine")
    need = (1 + 2 * world) * MAX_ROWS * spec.hidden_size * 2 + (2 * n_blocks + 2) * 8 + (1 << 20)
    heap = SymmetricHeap(heap_bytes or (max(need, 8 << 20) + (256 << 20)), group=group, device=device)  # + room for the bandwidth probe
    if blocks is not None:
        shards = [shard_block(b, spec, rank, world, device) for b in blocks]
    else:
        shards = [random_shard(spec, rank, world, layer, device, seed) for layer in range(n_blocks)]
    ls = local_spec(spec, world)
    cache = MemoryCache(attn_cache_tokens, None, n_blocks=n_blocks, spec=ls, dtype=torch.bfloat16, device=device, paged=True,
                        max_length=inference_max_length)
    engine = TPDecodeEngine(spec, shards, heap, cache, use_cuda_graphs=use_cuda_graphs)
    torch.cuda.synchronize(device)
    dist.barrier(group=group)
    return engine, cache, heap


def make_ring(group=None, n_followers: Optional[int] = None) -> CommandRing:
    """Collective: the leader creates the ring, everyone else attaches."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    names: List[Optional[str]] = [None]
    ring = None
    if rank == 0:
        ring = CommandRing(None, create=True, n_consumers=(world - 1) if n_followers is None else n_followers)
        names[0] = ring.name
    dist.broadcast_object_list(names, src=0, group=group)
    if rank != 0:
        ring = CommandRing(names[0], create=False)
    dist.barrier(group=group)
    return ring

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
import threading
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from petals_b200.parallel.control import CommandRing
from petals_b200.parallel.symmetric import SymmetricHeap, host_barrier
from petals_b200.parallel.tensor_parallel import MAX_ROWS, TPDecodeEngine, local_spec
from petals_b200.server.memory_cache import MemoryCache, SessionCache
from petals_b200.utils.logging import get_logger
from petals_b200.utils.misc import is_dummy

logger = get_logger(__name__)


class TPLeaderEngine:
    """Drop-in for ``StageEngine`` on the leader rank."""

    whole_span_only = True  # the handler must not split a step into per-block tasks

    def __init__(self, engine: TPDecodeEngine, ring: CommandRing):
        self.engine, self.ring = engine, ring
        self.n_blocks = engine.n_blocks
        self._sids: Dict[int, int] = {}
        self._next_sid = itertools.count(1)
        self._lock = threading.RLock()  # one command stream: ring order must equal launch order on the leader

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
        with self._lock:
            return self._inference_step(session, hidden, prompts, hypo_ids, block_range)

    def _inference_step(self, session, hidden, prompts, hypo_ids, block_range) -> torch.Tensor:
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
        hidden = hidden.to(torch.bfloat16)
        out = torch.empty(B, T, H, dtype=torch.bfloat16, device=hidden.device)
        P = self.engine.max_prefill_rows
        prefill = B * T > MAX_ROWS and P >= B
        if not prefill and B > MAX_ROWS:
            raise ValueError(f"batch of {B} sequences exceeds the tensor-parallel decode engine's {MAX_ROWS} rows")
        # decode-shaped steps replay the GEMV graph; prompt ingestion goes through sequence-parallel GEMM chunks
        step_t = max(1, (P if prefill else MAX_ROWS) // B)
        for t0 in range(0, T, step_t):
            t1 = min(T, t0 + step_t)
            chunk_prefill = prefill and B * (t1 - t0) > MAX_ROWS
            cmd = {"op": "prefill" if chunk_prefill else "step", "sid": sid, "B": B, "T": t1 - t0, "pos": session.position}
            if hypo is not None and t0 == 0:
                cmd["hypo"] = hypo
            self.ring.send(cmd)
            rows = hidden[:, t0:t1].reshape(B * (t1 - t0), H)
            if chunk_prefill:
                self.engine.push_prefill_inputs(rows)
                y = self.engine.run_prefill(session, B, t1 - t0)
            else:
                self.engine.push_inputs(rows)
                y = self.engine.run_step(session, B, t1 - t0)
            out[:, t0:t1] = y.view(B, t1 - t0, H)
        return out

    def forward(self, hidden: torch.Tensor, prompts=None, block_range: Optional[Tuple[int, int]] = None) -> torch.Tensor:
        """Cache-less parallel forward (rpc_forward): the prompt chunks run through the sequence-parallel prefill path against a
        scratch KV session that is dropped afterwards (every rank mirrors it)."""
        B, T, _ = hidden.shape
        with self._lock:
            session = self.engine.cache.open_session(B, T, timeout=None)
            try:
                return self._inference_step(session, hidden, prompts, None, block_range).clone()
            finally:
                session.close()

    def backward(self, hidden: torch.Tensor, grad_out: torch.Tensor, prompts=None, block_range: Optional[Tuple[int, int]] = None):
        """Gradient of the span wrt its input and deep prompts (rpc_backward). No fused kernels here: every rank runs the
        autograd recompute on its shard and the two all-reduces per block go through NCCL (parallel/tp_generic.py)."""
        lo, hi = block_range or (0, self.n_blocks)
        B, T, H = hidden.shape
        n = hi - lo
        plist = [None] * n if prompts is None else [None if (p is None or is_dummy(p)) else p for p in prompts]
        if len(plist) != n or grad_out.shape != hidden.shape:
            raise ValueError("backward: mismatched prompts / gradient shapes")
        shapes = [None if p is None else [int(p.shape[0]), int(p.shape[1])] for p in plist]
        with self._lock:
            self.ring.send({"op": "backward", "B": B, "T": T, "lo": lo, "hi": hi, "prompts": shapes})
            return tp_collective_backward(self.engine, hidden, grad_out, plist, lo, hi, shapes)

    def check_errors(self) -> None:
        self.engine.check_errors()

    def shutdown(self) -> None:
        self.ring.send({"op": "stop"})


def _sharded_blocks(engine: TPDecodeEngine):
    """ShardedBlock views of the engine's weight shards (built once; the parameters alias the shard tensors)."""
    blocks = getattr(engine, "_sharded_blocks", None)
    if blocks is None:
        from petals_b200.parallel.tp_generic import ShardedBlock

        blocks = [ShardedBlock.from_tensors(t, engine.spec, engine.rank, engine.world, getattr(engine.heap, "group", None)) for t in engine.shards]
        engine._sharded_blocks = blocks
    return blocks


def tp_collective_backward(engine: TPDecodeEngine, hidden, grad_out, prompts, lo: int, hi: int, prompt_shapes):
    """Collective over the TP group. The leader passes the real tensors; followers pass ``None`` and receive them by broadcast."""
    from petals_b200.parallel.tp_generic import span_backward

    dev, H = engine.device, engine.spec.hidden_size
    group = getattr(engine.heap, "group", None)
    src = dist.get_global_rank(group, 0) if group is not None else 0

    def bcast(t, shape):
        buf = t.to(dev, torch.bfloat16).contiguous() if t is not None else torch.empty(shape, dtype=torch.bfloat16, device=dev)
        dist.broadcast(buf, src=src, group=group)
        return buf

    B, T = (hidden.shape[0], hidden.shape[1]) if hidden is not None else prompt_shapes["BT"]
    shapes = prompt_shapes["prompts"] if isinstance(prompt_shapes, dict) else prompt_shapes
    x = bcast(hidden, (B, T, H))
    g = bcast(grad_out, (B, T, H))
    ps = [None if sh is None else bcast(prompts[i] if prompts is not None else None, (sh[0], sh[1], H)) for i, sh in enumerate(shapes)]
    grad, grad_prompts = span_backward(_sharded_blocks(engine)[lo:hi], x, g, ps)
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    return grad, grad_prompts


def follower_loop(engine: TPDecodeEngine, cache: MemoryCache, ring: CommandRing, consumer: int, idle_timeout: Optional[float] = None) -> None:
    """Followers: mirror the leader's sessions and replay the same step graphs until told to stop."""
    sessions: Dict[int, SessionCache] = {}
    while True:
        cmd = ring.recv(consumer, timeout=idle_timeout)
        op = cmd["op"]
        if op in ("step", "prefill"):
            s = sessions[cmd["sid"]]
            if cmd["pos"] != s.position:
                s.set_position(cmd["pos"])  # rollback (speculative decoding) decided on the leader
            if "hypo" in cmd:
                s.reorder(torch.tensor(cmd["hypo"], dtype=torch.int64))
            (engine.run_step if op == "step" else engine.run_prefill)(s, cmd["B"], cmd["T"])
        elif op == "backward":
            tp_collective_backward(engine, None, None, None, cmd["lo"], cmd["hi"], {"BT": (cmd["B"], cmd["T"]), "prompts": cmd["prompts"]})
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
                    seed: int = 0, heap_bytes: Optional[int] = None, use_cuda_graphs: bool = True, max_prefill_rows: int = 4096):
    """Collective: every rank of ``group`` builds its shard of an ``n_blocks`` span. ``blocks`` (full GenericBlocks, same on
    all ranks) are sharded if given, otherwise shards are random-initialised in place. Returns (engine, cache, heap)."""
    from petals_b200.parallel.tensor_parallel import prefill_heap_bytes, random_shard, shard_block, tp_supported

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    device = torch.device("cuda", torch.cuda.current_device())
    spec = config.block_spec()
    if not tp_supported(spec, world):
        raise ValueError(f"{spec.family} blocks cannot be tensor-parallelised over {world} ranks by this engine")
    need = (1 + 2 * world) * MAX_ROWS * spec.hidden_size * 2 + (2 * n_blocks + 2) * 8 + (1 << 20)
    need += 2 * world * MAX_ROWS * spec.hidden_size * 4 + (1 << 16)  # LL all-reduce buffers
    need += prefill_heap_bytes(spec.hidden_size, world, n_blocks, max_prefill_rows)
    heap = SymmetricHeap(heap_bytes or (max(need, 8 << 20) + (256 << 20)), group=group, device=device)  # + room for the bandwidth probe
    if blocks is not None:
        shards = [shard_block(b, spec, rank, world, device) for b in blocks]
    else:
        shards = [random_shard(spec, rank, world, layer, device, seed) for layer in range(n_blocks)]
    ls = local_spec(spec, world)
    cache = MemoryCache(attn_cache_tokens, None, n_blocks=n_blocks, spec=ls, dtype=torch.bfloat16, device=device, paged=True,
                        max_length=inference_max_length)
    engine = TPDecodeEngine(spec, shards, heap, cache, use_cuda_graphs=use_cuda_graphs, max_prefill_rows=max_prefill_rows)
    torch.cuda.synchronize(device)
    host_barrier(group)
    return engine, cache, heap


def make_ring(group=None, n_followers: Optional[int] = None) -> CommandRing:
    """Collective: the leader creates the ring, everyone else attaches."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    names: List[Optional[str]] = [None]
    ring = None
    if rank == 0:
        ring = CommandRing(None, create=True, n_consumers=(world - 1) if n_followers is None else n_followers)
        names[0] = ring.name
    dist.broadcast_object_list(names, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if rank != 0:
        ring = CommandRing(names[0], create=False)
    host_barrier(group)
    return ring


# ---- a TP group started from ONE server process (`run_server --tensor_parallel_devices cuda:0 cuda:1 ...`) -------------------
class TPGroup:
    """Owns the follower processes of a tensor-parallel stage. The calling (server) process becomes rank 0 / the leader: it
    initialises ``torch.distributed`` for the group, loads + shards its blocks, builds the engine and the command ring;
    every other device gets a spawned worker running :func:`_follower_main` on the same checkpoint."""

    def __init__(self, config, converted_model_name_or_path: str, block_indices: Sequence[int], devices: Sequence[torch.device], *,
                 torch_dtype: torch.dtype, attn_cache_tokens: int, inference_max_length: int, use_cuda_graphs: bool = True,
                 max_prefill_rows: int = 4096, start_timeout: float = 600.0):
        import socket

        import torch.multiprocessing as mp

        from petals_b200.parallel.tensor_parallel import tp_supported

        if torch_dtype != torch.bfloat16:
            raise ValueError("tensor-parallel stages run in bfloat16")
        spec = config.block_spec()
        world = len(devices)
        if not tp_supported(spec, world):
            raise ValueError(f"{spec.family} blocks cannot be tensor-parallelised over {world} devices by this engine "
                             f"(needs kv heads, heads and FFN columns divisible by {world}, sequential attention/MLP, no fused-interleaved QKV)")
        if dist.is_available() and dist.is_initialized():
            raise RuntimeError("this process already belongs to a torch.distributed job; start tensor-parallel stages from a fresh server process")
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        self.world, self.devices = world, [torch.device(d) for d in devices]
        ctx = mp.get_context("spawn")
        self.procs = []
        common = dict(world=world, port=port, path=converted_model_name_or_path, block_indices=list(block_indices),
                      attn_cache_tokens=attn_cache_tokens, inference_max_length=inference_max_length, use_cuda_graphs=use_cuda_graphs,
                      max_prefill_rows=max_prefill_rows)
        for rank in range(1, world):
            p = ctx.Process(target=_follower_main, kwargs=dict(rank=rank, device=str(self.devices[rank]), **common), daemon=True,
                            name=f"tp-follower-{rank}")
            p.start()
            self.procs.append(p)
        self.engine, self.cache, self.heap, self.ring = _join_group(rank=0, device=self.devices[0], config=config, **common)
        self.leader = TPLeaderEngine(self.engine, self.ring)

    def shutdown(self, timeout: float = 30.0) -> None:
        try:
            self.leader.shutdown()
        finally:
            for p in self.procs:
                p.join(timeout)
                if p.is_alive():
                    p.terminate()
            try:
                self.heap.close()
                dist.destroy_process_group()
            except Exception:  # noqa: BLE001 - best effort on the way out
                pass


def _join_group(*, rank: int, device, world: int, port: int, path: str, block_indices, attn_cache_tokens: int, inference_max_length: int,
                use_cuda_graphs: bool, max_prefill_rows: int, config=None):
    """Collective over the TP group: initialise the process group, load + shard this rank's blocks, build engine and ring."""
    from petals_b200.parallel.tensor_parallel import TPDecodeEngine, local_spec, prefill_heap_bytes, shard_block
    from petals_b200.server.from_pretrained import load_pretrained_block
    from petals_b200.utils.auto_config import AutoDistributedConfig

    device = torch.device(device)
    torch.cuda.set_device(device)
    dist.init_process_group(backend="cpu:gloo,cuda:nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=device)
    config = config or AutoDistributedConfig.from_pretrained(path)
    spec = config.block_spec()
    n_blocks = len(block_indices)
    need = (1 + 2 * world) * MAX_ROWS * spec.hidden_size * 2 + (2 * n_blocks + 2) * 8 + (1 << 20)
    need += 2 * world * MAX_ROWS * spec.hidden_size * 4 + (1 << 16)
    need += prefill_heap_bytes(spec.hidden_size, world, n_blocks, max_prefill_rows)
    heap = SymmetricHeap(max(need, 8 << 20) + (64 << 20), device=device)
    shards = []
    for block_index in block_indices:  # one block at a time: the dense block never lives on the GPU
        block = load_pretrained_block(path, block_index, config=config, torch_dtype=torch.bfloat16)
        shards.append(shard_block(block, spec, rank, world, device))
        del block
    cache = MemoryCache(attn_cache_tokens, None, n_blocks=n_blocks, spec=local_spec(spec, world), dtype=torch.bfloat16, device=device, paged=True,
                        max_length=inference_max_length)
    engine = TPDecodeEngine(spec, shards, heap, cache, use_cuda_graphs=use_cuda_graphs, max_prefill_rows=max_prefill_rows)
    torch.cuda.synchronize(device)
    host_barrier()
    ring = make_ring()
    return engine, cache, heap, ring


def _follower_main(*, rank: int, device: str, **kwargs) -> None:
    engine, cache, heap, ring = _join_group(rank=rank, device=device, **kwargs)
    try:
        follower_loop(engine, cache, ring, rank - 1)
    finally:
        heap.close()
        dist.destroy_process_group()

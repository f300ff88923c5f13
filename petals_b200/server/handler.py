"""Request front-end of a stage (reference: src/petals/server/handler.py:55-592 and
src/petals/server/block_functions.py:1-237, which this module + block_functions.py re-implement).

The reference forks ``num_handlers`` processes that deserialise protobufs, talk to the runtime through mp
queues and keep a cross-process session registry. One B200 worker is a single process, so the handler is an
ordinary object whose methods *are* the RPCs; ``parallel/transport.py`` exposes the same methods to other
processes. Contract kept from the reference:

* ``rpc_inference(uids, metadata)`` opens a stream with a server-side KV session (``max_length`` checked
  against ``inference_max_length``, cache reserved for all requested blocks with ``alloc_timeout``); every
  ``step(hidden, prompts, hypo_ids, metadata)`` may carry ``start_from_position`` (KV rollback, *validated*
  here unlike the reference's always-true assert), 0-token steps are legal, exceeding
  ``max_length`` raises ``ValueError("Maximum length exceeded ...")``;
* short steps (``B*T <= MAX_SHORT_INFERENCE_TOKENS``) run the whole requested span as ONE runtime task, longer
  ones are submitted block by block so latency-critical tasks of other sessions can interleave;
* ``rpc_forward`` / ``rpc_backward`` are stateless; ``rpc_backward`` returns ``grad_inputs`` and, if prompts
  were given, ``grad_prompts``; ``rpc_push`` lets the previous stage deliver a step's input directly
  (de-duplicated by ``step_id``); ``rpc_info`` reports version, cache budget and schemas.
"""
from __future__ import annotations

import collections
import concurrent.futures
import threading
import time
import uuid
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

import petals_b200
from petals_b200.data_structures import ModuleUID, split_uids
from petals_b200.server.backend import Stage, TransformerBackend
from petals_b200.server.block_functions import MAX_SHORT_INFERENCE_TOKENS, run_rpc_backward, run_rpc_forward
from petals_b200.server.memory_cache import SessionCache
from petals_b200.server.task_pool import PrioritizedTaskPool
from petals_b200.server.task_prioritizer import DummyTaskPrioritizer, TaskPrioritizerBase
from petals_b200.utils.logging import get_logger
from petals_b200.utils.metrics import ServerMetrics
from petals_b200.utils.fault_injection import maybe_fail
from petals_b200.utils.misc import is_dummy

CACHE_TOKENS_AVAILABLE = "cache_tokens_available"  # rpc_info key (reference: handler.py:52)

logger = get_logger(__name__)


def fabric_endpoints(metadata: Dict[str, Any]) -> Tuple[Optional[tuple], Optional[tuple]]:
    """Validated ``(take_from, push_to)`` of a request that rides the NVLink fabric (parallel/fabric.py): ``fabric_in = {B, T, src_rank,
    slot}`` says the tensor argument already sits in a landing slot of this rank, ``fabric_out = {kind, rank, slot}`` asks for the result to
    be stored into a landing slot of ``rank``. Never trust the wire: sizes and slots index device memory."""
    from petals_b200.parallel.fabric import KINDS, get_fabric

    fin, fout = metadata.get("fabric_in"), metadata.get("fabric_out")
    if fin is None and fout is None:
        return None, None
    fabric = get_fabric()
    if fabric is None:
        raise RuntimeError("this stage has no NVLink fabric but the request names landing slots")
    n_slots, world = getattr(fabric, "n_slots", 1), getattr(fabric, "world", None)
    take_from = push_to = None
    if fin is not None:
        fb, ft, src, slot = int(fin["B"]), int(fin["T"]), int(fin["src_rank"]), int(fin.get("slot", 0))
        if fb < 1 or ft < 1 or fb * ft > fabric.max_tokens or src < 0 or (world is not None and src >= world):
            raise ValueError(f"fabric_in describes {fb} x {ft} rows from rank {src}: outside this stage's landing zone ({fabric.max_tokens} rows)")
        if not 0 <= slot < n_slots:
            raise ValueError(f"fabric_in names landing slot {slot}; this stage's ring has {n_slots}")
        take_from = (fabric, src, fb, ft, slot)
    if fout is not None:
        kind, rank = str(fout["kind"]), int(fout["rank"])
        if kind not in KINDS or rank < 0 or (world is not None and rank >= world):
            raise ValueError(f"fabric_out names landing zone {kind!r} of rank {rank}")
        push_to = (fabric, kind, rank, int(fout.get("slot", 0)) % max(1, n_slots))
    return take_from, push_to


def drain_landing(take_from: Optional[tuple], kind: str, device) -> None:
    """A request that announced a transfer into a landing slot of this rank failed BEFORE the stage consumed it. The ring protocol counts
    transfers (parallel/fabric.py), so the slot must still be consumed and acknowledged once — otherwise the producer's next push
    through that slot waits for an acknowledgement that never comes. Best effort: the wait is enqueued like a normal take (device
    watchdog / host time-out apply when the producer never delivered)."""
    if take_from is None:
        return
    fabric, src_rank, B, T = take_from[:4]
    try:
        if torch.device(device).type == "cuda":
            with torch.cuda.device(device):
                fabric.recv(B * T, kind, src_rank, *take_from[4:5])
        else:
            fabric.recv(B * T, kind, src_rank, *take_from[4:5])
    except Exception as e:  # noqa: BLE001
        logger.warning(f"could not drain landing slot {kind}{list(take_from[4:5])} after a failed request: {e!r}")


class InferenceStream:
    """Server side of one ``rpc_inference`` stream: a KV session + step loop state."""

    MAX_TRACKED_STEPS = 64  # step ids remembered for push de-duplication (a push older than that can only be stale)

    def __init__(self, handler: "TransformerConnectionHandler", uids: Sequence[ModuleUID], metadata: Dict[str, Any]):
        self.handler = handler
        self.uids = list(uids)
        self.backends = [handler.module_backends[u] for u in self.uids]
        self.lo, self.hi = self.backends[0].slot, self.backends[-1].slot + 1
        self.max_length = int(metadata.get("max_length", 0))
        if self.max_length <= 0:
            raise ValueError("rpc_inference requires a positive max_length")
        if self.max_length > handler.inference_max_length:
            raise ValueError(f"Cannot allocate KV cache for {self.max_length} tokens, max = {handler.inference_max_length}")
        self.session_id = metadata.get("session_id") or str(uuid.uuid4())
        self.active_adapter = metadata.get("active_adapter") or None
        handler.check_adapter(self.active_adapter)
        self.points = float(metadata.get("points", 0))
        self.alloc_timeout = float(metadata.get("alloc_timeout", 0.0))
        self.cache: Optional[SessionCache] = None
        self.batch_size: Optional[int] = None
        self.opened_at = self.last_step_at = time.monotonic()
        self.closed = False
        self._pushed: Dict[str, Tuple[torch.Tensor, ...]] = {}
        self._last_tokens = 0
        self._done_steps: "collections.OrderedDict[str, None]" = collections.OrderedDict()  # most recent step ids only (bounded)
        self._lock = threading.Lock()
        handler._register(self)

    # ---- lifecycle ----------------------------------------------------------------------------------------
    def _ensure_cache(self, batch_size: int) -> SessionCache:
        if self.cache is None:
            self.batch_size = batch_size
            self.cache = self.handler.stage.memory_cache.open_session(batch_size, self.max_length, self.alloc_timeout)
        elif batch_size != self.batch_size:
            raise ValueError(f"batch size changed within a session ({self.batch_size} -> {batch_size})")
        return self.cache

    def close(self) -> None:
        with self._lock:
            if self.closed:
                return
            self.closed = True
            if self.cache is not None:
                self.cache.close()
            self.handler._unregister(self)

    @property
    def position(self) -> int:
        return 0 if self.cache is None else self.cache.position

    def expired(self, now: float) -> bool:
        return (now - self.opened_at > self.handler.session_timeout) or (now - self.last_step_at > self.handler.step_timeout)

    # ---- pushed inputs (server-to-server) ---------------------------------------------------------------------
    def push(self, step_id: str, tensors: Tuple[torch.Tensor, ...]) -> None:
        with self._lock:
            if step_id not in self._done_steps:
                self._pushed[step_id] = tensors
                while len(self._pushed) > self.MAX_TRACKED_STEPS:  # pushes for steps the client never sent
                    self._pushed.pop(next(iter(self._pushed)))

    # ---- one step -------------------------------------------------------------------------------------------
    def step(self, hidden: torch.Tensor, prompts: Optional[torch.Tensor] = None, hypo_ids: Optional[torch.Tensor] = None,
             metadata: Optional[Dict[str, Any]] = None) -> torch.Tensor:
        t0 = time.perf_counter()
        try:
            out = self._step(hidden, prompts, hypo_ids, metadata)
        except Exception:
            self.handler.metrics.error("inference")
            raise
        self.handler.metrics.observe("inference", self._last_tokens, time.perf_counter() - t0)
        return out

    def _step(self, hidden: torch.Tensor, prompts: Optional[torch.Tensor] = None, hypo_ids: Optional[torch.Tensor] = None,
              metadata: Optional[Dict[str, Any]] = None) -> torch.Tensor:
        metadata = metadata or {}
        take_from, push_to = fabric_endpoints(metadata)
        state = {"taken": take_from is None}
        try:
            return self._step_guarded(hidden, prompts, hypo_ids, metadata, take_from, push_to, state)
        except BaseException:
            if not state["taken"]:
                drain_landing(take_from, "x_in", self.handler.stage.device)
            raise

    def _step_guarded(self, hidden, prompts, hypo_ids, metadata, take_from, push_to, state) -> torch.Tensor:
        maybe_fail("rpc_inference", self.handler.peer_id)
        if self.closed:
            raise RuntimeError("inference session is closed")
        now = time.monotonic()
        if self.expired(now):
            self.close()
            raise TimeoutError("inference session timed out on the server")
        self.last_step_at = now
        step_id = metadata.get("step_id")
        if step_id is not None:
            with self._lock:
                pushed = self._pushed.pop(step_id, None)
                self._done_steps[step_id] = None
                while len(self._done_steps) > self.MAX_TRACKED_STEPS:
                    self._done_steps.popitem(last=False)
            # the previous stage may already have delivered this step's input (server-to-server push). It is only a latency shortcut for
            # the tensor the client sends anyway, and the client is the authority on what this stage must see: a predecessor that is
            # replaying a longer history after a fail-over pushes more positions than this (healthy) session needs — such a push is dropped
            if pushed is not None and isinstance(hidden, torch.Tensor) and tuple(pushed[0].shape) == tuple(hidden.shape):
                hidden = pushed[0]
                prompts = pushed[1] if len(pushed) > 1 else prompts
                hypo_ids = pushed[2] if len(pushed) > 2 else hypo_ids
        # NVLink fabric (parallel/fabric.py): the input may already sit in this rank's landing zone, and/or the output
        # may have to be stored straight into the next stage's (or the client's) landing zone by the span's last kernel
        if take_from is not None:
            hidden = torch.empty(take_from[2], take_from[3], self.handler.stage.spec.hidden_size, dtype=self.handler.stage.dtype,
                                 device=self.handler.stage.device)  # shape carrier only
        if hidden.dim() != 3 or not hidden.is_floating_point():
            raise ValueError(f"hidden states must be a floating-point tensor [batch, seq, hidden], got {tuple(hidden.shape)} {hidden.dtype}")
        if hidden.shape[2] != self.handler.stage.spec.hidden_size:  # kernels index by the model's hidden size: never trust the wire
            raise ValueError(f"hidden states have hidden size {hidden.shape[2]}, this model's is {self.handler.stage.spec.hidden_size}")
        B, T, H = hidden.shape
        self._last_tokens = B * T  # for the metrics: a fused stage hop returns an empty tensor
        cache = self._ensure_cache(B)
        start = metadata.get("start_from_position")
        if start is not None:
            start = int(start)
            if not 0 <= start <= cache.position:
                raise ValueError(f"start_from_position={start} must be within the current prefix [0, {cache.position}]")
            cache.set_position(start)
        prefix = cache.position
        if prefix + T > self.max_length:
            raise ValueError(f"Maximum length exceeded: prefix {prefix} + current {T} exceeds pre-allocated maximum {self.max_length}")
        n = self.hi - self.lo
        if prompts is None or is_dummy(prompts):
            block_prompts = None
        else:
            if prompts.dim() != 4 or prompts.shape[0] != n or prompts.shape[1] not in (1, B) or prompts.shape[3] != H:
                raise ValueError(f"prompts must be [{n}, {B} or 1, pre_seq_len, {H}], got {tuple(prompts.shape)}")
            block_prompts = list(prompts.unbind(0))
        if hypo_ids is not None and not is_dummy(hypo_ids):
            if hypo_ids.dtype != torch.int64 or hypo_ids.shape != (B,):
                raise ValueError(f"hypo_ids must be int64 [{B}]")
            if bool(((hypo_ids < 0) | (hypo_ids >= B)).any()):  # they index KV pages / cache rows: never trust them unchecked
                raise ValueError(f"hypo_ids must index the {B} sequences of the batch")
        else:
            hypo_ids = None
        priority = self.handler.prioritizer.prioritize(hidden, hypo_ids, points=self.points / max(n, 1), type="inference")
        h = self.handler
        whole_span = getattr(h.stage.engine, "whole_span_only", False)  # tensor-parallel groups step their span as one unit
        if B * T <= MAX_SHORT_INFERENCE_TOKENS or n == 1 or take_from is not None or push_to is not None or whole_span:
            state["taken"] = True  # from here on the stage consumes the landing slot itself (first thing it does)
            fut = h.inference_pool.submit_task(hidden, hypo_ids, cache, self.lo, self.hi, block_prompts, self.active_adapter, take_from, push_to,
                                               priority=priority, size=B * T)
            out = fut.result(timeout=h.step_timeout)
        else:
            out = hidden
            for i, backend in enumerate(self.backends):
                cache.set_position(prefix)  # every block sees the same prefix; advanced once at the end
                p = [block_prompts[i]] if block_prompts is not None else None
                fut = h.inference_pool.submit_task(out, hypo_ids if i == 0 else None, cache, backend.slot, backend.slot + 1, p,
                                                   self.active_adapter, priority=priority, size=B * T)
                out = fut.result(timeout=h.step_timeout)
            cache.set_position(prefix + T)
        next_servers = metadata.get("next_servers")
        if next_servers:
            # asynchronous like the reference (handler.py:337 asyncio.create_task): the response to the client never waits for
            # the connection + transfer + acknowledgement of the server-to-server push
            h._push_pool.submit(h._push_outputs, out, metadata, next_servers)
        return out


class TransformerConnectionHandler:
    """All RPCs of one stage worker."""

    def __init__(self, swarm, module_backends: Dict[ModuleUID, TransformerBackend], *, stage: Stage, peer_id: str,
                 inference_pool: PrioritizedTaskPool, forward_pool: PrioritizedTaskPool, backward_pool: PrioritizedTaskPool,
                 adapters: Sequence[str] = (), inference_max_length: int = 8192, request_timeout: float = 3 * 60,
                 session_timeout: float = 30 * 60, step_timeout: float = 5 * 60,
                 task_prioritizer: Optional[TaskPrioritizerBase] = None, quant_type=None):
        self.swarm, self.module_backends, self.stage, self.peer_id = swarm, module_backends, stage, peer_id
        self.inference_pool, self.forward_pool, self.backward_pool = inference_pool, forward_pool, backward_pool
        self.adapters = tuple(adapters)
        self.inference_max_length = inference_max_length
        self.request_timeout, self.session_timeout, self.step_timeout = request_timeout, session_timeout, step_timeout
        self.prioritizer = task_prioritizer or DummyTaskPrioritizer()
        self.quant_type = quant_type
        self.metrics = ServerMetrics(peer_id)
        self.metrics.gauge("cache_tokens_left", lambda: stage.memory_cache.tokens_left * len(stage))
        self.metrics.gauge("queue_size", lambda: inference_pool.runtime.queue_size if getattr(inference_pool, "runtime", None) is not None else None)
        self.compression = None  # default wire codec of the responses on the socket transport (utils/compression.py)
        self._sessions: Dict[str, InferenceStream] = {}
        self._sessions_lock = threading.Lock()
        self._push_pool = concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix=f"push-{peer_id}")

    # ---- validation ------------------------------------------------------------------------------------------
    def _check_uids(self, uids) -> List[ModuleUID]:
        if isinstance(uids, str):
            uids = split_uids(uids)
        uids = list(uids)
        if not uids:
            raise RuntimeError("User must specify at least one block for inference, but got none")
        for uid in uids:
            if uid not in self.module_backends:
                raise RuntimeError(f"Remote peer does not serve {uid}")
        slots = [self.module_backends[u].slot for u in uids]
        if slots != list(range(slots[0], slots[0] + len(slots))):
            raise RuntimeError(f"requested blocks must be consecutive, got {uids}")
        return uids

    def check_adapter(self, active_adapter: Optional[str]) -> None:
        if active_adapter and active_adapter not in self.adapters:
            raise KeyError(f"adapter {active_adapter} not found (this server holds {list(self.adapters)})")

    # ---- session registry --------------------------------------------------------------------------------------
    def _register(self, stream: InferenceStream) -> None:
        with self._sessions_lock:
            self._sessions[stream.session_id] = stream
        self.metrics.session_opened()

    def _unregister(self, stream: InferenceStream) -> None:
        with self._sessions_lock:
            known = self._sessions.pop(stream.session_id, None)
        if known is not None:
            self.metrics.session_closed()

    def sweep_sessions(self) -> int:
        """Close sessions whose session/step timeout expired; returns how many were closed."""
        now = time.monotonic()
        with self._sessions_lock:
            stale = [s for s in self._sessions.values() if s.expired(now)]
        for s in stale:
            logger.info(f"closing expired inference session {s.session_id}")
            s.close()
        return len(stale)

    # ---- RPCs ----------------------------------------------------------------------------------------------------
    def rpc_inference(self, uids, metadata: Optional[Dict[str, Any]] = None) -> InferenceStream:
        return InferenceStream(self, self._check_uids(uids), metadata or {})

    def rpc_forward(self, uids, hidden: torch.Tensor, prompts: Optional[torch.Tensor] = None, metadata: Optional[Dict[str, Any]] = None) -> torch.Tensor:
        t0 = time.perf_counter()
        try:
            out = self._rpc_forward(uids, hidden, prompts, metadata)
        except Exception:
            self.metrics.error("forward")
            raise
        self.metrics.observe("forward", hidden.shape[0] * hidden.shape[1] if hidden.dim() == 3 else 0, time.perf_counter() - t0)
        return out

    def _rpc_forward(self, uids, hidden: torch.Tensor, prompts: Optional[torch.Tensor] = None, metadata: Optional[Dict[str, Any]] = None) -> torch.Tensor:
        metadata = metadata or {}
        take_from, push_to = fabric_endpoints(metadata)
        state = {"taken": take_from is None}
        try:
            maybe_fail("rpc_forward", self.peer_id)
            uids = self._check_uids(uids)
            self.check_adapter(metadata.get("active_adapter"))
            backends = [self.module_backends[u] for u in uids]
            if take_from is not None:  # the micro-batch already sits in this rank's landing slot: `hidden` is a shape carrier
                hidden = torch.empty(take_from[2], take_from[3], self.stage.spec.hidden_size, dtype=self.stage.dtype, device="meta")
            stash = self._stash_key(metadata)
            state["taken"] = True  # Stage.forward consumes the slot before anything else
            return run_rpc_forward(hidden, prompts, backends=backends, handler=self, active_adapter=metadata.get("active_adapter"),
                                   points=float(metadata.get("points", 0)), take_from=take_from, push_to=push_to, stash=stash)
        except BaseException:
            if not state["taken"]:
                drain_landing(take_from, "x_in", self.stage.device)
            raise

    def rpc_backward(self, uids, inputs: torch.Tensor, grad_outputs: torch.Tensor, prompts: Optional[torch.Tensor] = None,
                     metadata: Optional[Dict[str, Any]] = None) -> List[torch.Tensor]:
        t0 = time.perf_counter()
        try:
            out = self._rpc_backward(uids, inputs, grad_outputs, prompts, metadata)
        except Exception:
            self.metrics.error("backward")
            raise
        self.metrics.observe("backward", inputs.shape[0] * inputs.shape[1] if inputs.dim() == 3 else 0, time.perf_counter() - t0)
        return out

    def _rpc_backward(self, uids, inputs: torch.Tensor, grad_outputs: torch.Tensor, prompts: Optional[torch.Tensor] = None,
                      metadata: Optional[Dict[str, Any]] = None) -> List[torch.Tensor]:
        metadata = metadata or {}
        grad_from, push_to = fabric_endpoints(metadata)
        state = {"taken": grad_from is None}
        try:
            maybe_fail("rpc_backward", self.peer_id)
            uids = self._check_uids(uids)
            self.check_adapter(metadata.get("active_adapter"))
            backends = [self.module_backends[u] for u in uids]
            stash = self._stash_key(metadata)
            if grad_from is not None:
                grad_outputs = torch.empty(grad_from[2], grad_from[3], self.stage.spec.hidden_size, dtype=self.stage.dtype, device="meta")
            if stash is not None:  # the span input stayed on this stage since the forward (see Stage.stash_put)
                inputs = torch.empty(tuple(grad_outputs.shape), dtype=self.stage.dtype, device="meta")
                if stash not in self.stage._stash:
                    raise KeyError(f"no activations are stashed under {stash!r} (expired, evicted, or this stage never ran that forward)")
            state["taken"] = True  # Stage.backward reads (and acknowledges) the gradient slot itself
            return run_rpc_backward(inputs, grad_outputs, prompts, backends=backends, handler=self,
                                    active_adapter=metadata.get("active_adapter"), points=float(metadata.get("points", 0)),
                                    grad_from=grad_from, push_to=push_to, stash=stash)
        except BaseException:
            if not state["taken"]:
                drain_landing(grad_from, "g_in", self.stage.device)
            raise

    @staticmethod
    def _stash_key(metadata: Dict[str, Any]) -> Optional[str]:
        key = metadata.get("stash")
        if key is None:
            return None
        key = str(key)
        if not 0 < len(key) <= 128:
            raise ValueError("stash keys are 1..128 characters")
        return key

    def rpc_push(self, uids, *tensors: torch.Tensor, metadata: Optional[Dict[str, Any]] = None) -> None:
        metadata = metadata or {}
        session_id, step_id = metadata.get("session_id"), metadata.get("step_id")
        with self._sessions_lock:
            stream = self._sessions.get(session_id)
        if stream is None:
            logger.debug(f"rpc_push for unknown session {session_id} ignored")
            return
        stream.push(step_id, tuple(tensors))

    def _push_outputs(self, out: torch.Tensor, metadata: Dict[str, Any], next_servers) -> None:
        try:
            next_peer, next_session_id, start, end = next_servers[0]
            stub = self.swarm.connect(next_peer)
            prefix = next(iter(self.module_backends)).rsplit(".", 1)[0]
            next_uids = [f"{prefix}.{i}" for i in range(start, end)]
            meta = dict(session_id=next_session_id, step_id=metadata.get("step_id"), pushed=True, next_servers=next_servers[1:])
            stub.rpc_push(next_uids, out, metadata=meta)
        except Exception as e:  # noqa: BLE001 - pushing is an optimisation; the client resends anyway
            logger.debug(f"failed to push outputs to {next_servers[0]}: {e!r}")

    def rpc_check(self, check_peer: str, wait_timeout: float = 5.0) -> bool:
        """Peer-assisted reachability (reference src/petals/server/reachability.py:55-164 ``ReachabilityProtocol.rpc_check``):
        can *this* process dial ``check_peer``'s announced endpoint?  The address is looked up afresh in the registry."""
        from petals_b200.utils.ping import ping

        if hasattr(self.swarm, "forget"):
            self.swarm.forget(check_peer)
        return ping(check_peer, self.swarm, wait_timeout=wait_timeout) != float("inf")

    def rpc_ping(self) -> None:
        return None

    def rpc_info(self, uids=None) -> Dict[str, Any]:
        cache = self.stage.memory_cache
        spec = self.stage.spec
        return dict(
            version=petals_b200.__version__, dht_client_mode=False, peer_id=self.peer_id,
            **{CACHE_TOKENS_AVAILABLE: cache.tokens_left * len(self.stage)}, inference_max_length=self.inference_max_length,
            start_block=self.stage.start_block, end_block=self.stage.end_block, torch_dtype=str(self.stage.dtype).replace("torch.", ""),
            quant_type=(self.quant_type.name.lower() if self.quant_type is not None else "none"), adapters=list(self.adapters),
            keyword_names=("prompts", "hypo_ids"),
            forward_schema=dict(args=("hidden_states", "prompts"), hidden_size=spec.hidden_size),
            outputs_schema=dict(hidden_size=spec.hidden_size),
            inference_schema=dict(args=("hidden_states", "prompts", "hypo_ids"), hidden_size=spec.hidden_size),
            device=str(self.stage.device), engine="sm_100a" if self.stage.engine is not None else "oracle", metrics=self.metrics.snapshot(),
            fabric_rank=self._fabric_rank(), fabric=self._fabric_info())

    @staticmethod
    def _fabric_info() -> Optional[dict]:
        from petals_b200.parallel.fabric import fabric_info

        return fabric_info()

    @staticmethod
    def _fabric_rank() -> Optional[int]:
        from petals_b200.parallel.fabric import get_fabric

        fabric = get_fabric()
        return None if fabric is None else fabric.rank

    def shutdown(self) -> None:
        with self._sessions_lock:
            streams = list(self._sessions.values())
        for s in streams:
            s.close()
        self._push_pool.shutdown(wait=False, cancel_futures=True)

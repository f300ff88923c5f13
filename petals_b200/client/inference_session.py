"""Client side of multi-step inference.

Behavioural contract (what a user of the reference's ``InferenceSession`` relies on, src/petals/client/inference_session.py:220-414):
``step(inputs, prompts, hypo_ids)`` sends the new hidden states through every stage and returns the last stage's output;
``position`` may be moved backwards (speculative roll-back); a failing stage is blacklisted, the uncovered blocks are
re-routed and the replacement rebuilds its KV cache from the inputs its predecessor had been fed; ``max_length`` is a hard
budget (``Maximum length exceeded``); 0-token steps are legal.

Design here (one NVLink box, one process per GPU):

* every stage stream owns an :class:`InputLog` — the exact tensor sequence that stage has consumed — which is the only state
  needed to re-create the stage elsewhere;
* a step is a *wave* over the chain, driven by :meth:`InferenceSession._run_wave`: one loop with a cursor and a failure
  counter; on an error the cursor's stage is replaced (:meth:`_reroute`) and the wave continues from the same activations;
* when every hop of the chain rides the NVLink fabric (``parallel/fabric.py``) and every stream is primed, the wave is
  **dispatched to all stages at once** (:meth:`_fabric_wave`): stage *i*'s first kernel spins on its landing-zone flag and
  stage *i-1*'s last kernel fills it, so the hops are ordered on the devices, not by one client round trip per hop;
* a long step (prompt ingestion) over several stages is cut into chunks along the sequence and run as a **wavefront**
  (:meth:`_pipelined_wave`, client/pipeline.py): chunk *c* enters stage *s+1* while chunk *c+1* enters stage *s* — the stages'
  KV sessions make the causal attention of later chunks see the earlier ones, so S stages ingest a prompt concurrently instead
  of one after the other. Over the fabric the chunks travel through the landing RINGS (one slot per chunk in flight);
* activations that only ever lived in landing zones are unknown to the client; such a chain is rebuilt from the first
  stage's log as a whole.
"""
from __future__ import annotations

import time
import uuid
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from petals_b200.client.routing import RemoteSequenceManager, maybe_log_traceback
from petals_b200.data_structures import ModuleUID, RemoteSpanInfo
from petals_b200.utils.logging import get_logger
from petals_b200.utils.misc import DUMMY, DUMMY_INT64, is_dummy

logger = get_logger(__name__)
_dispatch = ThreadPoolExecutor(max_workers=16, thread_name_prefix="petals-wave")


class InputLog:
    """The inputs a stage has consumed so far, ``[B, position, H]`` (``None`` while empty). Replaying it into a fresh stream
    re-creates the stage's attention cache."""

    __slots__ = ("tokens",)

    def __init__(self, tokens: Optional[torch.Tensor] = None):
        self.tokens = tokens

    def __len__(self) -> int:
        return 0 if self.tokens is None else self.tokens.shape[1]

    def truncate(self, length: int) -> None:
        if length <= 0:
            self.tokens = None
        elif len(self) > length:
            self.tokens = self.tokens[:, :length]

    def extend(self, new: torch.Tensor) -> None:
        if new.shape[1] == 0 and self.tokens is not None:
            return
        self.tokens = new if self.tokens is None else torch.cat([self.tokens, new.to(self.tokens.device)], dim=1)

    def forget(self) -> None:
        self.tokens = None


class StageStream:
    """One ``rpc_inference`` stream: a span of blocks on one stage, its KV session there, and the log of what it was fed."""

    def __init__(self, config, span: RemoteSpanInfo, uids: Sequence[ModuleUID], stream, *, max_length: int, session_id: str):
        self.config, self.span, self.uids, self.stream = config, span, list(uids), stream
        self.max_length, self.session_id = max_length, session_id
        self.log = InputLog()
        self.cursor = 0  # tokens of this session held by the stage's cache
        self.primed = False  # has the stage executed at least one step of this stream?
        self.closed = False
        self.successor: Optional["StageStream"] = None
        self.fabric_rank: Optional[int] = None  # rank of the stage in the NVLink fabric (None: tensors travel with the RPC)
        self.fabric_info: Optional[dict] = None  # what the stage announced about that fabric: id, rank, max_tokens, hidden_size, n_slots
        self.no_history = False  # some input arrived over the fabric: the client cannot replay this stream

    @classmethod
    def open(cls, manager: RemoteSequenceManager, span: RemoteSpanInfo, *, max_length: int, alloc_timeout: float) -> "StageStream":
        uids = manager.block_uids[span.start: span.end]
        request = {k: v for k, v in manager.get_request_metadata("rpc_inference", None, *uids).items() if k != "args_structure"}
        session_id = str(uuid.uuid4())
        request.update(max_length=max_length, session_id=session_id, alloc_timeout=float(alloc_timeout))
        stub = manager.connect(span.peer_id)
        self = cls(manager.config, span, uids, stub.rpc_inference(list(uids), request), max_length=max_length, session_id=session_id)
        if manager.config.use_server_to_server:
            # which NVLink fabric (if any) the stage sits on: two stages announcing the same fabric id hop through each other's landing
            # rings whether or not this client is a member of that fabric
            try:
                info = stub.rpc_info()
                self.fabric_info = info.get("fabric")
                self.fabric_rank = self.fabric_info["rank"] if self.fabric_info else info.get("fabric_rank")
                if self.fabric_info is None and self.fabric_rank is not None:  # older stage: rank only; usable when this process is a member
                    from petals_b200.parallel.fabric import fabric_info, get_fabric

                    mine = fabric_info(get_fabric())
                    self.fabric_info = dict(mine, rank=self.fabric_rank) if mine else None
            except Exception:  # noqa: BLE001 - a stage that cannot tell simply gets its tensors with the RPC
                self.fabric_info, self.fabric_rank = None, None
        return self

    # the attributes below keep older call sites (tests, tools) readable
    @property
    def position(self) -> int:
        return self.cursor

    @property
    def history(self) -> Optional[torch.Tensor]:
        return self.log.tokens

    def rewind(self, position: int) -> None:
        if position > self.cursor:
            raise ValueError(f"a stage stream can only move backwards (at {self.cursor}, asked for {position})")
        self.cursor = position
        self.log.truncate(position)

    def _request(self, step_id: str, deliver_to: Optional[dict]) -> Dict[str, Any]:  # noqa: D401
        meta: Dict[str, Any] = {"step_id": step_id}
        if self.primed:
            meta["start_from_position"] = self.cursor  # a no-op unless the session was rolled back
        if deliver_to is not None:
            meta["fabric_out"] = deliver_to
        return meta

    def feed(self, arriving: torch.Tensor, prompts: torch.Tensor, hypo_ids: torch.Tensor, *, step_id: str, n_new: int,
             deliver_to: Optional[dict] = None):
        """Run one step with a tensor the client holds. ``arriving`` is either just the ``n_new`` new positions or — when the
        predecessor is itself rebuilding — the predecessor's output for the *whole* prefix, which lets an unprimed stream
        rebuild too. Returns the stage's output, or only its shape when the output went to a landing zone (``deliver_to``)."""
        if self.closed:
            raise RuntimeError("this stage stream is closed")
        arriving = arriving.detach()
        fresh = arriving[:, arriving.shape[1] - n_new:] if n_new else arriving[:, :0]
        if not self.primed and arriving.shape[1] == self.cursor + n_new and arriving.shape[1] > n_new:
            self.log.tokens = arriving  # the predecessor replayed everything: adopt it as this stream's log
        elif len(self.log) == self.cursor:
            self.log.extend(fresh)
        if len(self.log) != self.cursor + n_new:
            raise RuntimeError(f"input log of {self.span} holds {len(self.log)} positions, expected {self.cursor} + {n_new}")
        payload = fresh if self.primed else self.log.tokens  # an unprimed stage must see the whole prefix
        meta = self._request(step_id, deliver_to)
        if self.config.use_server_to_server and self.successor is not None:
            followers = []
            nxt = self.successor
            while nxt is not None and nxt.primed:
                followers.append((nxt.span.peer_id, nxt.session_id, nxt.span.start, nxt.span.end))
                nxt = nxt.successor
            if followers:
                meta["next_servers"] = followers
        result = self.stream.step(payload, prompts, hypo_ids, metadata=meta)
        if deliver_to is None:
            if tuple(result.shape) != tuple(payload.shape):
                raise RuntimeError(f"{self.span} returned {tuple(result.shape)} for an input of {tuple(payload.shape)}")
        else:
            result = tuple(payload.shape)
        self.primed, self.cursor = True, self.cursor + n_new
        return result

    def feed_landed(self, shape: Tuple[int, int, int], src_rank: int, prompts: torch.Tensor, hypo_ids: torch.Tensor, *, step_id: str,
                    n_new: int, deliver_to: Optional[dict] = None, slot: int = 0):
        """Run one step whose input ``src_rank`` stored (or is about to store) into landing slot ``slot`` of this stage. Returns the
        stage's output, or only its shape when the output went to a landing zone (``deliver_to``)."""
        if self.closed:
            raise RuntimeError("this stage stream is closed")
        B, L, _ = shape
        if not self.primed and L != self.cursor + n_new:
            raise RuntimeError("an unprimed stage needs the whole prefix, but only the new positions were pushed to it")
        self.log.forget()
        self.no_history = True
        meta = self._request(step_id, deliver_to)
        meta["fabric_in"] = {"src_rank": src_rank, "B": B, "T": L, "slot": slot}
        result = self.stream.step(torch.empty(0), prompts, hypo_ids, metadata=meta)
        self.primed, self.cursor = True, self.cursor + n_new
        return tuple(shape) if deliver_to is not None else result

    def close(self) -> None:
        if not self.closed:
            self.closed = True
            try:
                self.stream.close()
            except Exception as e:  # noqa: BLE001 - a dead stage cannot be closed politely
                logger.debug(f"closing the stream to {self.span.peer_id}: {e!r}")

    def __del__(self):
        self.close()


def _validate_step(inputs, prompts, hypo_ids, *, num_blocks: int, hidden_size: Optional[int], batch_size: Optional[int]):
    """Mistakes of the caller are reported on the client, once: a stage would reject them too, but for the retry machinery a
    rejection is indistinguishable from a broken stage and would be re-routed and retried."""
    if not isinstance(inputs, torch.Tensor) or inputs.ndim != 3 or not inputs.is_floating_point():
        raise ValueError("inputs must be a floating-point tensor [batch_size, seq_length, hidden_size]")
    B, T, H = inputs.shape
    if hidden_size is not None and H != hidden_size:
        raise ValueError(f"inputs have hidden size {H}, the model's is {hidden_size}")
    if batch_size is None and B < 1:
        raise ValueError("inputs must contain at least one sequence")
    if batch_size is not None and B != batch_size:
        raise ValueError(f"batch size changed within a session ({batch_size} -> {B})")
    if hypo_ids is None or is_dummy(hypo_ids):
        hypo_ids = DUMMY_INT64
    elif (hypo_ids.dtype != torch.int64 or hypo_ids.ndim != 1 or hypo_ids.shape[0] != B
          or bool(((hypo_ids < 0) | (hypo_ids >= B)).any())):
        raise ValueError(f"hypo_ids must be an int64 vector of {B} indices into the batch")
    if prompts is None or is_dummy(prompts):
        prompts = DUMMY
    else:
        ok = prompts.ndim == 4 and prompts.shape[0] == num_blocks and prompts.shape[1] in (1, B) and prompts.shape[2] <= T and prompts.shape[3] == H
        assert ok, f"deep prompts must be [num_blocks={num_blocks}, {B} or 1, prefix_len <= {T}, {H}], got {tuple(prompts.shape)}"
    return prompts, hypo_ids


class InferenceSession:
    """Multi-step inference over a chain of stages, with fail-over."""

    def __init__(self, sequence_manager: RemoteSequenceManager, max_length: int, *, alloc_timeout: float = 0.0):
        """``alloc_timeout``: how long a stage may keep this session waiting for KV-cache room before refusing it (the request
        field of the same name, reference handler.py:148-154); 0 fails fast so that routing can try another stage."""
        if isinstance(max_length, bool) or not isinstance(max_length, int) or max_length < 1:
            raise ValueError(f"max_length must be a positive number of tokens to reserve KV caches for, got {max_length!r}")
        self._manager = sequence_manager
        self._alloc_timeout = float(alloc_timeout)
        self._max_length = max_length
        self._chain: List[StageStream] = []
        self._position = 0
        self._batch_size: Optional[int] = None  # fixed by the first step
        self._closed = False
        self.output_ids: Optional[torch.Tensor] = None
        self.past_key_values = None

    # ---- simple accessors -----------------------------------------------------------------------------------------
    @property
    def _server_sessions(self) -> List[StageStream]:  # older name of the chain
        return self._chain

    @property
    def _sequence_manager(self) -> RemoteSequenceManager:
        return self._manager

    @property
    def num_blocks(self) -> int:
        return len(self._manager)

    @property
    def max_length(self) -> int:
        return self._max_length

    @property
    def position(self) -> int:
        return self._position

    @position.setter
    def position(self, target: int) -> None:
        if not 0 <= target <= self._position:
            raise ValueError(f"position can only be moved backwards within [0, {self._position}], got {target}")
        self._position = target
        for stage in self._chain:
            stage.rewind(target)

    @property
    def last_token_id(self) -> Optional[torch.Tensor]:  # petals <= 2.1 spelling
        return None if self.output_ids is None else self.output_ids[:, -1:]

    @last_token_id.setter
    def last_token_id(self, value: torch.Tensor) -> None:
        if self.output_ids is None:
            raise RuntimeError("Can't override `last_token_id` since the session has not stepped yet")
        self.output_ids[:, -1:] = value

    # ---- chain maintenance ----------------------------------------------------------------------------------------
    def _open_streams(self, spans: Sequence[RemoteSpanInfo]) -> List[StageStream]:
        opened: List[StageStream] = []
        try:
            for span in spans:
                opened.append(StageStream.open(self._manager, span, max_length=self._max_length, alloc_timeout=self._alloc_timeout))
        except BaseException:
            self._close_streams(opened)
            raise
        return opened

    @staticmethod
    def _close_streams(streams: Sequence[StageStream]) -> None:
        for stage in reversed(list(streams)):
            stage.close()

    def _relink(self) -> None:
        for a, b in zip(self._chain, self._chain[1:] + [None]):
            a.successor = b

    def _reroute(self, idx: int, frontier: int, failed_before: bool) -> None:
        """Put fresh streams in place of chain[idx] (or, past the end of the chain, route the uncovered blocks). The log of the
        replaced stream — everything that had entered block ``frontier`` — moves to the first replacement, which replays it
        on its first step and hands its full-length output on, so every replacement rebuilds its cache."""
        replaced = self._chain[idx: idx + 1]
        stop = replaced[0].span.end if replaced else self.num_blocks
        if failed_before:
            logger.debug(f"attention caches of blocks [{frontier}, {stop}) will be rebuilt on other stages")
        self._close_streams(replaced)
        spans = self._manager.make_sequence(frontier, stop, mode="min_latency", cache_tokens_needed=self._max_length)
        spans[-1].end = min(spans[-1].end, stop)  # the router may offer a stage that serves more than we asked for
        fresh = self._open_streams(spans)
        carried = replaced[0].log.tokens if replaced else None
        for k, stage in enumerate(fresh):
            stage.cursor = self._position
            if k == 0 and carried is not None and self._position > 0:
                stage.log.tokens = carried[:, : self._position]
        if self._position > 0 and fresh and len(fresh[0].log) == 0 and (idx, frontier) != (0, 0):
            self._close_streams(fresh)
            raise RuntimeError("cannot rebuild a remote attention cache: no input log for the failed span")
        self._chain[idx: idx + 1] = fresh
        self._relink()

    def _first_stage_inputs(self, step_inputs: torch.Tensor, n_new: int) -> torch.Tensor:
        """Everything block 0 has ever been fed in this session, ending with the current step's inputs."""
        head = self._chain[0].log.tokens if self._chain else None
        new = step_inputs[:, step_inputs.shape[1] - n_new:]
        if self._position == 0:
            return new
        if head is None:
            raise RuntimeError("cannot rebuild remote attention caches: the input log of the first stage is gone")
        if head.shape[1] == self._position + n_new:
            return head
        return torch.cat([head[:, : self._position], new.to(head.device)], dim=1)

    def _landing(self, fabric, idx: int, shape: Tuple[int, int, int]) -> Optional[dict]:
        """Where stage ``idx`` should store its output: the next stage's landing zone, this process's (last stage), or None =
        return it with the RPC."""
        if time.monotonic() < getattr(self._manager, "fabric_broken_until", 0.0):
            return None  # a hop failed recently (the rings may be out of step until the stages have drained them)
        here = self._chain[idx].fabric_info
        rows = shape[0] * shape[1]
        if here is None or rows == 0 or rows > min(here["max_tokens"], 4096) or shape[2] != here["hidden_size"]:
            return None
        if idx + 1 < len(self._chain):
            there = self._chain[idx + 1].fabric_info
            if there is None or there.get("id") != here.get("id") or there["rank"] == here["rank"]:
                return None  # not on the same fabric (or the same GPU): the tensor travels with the RPCs
            return {"kind": "x_in", "rank": there["rank"]}
        # the last stage can only return through a landing ring when THIS process is a member of the same fabric
        if (fabric is not None and getattr(fabric, "fabric_id", None) == here.get("id") and self._chain[idx].span.end == self.num_blocks
                and here["rank"] != fabric.rank):
            return {"kind": "y_ret", "rank": fabric.rank}
        return None

    # ---- one step --------------------------------------------------------------------------------------------------
    def step(self, inputs: torch.Tensor, prompts: Optional[torch.Tensor] = None, hypo_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert not self._closed, "the session is closed"
        if torch.is_grad_enabled():
            logger.warning("Running inference session with grad enabled. Gradients will *not* be propagated correctly.")
        prompts, hypo_ids = _validate_step(inputs, prompts, hypo_ids, num_blocks=self.num_blocks, batch_size=self._batch_size,
                                           hidden_size=getattr(self._manager.config, "hidden_size", None))
        n_new = inputs.shape[1]
        if self._position + n_new > self._max_length:
            raise ValueError(f"Maximum length exceeded: prefix {self._position} + current {n_new} exceeds pre-allocated maximum {self._max_length}")
        out = self._run_wave(inputs.detach(), prompts, hypo_ids, n_new)
        self._position += n_new
        self._batch_size = out.shape[0]
        return out[:, out.shape[1] - n_new:].to(device=inputs.device, dtype=inputs.dtype)

    def _fabric_wave(self, fabric, x: torch.Tensor, prompts, hypo_ids, n_new: int, step_id: str) -> Optional[torch.Tensor]:
        """Steady state of a chain that lives entirely on the NVLink fabric: issue this step to ALL stages at once. Returns
        None when the chain does not qualify (the caller then walks it stage by stage)."""
        chain = self._chain
        if (len(chain) < 2 or chain[0].span.start != 0 or chain[-1].span.end != self.num_blocks
                or any(not s.primed or s.closed or s.cursor != self._position for s in chain)):
            return None
        shape = tuple(x.shape)
        targets = [self._landing(fabric, i, shape) for i in range(len(chain))]
        if any(t is None for t in targets[:-1]):
            return None  # some hop between stages does not ride a fabric
        # targets[-1] is None when this process is not a member of the stages' fabric (or shares the last stage's GPU): the last stage
        # then answers with the tensor

        def span_prompts(stage: StageStream):
            return DUMMY if is_dummy(prompts) else prompts[stage.span.start: stage.span.end]

        jobs = [_dispatch.submit(chain[0].feed, x, span_prompts(chain[0]), hypo_ids, step_id=step_id, n_new=n_new, deliver_to=targets[0])]
        for i in range(1, len(chain)):
            jobs.append(_dispatch.submit(chain[i].feed_landed, shape, chain[i - 1].fabric_rank, span_prompts(chain[i]), hypo_ids,
                                         step_id=step_id, n_new=n_new, deliver_to=targets[i]))
        errors, last = [], None
        for stage, job in zip(chain, jobs):
            try:
                last = job.result()
                self._manager.on_request_success(stage.span.peer_id)
            except Exception as e:  # noqa: BLE001
                self._manager.on_request_failure(stage.span.peer_id)
                errors.append((stage, e))
        if errors:
            for stage in chain:  # landing-zone contents are unknown now: the whole chain is rebuilt from the first stage's log
                stage.no_history = True
            self._manager.fabric_broken_until = time.monotonic() + 60.0  # tensors travel with the RPCs for a while
            raise errors[0][1]
        if targets[-1] is None:
            return last
        B, L, H = shape
        return fabric.recv(B * L, "y_ret", chain[-1].fabric_rank).view(B, L, H)

    def _pipelined_wave(self, fabric, x: torch.Tensor, prompts, hypo_ids, n_new: int) -> Optional[torch.Tensor]:
        """Chunked prompt ingestion as a wavefront over the chain. Returns None when the step does not qualify (short step, one
        stage, deep prompts / beam reordering in play, chain not in step); raises when a stage fails (the caller rebuilds)."""
        from petals_b200.client.pipeline import run_wave

        chain, config = self._chain, self._manager.config
        chunk = int(getattr(config, "pipeline_chunk_tokens", 0) or 0)
        B = x.shape[0]
        if (chunk <= 0 or len(chain) < 2 or n_new < 2 * chunk or not is_dummy(prompts) or not is_dummy(hypo_ids)
                or chain[0].span.start != 0 or chain[-1].span.end != self.num_blocks
                or any(s.closed or s.cursor != self._position for s in chain) or (self._position > 0 and any(not s.primed for s in chain))):
            return None
        infos = [s.fabric_info for s in chain if s.fabric_info is not None]
        if infos:  # chunks must fit the smallest landing slot on the way; one ring slot per chunk in flight
            chunk = max(1, min(chunk, min(i["max_tokens"] for i in infos) // max(B, 1)))
        bounds = [(t0, min(n_new, t0 + chunk)) for t0 in range(0, n_new, chunk)]
        n_slots = min((i.get("n_slots", 1) for i in infos), default=1)
        step_id = str(uuid.uuid4())
        H = x.shape[2]

        class Chunk:  # one item of the wavefront
            def __init__(self, index: int, t0: int, t1: int):
                self.index, self.t0, self.t1 = index, t0, t1
                self.x: Any = None  # tensor, or (src_rank, shape) when the activations sit in the next landing slot
                self.error: Optional[BaseException] = None
                self.detached = False

        items = [Chunk(i, a, b) for i, (a, b) in enumerate(bounds)]
        for it in items:
            it.x = x[:, x.shape[1] - n_new + it.t0: x.shape[1] - n_new + it.t1]

        def stage_work(i: int):
            stage = chain[i]

            def work(it: "Chunk") -> None:
                n = it.t1 - it.t0
                shape = (B, n, H)
                target = self._landing(fabric, i, shape)
                if target is not None:
                    target = dict(target, slot=it.index % n_slots)
                sid = f"{step_id}:{it.index}"
                if isinstance(it.x, tuple):  # landed in this stage's ring by the previous stage
                    result = stage.feed_landed(shape, it.x[0], DUMMY, DUMMY_INT64, step_id=sid, n_new=n, deliver_to=target, slot=it.index % n_slots)
                else:
                    result = stage.feed(it.x, DUMMY, DUMMY_INT64, step_id=sid, n_new=n, deliver_to=target)
                it.x = (stage.fabric_rank, shape) if target is not None else result
                self._manager.on_request_success(stage.span.peer_id)
            return work

        def collect(it: "Chunk") -> None:  # last lane: bring the chunk's result home (frees the y_ret slot for a later chunk)
            if isinstance(it.x, tuple):
                n = it.t1 - it.t0
                it.x = fabric.recv(B * n, "y_ret", it.x[0], it.index % n_slots).view(B, n, H)

        run_wave(items, [stage_work(i) for i in range(len(chain))] + [collect], threaded=True)
        failed = [it for it in items if it.error is not None]
        if failed:
            for stage in chain:  # some stages are chunks ahead of others: the chain is rebuilt from the first stage's log
                stage.no_history = True
            raise failed[0].error
        return torch.cat([it.x.to(x.device, x.dtype) for it in items], dim=1)

    def _run_wave(self, step_inputs: torch.Tensor, prompts, hypo_ids, n_new: int) -> torch.Tensor:
        from petals_b200.parallel.fabric import get_fabric

        config = self._manager.config
        fabric = get_fabric() if config.use_server_to_server else None
        step_id = str(uuid.uuid4())
        x = step_inputs  # activations entering chain[idx] (held by the client unless `landed`)
        landed: Optional[Tuple[int, Tuple[int, int, int]]] = None  # (source rank, shape): x sits in chain[idx]'s landing zone
        idx = failures = 0
        try_all_at_once = True
        while True:
            frontier = self._chain[idx - 1].span.end if idx > 0 else 0
            if frontier >= self.num_blocks:
                break
            stage: Optional[StageStream] = None
            try:
                if try_all_at_once and idx == 0:
                    try_all_at_once = False
                    if not self._chain:
                        self._reroute(0, 0, False)  # first step of the session: route the whole model now
                    done = self._fabric_wave(fabric, x, prompts, hypo_ids, n_new, step_id)
                    if done is None and self._chain:
                        done = self._pipelined_wave(fabric, x, prompts, hypo_ids, n_new)
                    if done is not None:
                        return done
                if idx >= len(self._chain) or failures > 0:
                    if failures > 0 and (landed is not None or any(s.no_history for s in self._chain)):
                        # some activations only ever existed in landing zones: start over from what block 0 was fed
                        x = self._first_stage_inputs(step_inputs, n_new)
                        self._close_streams(self._chain)
                        self._chain, idx, frontier, landed = [], 0, 0, None
                    self._reroute(idx, frontier, failures > 0)
                stage = self._chain[idx]
                if stage.cursor != self._position:
                    raise RuntimeError(f"{stage.span} is at position {stage.cursor}, the session at {self._position}")
                span_prompts = DUMMY if is_dummy(prompts) else prompts[stage.span.start: stage.span.end]
                shape = landed[1] if landed is not None else tuple(x.shape)
                deliver_to = self._landing(fabric, idx, shape)
                if landed is not None:
                    result = stage.feed_landed(shape, landed[0], span_prompts, hypo_ids, step_id=step_id, n_new=n_new, deliver_to=deliver_to)
                else:
                    result = stage.feed(x, span_prompts, hypo_ids, step_id=step_id, n_new=n_new, deliver_to=deliver_to)
                if deliver_to is not None:
                    landed = (stage.fabric_rank, tuple(result))
                else:
                    x, landed = result, None
                self._manager.on_request_success(stage.span.peer_id)
                idx, failures = idx + 1, 0
            except Exception as e:  # noqa: BLE001 - whatever went wrong on a stage, try another one
                if isinstance(e, ValueError) and "Maximum length exceeded" in str(e):
                    raise
                self._manager.on_request_failure(stage.span.peer_id if stage is not None else None)
                failures += 1
                if config.max_retries is not None and failures >= config.max_retries:
                    raise
                delay = self._manager.get_retry_delay(failures - 1)
                logger.warning(f"inference step failed on {stage.span if stage is not None else 'routing'} (retry in {delay:.0f} sec): {e!r}")
                maybe_log_traceback(e)
                time.sleep(delay)
        if landed is not None:  # the last stage stored the result into this process's landing zone
            B, L, H = landed[1]
            x = fabric.recv(B * L, "y_ret", landed[0]).view(B, L, H)
        return x

    # ---- lifecycle ---------------------------------------------------------------------------------------------------
    def __enter__(self) -> "InferenceSession":
        assert not self._closed and not self._chain
        return self

    def close(self, *exc_details) -> None:
        if not getattr(self, "_closed", True):  # also safe when __init__ did not finish (reached from __del__)
            self._close_streams(self._chain)
            self._chain.clear()
            self._closed = True

    def __exit__(self, *exc_details) -> None:
        self.close()

    def __del__(self):
        self.close()

"""Pipeline-parallel forward / backward of a batch through the chain of remote stages.

What the reference offers here (src/petals/client/sequential_autograd.py:199-250): a batch is cut into micro-batches of at most
``MAX_TOKENS_IN_BATCH`` tokens, the micro-batches travel through the servers concurrently, servers are stateless between the
forward and the backward call, a failed server is replaced and its part of the forward recomputed, and the result is
differentiable w.r.t. the inputs and the deep prompts.

This module implements that as an explicit **wavefront schedule** (all-forward, then all-backward; GPipe order):

* :class:`Route` fixes the chain of spans once per pass; every stage of the route gets a :class:`_Lane` — a worker with an
  inbox — and micro-batch *i* enters stage *s+1* while micro-batch *i+1* enters stage *s*. With one stage per GPU that is the
  pipeline: S stages, M micro-batches, bubble fraction (S-1)/(M+S-1). The order on every stage is deterministic.
* every executed hop is recorded on its :class:`MicroBatch` (span + the activations that entered it): the recorded hops are
  the tape the backward wave walks in reverse, stage S-1 first.
* a stage that fails takes its micro-batch off the wavefront: :func:`finish_forward_alone` / :func:`backward_hops` complete it
  with back-off and fresh routes (for a backward hop that means recomputing the lost span's forward on the replacement
  first), while the other micro-batches keep flowing through the healthy lanes.
* when every stage of the route sits on the NVLink fabric (``parallel/fabric.py``), micro-batches and their gradients **hop from
  stage to stage through the landing rings** (:class:`FabricPlan`): stage *i* stores its output into stage *i+1*'s ``x_in`` slot from
  the epilogue of its last GEMM and keeps its own input ("stash"); in the backward wave the last kernel of stage *i+1*'s backward
  stores dL/dx into stage *i*'s ``g_in`` slot. The RPCs of such a pass carry no activations, and because a stage only *enqueues* its
  kernels before answering, the lanes dispatch a micro-batch to all stages almost at once: the order is kept on the devices by the
  landing flags. A failure on that path re-runs the micro-batch on the tensor-carrying path from the activations the client holds.
* the backward wave runs its lanes on worker threads too, except when an in-process stage would have to call the autograd
  engine for CUDA tensors from a foreign thread while this thread blocks the device's autograd worker; then the same wave
  order is executed inline.
"""
from __future__ import annotations

import os
import time
import uuid
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

from petals_b200.client.pipeline import run_wave as _run_wave
from petals_b200.client.remote_forward_backward import run_remote_backward, run_remote_forward
from petals_b200.client.routing import RemoteSequenceManager, maybe_log_traceback
from petals_b200.data_structures import RemoteSpanInfo
from petals_b200.utils.logging import get_logger
from petals_b200.utils.misc import DUMMY, is_dummy

logger = get_logger(__name__)
MAX_TOKENS_IN_BATCH = 1024
FABRIC_COOL_DOWN = 60.0  # seconds without fabric passes after a failed hop (the stages drain what they can; rings of a dead peer time out)


@dataclass
class Hop:
    """One span executed for one micro-batch: the tape entry the backward needs."""
    span: RemoteSpanInfo
    entered: Optional[torch.Tensor]  # activations that went into span.start (None: they stayed on the stage, see ``stash``)
    stash: Optional[str] = None  # key under which the stage keeps its input (fabric passes)
    stage: int = -1  # position in the FabricPlan


@dataclass
class MicroBatch:
    index: int
    x: torch.Tensor  # current activations (forward) / gradient w.r.t. them (backward)
    prompts: torch.Tensor  # [n_blocks, b or 1, pre, H] or DUMMY
    hops: List[Hop] = field(default_factory=list)
    grad_prompts: List[torch.Tensor] = field(default_factory=list)  # per executed span, in backward order
    error: Optional[BaseException] = None
    detached: bool = False  # left the wavefront (a stage failed): completed on its own
    x0: Optional[torch.Tensor] = None  # what entered the pass (forward input / output gradient): the restart point of a fabric pass
    landed: bool = False  # x currently sits in the next stage's landing slot, not here
    key: str = ""
    fwd_input: Optional[torch.Tensor] = None  # the forward input, kept for fabric passes (the stages hold everything else)
    plan: Optional[object] = None  # the FabricPlan this micro-batch's tape was recorded on

    def prompts_for(self, span: RemoteSpanInfo) -> torch.Tensor:
        return DUMMY if is_dummy(self.prompts) else self.prompts[span.start: span.end]


class Route:
    """The chain of spans a pass is scheduled on."""

    def __init__(self, manager: RemoteSequenceManager, start: int, end: int):
        self.manager = manager
        self.spans: List[RemoteSpanInfo] = list(manager.make_sequence(start, end, mode="max_throughput"))
        if not self.spans or self.spans[0].start != start or self.spans[-1].end < end:
            raise RuntimeError(f"no route covers blocks [{start}, {end})")

    def __len__(self) -> int:
        return len(self.spans)


class FabricPlan:
    """A route whose stages all sit on the NVLink fabric of this process: who pushes into whose landing ring.

    Forward of micro-batch m (ring slot m mod n_slots): the client hands x to stage 0 with the RPC (same process or Unix socket),
    stage i stores its output into ``x_in[slot]`` of stage i+1, the last stage into ``y_ret[slot]`` of this rank (or answers with
    the tensor when it shares this rank's GPU). Backward: the client stores dL/dy into ``g_in[slot]`` of the last stage, stage i
    stores dL/dx into ``g_in[slot]`` of stage i-1, stage 0 answers with the tensor. Every stage keeps its own input between the two
    calls (``stash``). One producer per (consumer, kind, slot) at any time, as the ring protocol requires."""

    hops_done = {"forward": 0, "backward": 0}  # process-wide counters (self-tests check that this path really ran)

    def __init__(self, fabric, spans: Sequence[RemoteSpanInfo], ranks: Sequence[int], n_slots: Optional[int] = None):
        # ``fabric``: this process's own membership of the stages' fabric, or None — the stages then still hop among themselves and the
        # two ends of a pass (output of the last stage, gradient into the last stage) travel with the RPCs
        self.fabric, self.spans, self.ranks = fabric, list(spans), list(ranks)
        self.n_slots = max(1, n_slots if n_slots is not None else getattr(fabric, "n_slots", 1))

    @classmethod
    def probe(cls, manager: RemoteSequenceManager, route: Optional[Route], shape: Tuple[int, int, int]) -> Optional["FabricPlan"]:
        if route is None or len(route) < 2 or os.environ.get("PETALS_B200_FABRIC_TRAINING", "1") == "0":
            return None
        if time.monotonic() < getattr(manager, "fabric_broken_until", 0.0):  # a hop failed recently: rings may be out of step, carry tensors
            return None
        from petals_b200.parallel.fabric import fabric_info, get_fabric

        fabric = get_fabric()
        mine = fabric_info(fabric)
        B, T, H = shape
        infos = []
        for span in route.spans:
            try:
                announced = manager.connect(span.peer_id).rpc_info()
            except Exception:  # noqa: BLE001 - an unreachable stage: the tensor-carrying path deals with it
                return None
            info = announced.get("fabric")
            if info is None and announced.get("fabric_rank") is not None and mine is not None:  # older stage: rank only
                info = dict(mine, rank=announced["fabric_rank"])
            if info is None or (infos and (info.get("id") != infos[-1].get("id") or info["rank"] == infos[-1]["rank"])):
                return None  # some pair of neighbours does not share a fabric (or shares a GPU)
            infos.append(info)
        if H != infos[0]["hidden_size"] or B * T == 0 or B * T > min(i["max_tokens"] for i in infos):
            return None
        member = mine is not None and mine.get("id") == infos[0].get("id")
        return cls(fabric if member else None, route.spans, [int(i["rank"]) for i in infos], n_slots=min(i.get("n_slots", 1) for i in infos))

    def slot(self, mb: "MicroBatch") -> int:
        return mb.index % self.n_slots

    def forward_hop(self, manager: RemoteSequenceManager, mb: "MicroBatch", i: int) -> None:
        span, last = self.spans[i], i == len(self.spans) - 1
        uids = manager.block_uids[span.start: span.end]
        meta = dict(manager.get_request_metadata("rpc_forward", None, *uids))
        B, T, H = mb.x0.shape
        slot, key = self.slot(mb), f"{mb.key}:{i}"
        meta["stash"] = key
        if i > 0:
            meta["fabric_in"] = {"src_rank": self.ranks[i - 1], "B": B, "T": T, "slot": slot}
        returns = last and (self.fabric is None or self.ranks[i] == self.fabric.rank)  # not a member / same GPU: the tensor comes with the answer
        if not returns:
            meta["fabric_out"] = ({"kind": "y_ret", "rank": self.fabric.rank, "slot": slot} if last
                                  else {"kind": "x_in", "rank": self.ranks[i + 1], "slot": slot})
        p = mb.prompts_for(span)
        out = manager.connect(span.peer_id).rpc_forward(list(uids), mb.x.detach() if i == 0 else torch.empty(0), None if is_dummy(p) else p, metadata=meta)
        mb.hops.append(Hop(span, None, stash=key, stage=i))
        if returns:
            mb.x, mb.landed = out, False
        elif last:
            mb.x, mb.landed = self.fabric.recv(B * T, "y_ret", self.ranks[i], slot).view(B, T, H), False
        else:
            mb.landed = True
        FabricPlan.hops_done["forward"] += 1
        manager.on_request_success(span.peer_id)

    def backward_hop(self, manager: RemoteSequenceManager, mb: "MicroBatch", hop: Hop) -> None:
        i, span = hop.stage, hop.span
        last = i == len(self.spans) - 1
        uids = manager.block_uids[span.start: span.end]
        meta = dict(manager.get_request_metadata("rpc_backward", None, *uids))
        B, T, H = mb.x0.shape
        slot = self.slot(mb)
        meta["stash"] = hop.stash
        grad = torch.empty(0)
        if last and (self.fabric is None or self.ranks[i] == self.fabric.rank):
            grad = mb.x.detach()
        else:
            if last:  # dL/dy is here: store it into the last stage's gradient slot
                self.fabric.send(mb.x.detach().reshape(B * T, H), self.ranks[i], "g_in", slot)
            meta["fabric_in"] = {"src_rank": self.fabric.rank if last else self.ranks[i + 1], "B": B, "T": T, "slot": slot}
        if i > 0:
            meta["fabric_out"] = {"kind": "g_in", "rank": self.ranks[i - 1], "slot": slot}
        p = mb.prompts_for(span)
        grads = manager.connect(span.peer_id).rpc_backward(list(uids), torch.empty(0), grad, None if is_dummy(p) else p, metadata=meta)
        if i == 0:
            mb.x, mb.landed = grads[0], False
        else:
            mb.landed = True
        if len(grads) > 1:
            mb.grad_prompts.append(grads[1])
        FabricPlan.hops_done["backward"] += 1
        manager.on_request_success(span.peer_id)


# ---- single hops ---------------------------------------------------------------------------------------------------
def _forward_hop(manager: RemoteSequenceManager, mb: MicroBatch, span: RemoteSpanInfo) -> None:
    uids = manager.block_uids[span.start: span.end]
    meta = manager.get_request_metadata("rpc_forward", None, *uids)
    (y,) = run_remote_forward(manager.connect(span.peer_id), uids, mb.x, mb.prompts_for(span), metadata=meta,
                              timeout=manager.config.request_timeout)
    mb.hops.append(Hop(span, mb.x))
    mb.x = y
    manager.on_request_success(span.peer_id)


def _backward_hop(manager: RemoteSequenceManager, mb: MicroBatch, hop: Hop) -> None:
    uids = manager.block_uids[hop.span.start: hop.span.end]
    meta = manager.get_request_metadata("rpc_backward", None, *uids)
    grads = run_remote_backward(manager.connect(hop.span.peer_id), uids, hop.entered, mb.x, mb.prompts_for(hop.span), metadata=meta,
                                timeout=manager.config.request_timeout)
    mb.x = grads[0]
    if len(grads) > 1:
        mb.grad_prompts.append(grads[1])
    manager.on_request_success(hop.span.peer_id)


def _give_up_or_wait(manager: RemoteSequenceManager, failures: int, what: str, err: BaseException) -> None:
    limit = manager.config.max_retries
    if limit is not None and failures >= limit:
        raise err
    delay = manager.get_retry_delay(failures - 1)
    logger.warning(f"{what} failed (retry in {delay:.0f} sec): {err!r}")
    maybe_log_traceback(err)
    time.sleep(delay)


# ---- off-wavefront completion (also the whole story when there is a single micro-batch and a single stage) ---------------
def finish_forward_alone(manager: RemoteSequenceManager, mb: MicroBatch, block: int, end: int, failures: int = 0) -> None:
    """Take ``mb`` from ``block`` to ``end`` on whatever stages routing offers, re-routing after each failure."""
    pending: List[RemoteSpanInfo] = []
    while block < end:
        span = None
        try:
            if not pending:
                pending = list(manager.make_sequence(block, end, mode="max_throughput"))
            span = pending.pop(0)
            _forward_hop(manager, mb, span)
            block, failures = span.end, 0
        except Exception as e:  # noqa: BLE001
            manager.on_request_failure(span.peer_id if span is not None else None)
            failures += 1
            pending = []
            _give_up_or_wait(manager, failures, f"forward of micro-batch {mb.index} via {span}", e)


def backward_hops(manager: RemoteSequenceManager, mb: MicroBatch) -> None:
    """Consume ``mb.hops`` from the back. A hop whose stage is gone is replaced by recomputing that span's forward on a fresh
    route (which may split it into several hops) and continuing with the new tail."""
    failures = 0
    while mb.hops:
        hop = mb.hops.pop()
        try:
            _backward_hop(manager, mb, hop)
            failures = 0
        except Exception as e:  # noqa: BLE001
            manager.on_request_failure(hop.span.peer_id)
            failures += 1
            _give_up_or_wait(manager, failures, f"backward of micro-batch {mb.index} via {hop.span}", e)
            redo = MicroBatch(mb.index, hop.entered, mb.prompts)
            finish_forward_alone(manager, redo, hop.span.start, hop.span.end)
            mb.hops.extend(redo.hops)  # same blocks, new stages; the activations entering them were recomputed


# ---- the wavefront (client/pipeline.py) ------------------------------------------------------------------------------------
def pipelined_forward(manager: RemoteSequenceManager, micro_batches: Sequence[MicroBatch], start: int, end: int) -> None:
    """All micro-batches through blocks [start, end); fills ``mb.x`` (outputs) and ``mb.hops`` (the tape)."""
    try:
        route: Optional[Route] = Route(manager, start, end)
    except Exception as e:  # noqa: BLE001 - routing itself can fail transiently: every micro-batch then finds its own way
        logger.debug(f"no common route ({e!r}); micro-batches are routed individually")
        route = None

    def stage_work(span: RemoteSpanInfo):
        def work(mb: MicroBatch) -> None:
            try:
                _forward_hop(manager, mb, span)
            except Exception as e:  # noqa: BLE001 - this stage is out for this micro-batch: finish it off the wavefront
                manager.on_request_failure(span.peer_id)
                mb.detached = True
                _give_up_or_wait(manager, 1, f"forward of micro-batch {mb.index} via {span}", e)
                finish_forward_alone(manager, mb, span.start, end, failures=1)
        return work

    def fabric_work(plan: FabricPlan, i: int):
        def work(mb: MicroBatch) -> None:
            try:
                plan.forward_hop(manager, mb, i)
            except Exception as e:  # noqa: BLE001 - activations in flight are lost with the hop: restart this micro-batch with tensors
                manager.fabric_broken_until = time.monotonic() + FABRIC_COOL_DOWN
                manager.on_request_failure(plan.spans[i].peer_id)
                mb.detached, mb.landed, mb.x, mb.hops = True, False, mb.x0, []
                _give_up_or_wait(manager, 1, f"fabric forward of micro-batch {mb.index} via {plan.spans[i]}", e)
                finish_forward_alone(manager, mb, start, end, failures=1)
        return work

    for mb in micro_batches:
        mb.x0, mb.fwd_input, mb.key = mb.x, mb.x, uuid.uuid4().hex
    plan = FabricPlan.probe(manager, route, tuple(micro_batches[0].x.shape)) if micro_batches else None
    if plan is not None and all(tuple(mb.x.shape[1:]) == tuple(micro_batches[0].x.shape[1:]) and mb.x.shape[0] <= micro_batches[0].x.shape[0]
                                for mb in micro_batches):
        _run_wave(micro_batches, [fabric_work(plan, i) for i in range(len(plan.spans))], threaded=len(micro_batches) > 1)
        for mb in micro_batches:
            mb.plan = plan if (mb.error is None and all(h.stash is not None for h in mb.hops)) else None
    elif route is not None:
        _run_wave(micro_batches, [stage_work(s) for s in route.spans], threaded=len(micro_batches) > 1)
    else:
        _run_wave(micro_batches, [lambda mb: finish_forward_alone(manager, mb, start, end)], threaded=len(micro_batches) > 1)
    for mb in micro_batches:
        if mb.error is not None:
            raise mb.error
        mb.detached = False


def _autograd_needs_this_thread(manager: RemoteSequenceManager, micro_batches: Sequence[MicroBatch]) -> bool:
    """An in-process stage differentiates with a re-entrant autograd call. For CUDA tensors such a call from another thread is
    queued behind the device's autograd worker — the very thread that is blocked inside our ``backward``."""
    if not any(mb.x.is_cuda for mb in micro_batches):
        return False
    from petals_b200.parallel.transport import RemoteHandlerProxy

    for mb in micro_batches:
        for hop in mb.hops:
            try:
                if not isinstance(manager.connect(hop.span.peer_id), RemoteHandlerProxy):
                    return True
            except Exception:  # noqa: BLE001 - an unreachable peer is handled (and replaced) by the hop itself
                continue
    return False


def _redo_with_tensors(manager: RemoteSequenceManager, mb: MicroBatch, plan: Optional[FabricPlan], backward: bool = True) -> None:
    """A fabric pass lost this micro-batch's activations or gradients: recompute its forward on the tensor-carrying path from the
    input the client still holds, and (``backward``) walk the new tape with the output gradient the client still holds."""
    g0 = mb.x0
    redo = MicroBatch(mb.index, mb.fwd_input, mb.prompts)
    start = min(h.span.start for h in mb.hops) if (mb.hops and plan is None) else (plan.spans[0].start if plan is not None else 0)
    end = plan.spans[-1].end if plan is not None else max(h.span.end for h in mb.hops)
    finish_forward_alone(manager, redo, start, end)
    mb.hops, mb.grad_prompts, mb.landed, mb.detached = redo.hops, [], False, True
    mb.x = g0
    if backward:
        backward_hops(manager, mb)


def pipelined_backward(manager: RemoteSequenceManager, micro_batches: Sequence[MicroBatch]) -> None:
    """Reverse wave over the recorded tapes: ``mb.x`` holds grad_outputs on entry and grad_inputs on return; ``mb.grad_prompts``
    collects per-span prompt gradients (last span first)."""
    depth = max((len(mb.hops) for mb in micro_batches), default=0)
    same_shape = all(len(mb.hops) == depth for mb in micro_batches)

    def pop_one(mb: MicroBatch) -> None:  # the lane for tape position d pops exactly one (possibly repaired) hop
        target = len(mb.hops) - 1
        hop = mb.hops.pop()
        try:
            _backward_hop(manager, mb, hop)
        except Exception as e:  # noqa: BLE001
            manager.on_request_failure(hop.span.peer_id)
            _give_up_or_wait(manager, 1, f"backward of micro-batch {mb.index} via {hop.span}", e)
            redo = MicroBatch(mb.index, hop.entered, mb.prompts)
            finish_forward_alone(manager, redo, hop.span.start, hop.span.end, failures=1)
            mb.hops.extend(redo.hops)
            while len(mb.hops) > target:  # the replacement hops of this tape position
                h = mb.hops.pop()
                try:
                    _backward_hop(manager, mb, h)
                except Exception:  # noqa: BLE001 - a second failure in a row: hand the rest to the patient serial path
                    mb.hops.append(h)
                    mb.detached = True
                    backward_hops(manager, mb)
                    return

    for mb in micro_batches:
        mb.x0 = mb.x
    plans = {id(getattr(mb, "plan", None)) for mb in micro_batches}
    plan = getattr(micro_batches[0], "plan", None) if micro_batches and len(plans) == 1 else None
    if plan is not None and same_shape and depth == len(plan.spans):
        def pop_fabric(mb: MicroBatch) -> None:
            hop = mb.hops.pop()
            try:
                plan.backward_hop(manager, mb, hop)
            except Exception as e:  # noqa: BLE001 - gradients in flight are lost with the hop: redo this micro-batch with tensors
                manager.fabric_broken_until = time.monotonic() + FABRIC_COOL_DOWN
                manager.on_request_failure(hop.span.peer_id)
                _give_up_or_wait(manager, 1, f"fabric backward of micro-batch {mb.index} via {hop.span}", e)
                _redo_with_tensors(manager, mb, plan)

        _run_wave(micro_batches, [pop_fabric] * depth, threaded=len(micro_batches) > 1)
        for mb in micro_batches:
            if mb.error is not None:
                raise mb.error
        return
    for mb in micro_batches:
        if any(h.entered is None for h in mb.hops):  # a fabric tape without its plan (mixed pass): rebuild it with tensors
            _redo_with_tensors(manager, mb, None, backward=False)
    depth = max((len(mb.hops) for mb in micro_batches), default=0)
    same_shape = all(len(mb.hops) == depth for mb in micro_batches)
    threaded = len(micro_batches) > 1 and not _autograd_needs_this_thread(manager, micro_batches)
    if same_shape and depth > 0:
        _run_wave(micro_batches, [pop_one] * depth, threaded=threaded)
    else:  # tapes of different lengths (some micro-batches were repaired in the forward): no common stage structure
        _run_wave(micro_batches, [lambda mb: backward_hops(manager, mb)], threaded=threaded)
    for mb in micro_batches:
        if mb.error is not None:
            raise mb.error
        if mb.hops:
            backward_hops(manager, mb)


# ---- autograd glue ----------------------------------------------------------------------------------------------------
def _split(inputs: torch.Tensor, prompts: Optional[torch.Tensor]) -> Tuple[List[torch.Tensor], List[torch.Tensor], bool]:
    rows = max(MAX_TOKENS_IN_BATCH // max(inputs.shape[1], 1), 1)
    xs = list(inputs.detach().split(rows))
    if prompts is None or is_dummy(prompts):
        return xs, [DUMMY] * len(xs), False
    if prompts.shape[1] == 1:  # one prompt shared by the whole batch
        return xs, [prompts.detach()] * len(xs), True
    return xs, list(prompts.detach().split(rows, dim=1)), False


class PipelinedRemoteFunction(torch.autograd.Function):
    """``outputs = blocks(inputs, prompts)`` over the swarm; differentiable w.r.t. ``inputs`` and ``prompts`` (the blocks'
    weights live on the stages and are frozen)."""

    @staticmethod
    def forward(ctx, inputs: torch.Tensor, prompts: torch.Tensor, sequence_manager: RemoteSequenceManager):
        if inputs.ndim != 3:
            raise ValueError(f"inputs must be [batch, seq, hidden], got {tuple(inputs.shape)}")
        n_blocks = len(sequence_manager.block_uids)
        if not (prompts is None or is_dummy(prompts)) and len(prompts) != n_blocks:
            raise ValueError(f"deep prompts must have one entry per block ({n_blocks}), got {len(prompts)}")
        sequence_manager.rpc_info  # noqa: B018 - resolve the schema/route before fanning out
        xs, ps, shared = _split(inputs, prompts)
        mbs = [MicroBatch(i, x, p) for i, (x, p) in enumerate(zip(xs, ps))]
        pipelined_forward(sequence_manager, mbs, 0, n_blocks)
        ctx.manager, ctx.micro_batches, ctx.shared_prompts = sequence_manager, mbs, shared
        ctx.rows = xs[0].shape[0]
        return torch.cat([mb.x.to(device=inputs.device, dtype=inputs.dtype) for mb in mbs], dim=0)

    @staticmethod
    def backward(ctx, grad_outputs: torch.Tensor):
        mbs: List[MicroBatch] = ctx.micro_batches
        pieces = grad_outputs.split(ctx.rows)
        if len(pieces) != len(mbs):
            raise RuntimeError(f"gradient has {len(pieces)} micro-batches, the forward pass ran {len(mbs)}")
        ctx.manager.rpc_info  # noqa: B018
        for mb, g in zip(mbs, pieces):
            mb.x, mb.grad_prompts, mb.error, mb.detached = g, [], None, False
        pipelined_backward(ctx.manager, mbs)
        grad_inputs = torch.cat([mb.x.to(device=grad_outputs.device, dtype=grad_outputs.dtype) for mb in mbs], dim=0)
        per_mb = []
        for mb in mbs:  # spans were visited last-to-first: restore block order along dim 0
            gp = [g.to(device=grad_outputs.device, dtype=grad_outputs.dtype) for g in reversed(mb.grad_prompts)]
            per_mb.append(torch.cat(gp, dim=0) if gp else None)
        if all(g is None for g in per_mb):
            return grad_inputs, None, None
        if ctx.shared_prompts:
            return grad_inputs, torch.stack(per_mb, 0).sum(0), None
        return grad_inputs, torch.cat(per_mb, dim=1), None

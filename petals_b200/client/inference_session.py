"""Client side of multi-step inference (reference: src/petals/client/inference_session.py:26-414).

``InferenceSession`` drives a chain of per-stage streams; each ``step`` sends the new hidden states through every
span in order and returns the final hidden states on the caller's device/dtype. Contract kept from the reference:

* fault tolerance — every stage-facing call is retried with back-off; when a stage fails, the tail of the chain is
  re-routed (``make_sequence(mode="min_latency", cache_tokens_needed=max_length)``) and the **stored input history
  is replayed** into the replacement stage so that its KV cache is rebuilt (each per-stage session keeps its full
  input history for exactly this purpose);
* ``position`` can be set backwards (speculative decoding rollback): history is truncated and the next request
  carries ``start_from_position``;
* prompts ``[n_blocks, B, pre, H]`` / ``hypo_ids [B]`` validation, 0-token steps, ``Maximum length exceeded``.

Changed for one NVLink box: tensors are *not* moved to the CPU when the stage lives in this process (CUDA hidden
states are handed over by reference), and ``next_servers`` is really attached so a stage can push its output to the
next stage (the reference's code path for that is dead).
"""
from __future__ import annotations

import itertools
import time
import uuid
from typing import List, Optional, Sequence

import torch

from petals_b200.client.routing import RemoteSequenceManager, maybe_log_traceback
from petals_b200.data_structures import CHAIN_DELIMITER, ModuleUID, RemoteSpanInfo
from petals_b200.utils.logging import get_logger
from petals_b200.utils.misc import DUMMY, DUMMY_INT64, is_dummy

logger = get_logger(__name__)


class _ServerInferenceSession:
    """One ``rpc_inference`` stream to one stage (reference :26-217)."""

    def __init__(self, config, span: RemoteSpanInfo, uids: Sequence[ModuleUID], stream, *, max_length: int, session_id: str):
        self.config, self.span, self.uids, self.stream = config, span, list(uids), stream
        self.max_length, self.session_id = max_length, session_id
        self.stepped = False
        self.closed = False
        self._position = 0
        self.history: Optional[torch.Tensor] = None  # every input this stage has seen (for fail-over replay)
        self.next_session: Optional["_ServerInferenceSession"] = None
        self.fabric_rank: Optional[int] = None  # rank of the stage in the NVLink fabric (None: tensors travel with the RPC)
        self.no_history = False  # inputs arrived over the fabric: the client never saw them

    @classmethod
    def create(cls, config, sequence_manager: RemoteSequenceManager, span: RemoteSpanInfo, uids: Sequence[ModuleUID], *,
               max_length: int, **metadata) -> "_ServerInferenceSession":
        stub = sequence_manager.connect(span.peer_id)
        session_id = str(uuid.uuid4())
        meta = dict(max_length=max_length, session_id=session_id, alloc_timeout=float(metadata.pop("alloc_timeout", 0.0)), **metadata)
        stream = stub.rpc_inference(list(uids), meta)
        session = cls(config, span, uids, stream, max_length=max_length, session_id=session_id)
        from petals_b200.parallel.fabric import get_fabric

        if get_fabric() is not None and config.use_server_to_server:
            try:
                session.fabric_rank = stub.rpc_info().get("fabric_rank")
            except Exception:  # noqa: BLE001 - a stage without fabric info simply uses the tensor path
                session.fabric_rank = None
        return session

    @property
    def position(self) -> int:
        return self._position

    @position.setter
    def position(self, start_from_position: int) -> None:
        assert start_from_position <= self._position
        self._position = start_from_position
        if self.history is not None and self.history.shape[1] >= start_from_position:
            self.history = self.history[:, :start_from_position] if start_from_position > 0 else None

    def step_pushed(self, shape: tuple, src_rank: int, prompts: torch.Tensor, hypo_ids: torch.Tensor, *, step_id: str, n_new: int,
                    fabric_out: Optional[dict]) -> torch.Tensor:
        """A step whose input already sits in the stage's landing zone (pushed by ``src_rank`` over NVLink)."""
        if self.closed:
            raise Exception("Session is closed, cannot perform step")
        B, L, H = shape
        if not self.stepped and L != self._position + n_new:
            raise RuntimeError("a fresh server session needs the full input history, but only the new tokens were pushed")
        self.history, self.no_history = None, True
        metadata = dict(step_id=step_id, fabric_in=dict(src_rank=src_rank, B=B, T=L))
        if self.stepped:
            metadata["start_from_position"] = self._position
        if fabric_out is not None:
            metadata["fabric_out"] = fabric_out
        outputs = self.stream.step(torch.empty(0), prompts, hypo_ids, metadata=metadata)
        self.stepped = True
        self._position += n_new
        return outputs

    def step(self, inputs: torch.Tensor, prompts: torch.Tensor, hypo_ids: torch.Tensor, *, step_id: str,
             n_new: Optional[int] = None, fabric_out: Optional[dict] = None) -> torch.Tensor:
        """Send the new tokens (or, on a fresh stream after fail-over, the whole history) to the stage.

        ``inputs`` may be longer than ``n_new``: a predecessor that is itself replaying hands over its full-length
        output so that this (new) session can rebuild its KV cache too. The result has the same length as what
        was actually sent; callers slice the last ``n_new`` positions."""
        if self.closed:
            raise Exception("Session is closed, cannot perform step")
        n_new = inputs.shape[1] if n_new is None else n_new
        keep = inputs.detach()
        if keep.shape[1] == self._position + n_new and (not self.stepped or self.history is None) and keep.shape[1] > n_new:
            self.history = keep  # full replay from the predecessor
        elif self.history is None:
            self.history = keep[:, -n_new:] if n_new else keep[:, :0]
        elif self.history.shape[1] == self._position:
            self.history = torch.cat([self.history, keep[:, keep.shape[1] - n_new:].to(self.history.device)], dim=1)
        assert self.history.shape[1] == self._position + n_new, \
            f"Broken input cache: span={self.span} shape={self.history.shape} position={self._position} n_input_tokens={n_new}"
        metadata = dict(step_id=step_id)
        if not self.stepped:
            to_send = self.history  # (re)build the server-side KV from everything this stage should have seen
        else:
            to_send = inputs[:, inputs.shape[1] - n_new:]
            metadata["start_from_position"] = self._position  # cheap no-op unless a rollback happened
        if self.config.use_server_to_server and self.next_session is not None:
            metadata["next_servers"] = self._collect_next_servers()
        if fabric_out is not None:
            metadata["fabric_out"] = fabric_out
        outputs = self.stream.step(to_send, prompts, hypo_ids, metadata=metadata)
        if fabric_out is None:
            assert outputs.shape == to_send.shape, f"output activation shape is different from input shape: {outputs.shape} != {to_send.shape}"
        else:
            outputs = (to_send.shape[0], to_send.shape[1], to_send.shape[2])  # lives in the next landing zone: only its shape is known here
        self.stepped = True
        self._position += n_new
        return outputs

    def _collect_next_servers(self) -> List[tuple]:
        out, s = [], self.next_session
        while s is not None and s.stepped:
            out.append((s.span.peer_id, s.session_id, s.span.start, s.span.end))
            s = s.next_session
        return out

    def close(self) -> None:
        if self.closed:
            return
        self.closed = True
        try:
            self.stream.close()
        except Exception as e:  # noqa: BLE001 - closing must never raise
            logger.debug(f"Caught exception while closing connection: {e!r}")

    def __del__(self):
        self.close()

    def __enter__(self):
        assert not self.closed
        return self

    def __exit__(self, *exc):
        self.close()


class InferenceSession:
    """Multi-step inference over a chain of stages with fail-over (reference :220-414)."""

    def __init__(self, sequence_manager: RemoteSequenceManager, max_length: int, *, alloc_timeout: float = 0.0):
        """``alloc_timeout``: how long a server may keep this session waiting for KV-cache room before refusing it (the reference's
        ``alloc_timeout`` request field, handler.py:148-154; 0 = fail fast so that routing can try another server)."""
        if isinstance(max_length, bool) or not isinstance(max_length, int) or max_length < 1:
            raise ValueError(f"max_length must be a positive number of tokens to reserve KV caches for, got {max_length!r}")
        self._sequence_manager = sequence_manager
        self._alloc_timeout = float(alloc_timeout)
        self._batch_size: Optional[int] = None  # fixed by the first step
        self._closed = False
        self._server_sessions: List[_ServerInferenceSession] = []
        self._position = 0
        self._max_length = max_length
        self.output_ids: Optional[torch.Tensor] = None
        self.past_key_values = None

    @property
    def num_blocks(self) -> int:
        return len(self._sequence_manager)

    @property
    def max_length(self) -> int:
        return self._max_length

    @property
    def position(self) -> int:
        return self._position

    @position.setter
    def position(self, start_from_position: int) -> None:
        if not 0 <= start_from_position <= self._position:
            raise ValueError(f"position can only be moved backwards within [0, {self._position}], got {start_from_position}")
        self._position = start_from_position
        for session in self._server_sessions:
            assert isinstance(session, _ServerInferenceSession)
            session.position = start_from_position

    def _enter_server_sessions(self, chosen_spans: List[RemoteSpanInfo]) -> List[_ServerInferenceSession]:
        server_sessions = []
        try:
            for span in chosen_spans:
                uids = self._sequence_manager.block_uids[span.start: span.end]
                metadata = self._sequence_manager.get_request_metadata("rpc_inference", None, *uids)
                if self._alloc_timeout > 0:
                    metadata = dict(metadata, alloc_timeout=self._alloc_timeout)
                session = _ServerInferenceSession.create(self._sequence_manager.config, self._sequence_manager, span, uids,
                                                         max_length=self._max_length, **{k: v for k, v in metadata.items() if k != "args_structure"})
                server_sessions.append(session)
            return server_sessions
        except BaseException:
            self._exit_server_sessions(server_sessions)
            raise

    def _exit_server_sessions(self, server_sessions: List[_ServerInferenceSession]) -> None:
        for session in reversed(server_sessions):
            try:
                session.close()
            except Exception:  # noqa: BLE001
                logger.debug("Caught exception while closing connection to server:", exc_info=True)

    def __enter__(self) -> "InferenceSession":
        assert not self._closed and not self._server_sessions
        return self

    def step(self, inputs: torch.Tensor, prompts: Optional[torch.Tensor] = None, hypo_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert not self._closed
        if torch.is_grad_enabled():
            logger.warning("Running inference session with grad enabled. Gradients will *not* be propagated correctly.")
        # mistakes of the caller are reported here, once: a server would reject them too, and a rejection is indistinguishable from a
        # failing server for the retry loop below (it would re-route and retry for as long as max_retries allows)
        hidden_size = getattr(self._sequence_manager.config, "hidden_size", None)
        if not isinstance(inputs, torch.Tensor) or inputs.ndim != 3 or not inputs.is_floating_point():
            raise ValueError("inputs must be a floating-point tensor [batch_size, seq_length, hidden_size]")
        if hidden_size is not None and inputs.shape[2] != hidden_size:
            raise ValueError(f"inputs have hidden size {inputs.shape[2]}, the model's is {hidden_size}")
        if self._batch_size is None:
            if inputs.shape[0] < 1:
                raise ValueError("inputs must contain at least one sequence")
        elif inputs.shape[0] != self._batch_size:
            raise ValueError(f"batch size changed within a session ({self._batch_size} -> {inputs.shape[0]})")
        if hypo_ids is not None and not is_dummy(hypo_ids) and (hypo_ids.ndim != 1 or hypo_ids.shape[0] != inputs.shape[0] or hypo_ids.dtype != torch.int64
                                                               or bool(((hypo_ids < 0) | (hypo_ids >= inputs.shape[0])).any())):
            raise ValueError(f"hypo_ids must be an int64 vector of {inputs.shape[0]} indices into the batch")
        if prompts is None or is_dummy(prompts):
            prompts = DUMMY
        else:
            assert prompts.ndim == 4, "deep prompts should have shape [num_blocks, batch_size, prefix_len, hid_size]"
            assert prompts.shape[0] == self.num_blocks
            assert prompts.shape[1] in (inputs.shape[0], 1)
            assert prompts.shape[2] <= inputs.shape[1]
            assert prompts.shape[3] == inputs.shape[2]
        if hypo_ids is None or is_dummy(hypo_ids):
            hypo_ids = DUMMY_INT64
        else:
            assert len(hypo_ids) == len(inputs)
            assert hypo_ids.dtype == torch.int64
        inputs_device, inputs_dtype = inputs.device, inputs.dtype
        step_id = str(uuid.uuid4())
        n_input_tokens = inputs.shape[1]
        if self._position + n_input_tokens > self._max_length:
            raise ValueError(f"Maximum length exceeded: prefix {self._position} + current {n_input_tokens} exceeds pre-allocated maximum {self._max_length}")

        from petals_b200.parallel.fabric import get_fabric

        fabric = get_fabric() if self._sequence_manager.config.use_server_to_server else None
        server_idx = 0
        block_idx = 0
        inputs = inputs.detach()
        step_inputs = inputs
        pushed = None  # (src_rank, (B, L, H)): the current activations live in the next stage's landing zone, not here
        while block_idx < self.num_blocks:
            for attempt_no in itertools.count():
                logger.debug(f"Inference: block {block_idx}, attempt {attempt_no}")
                server_session = None
                try:
                    if not self._server_sessions or attempt_no >= 1:
                        if attempt_no >= 1 and (pushed is not None or any(s.no_history for s in self._server_sessions)):
                            # activations that travelled over the fabric were never seen by the client: rebuild the whole
                            # chain and replay the first stage's input history through it
                            inputs = self._full_history(step_inputs, n_input_tokens)
                            self._exit_server_sessions(self._server_sessions)
                            self._server_sessions, server_idx, block_idx, pushed = [], 0, 0, None
                        self._update_sequence(server_idx, block_idx, attempt_no)
                    server_session = self._server_sessions[server_idx]
                    assert server_session.position == self._position, f"Position mismatch: {server_session.position} and {self._position}"
                    span_prompts = prompts[server_session.span.start: server_session.span.end] if not is_dummy(prompts) else DUMMY
                    fabric_out = self._fabric_target(fabric, server_idx, inputs if pushed is None else pushed[1])
                    if pushed is not None:
                        result = server_session.step_pushed(pushed[1], pushed[0], span_prompts, hypo_ids, step_id=step_id, n_new=n_input_tokens,
                                                            fabric_out=fabric_out)
                    else:
                        result = server_session.step(inputs, span_prompts, hypo_ids, step_id=step_id, n_new=n_input_tokens, fabric_out=fabric_out)
                    if fabric_out is not None:
                        shape = result if isinstance(result, tuple) else pushed[1]
                        pushed = (server_session.fabric_rank, tuple(shape))
                    else:
                        inputs, pushed = result, None
                    server_idx += 1
                    block_idx = server_session.span.end
                    self._sequence_manager.on_request_success(server_session.span.peer_id)
                    break
                except Exception as e:  # noqa: BLE001 - any stage failure triggers re-routing
                    if isinstance(e, ValueError) and "Maximum length exceeded" in str(e):
                        raise
                    self._sequence_manager.on_request_failure(server_session.span.peer_id if server_session is not None else None)
                    if self._sequence_manager.config.max_retries is not None and attempt_no + 1 >= self._sequence_manager.config.max_retries:
                        raise
                    delay = self._sequence_manager.get_retry_delay(attempt_no)
                    logger.warning(f"Caught exception when running inference via {server_session.span if server_session is not None else None} "
                                   f"(retry in {delay:.0f} sec): {e!r}")
                    maybe_log_traceback(e)
                    time.sleep(delay)
        if pushed is not None:  # the last stage stored the result into this process's landing zone
            B, L, H = pushed[1]
            inputs = fabric.recv(B * L, "y_ret", pushed[0]).view(B, L, H)
        self._position += n_input_tokens
        self._batch_size = inputs.shape[0]
        outputs = inputs[:, -n_input_tokens:]
        return outputs.to(device=inputs_device, dtype=inputs_dtype)

    def _fabric_target(self, fabric, server_idx: int, like) -> Optional[dict]:
        """Where should server ``server_idx`` store its output? None = return it with the RPC (no fabric on that hop)."""
        if fabric is None:
            return None
        cur = self._server_sessions[server_idx]
        shape = tuple(like) if isinstance(like, tuple) else tuple(like.shape)
        rows = shape[0] * shape[1]
        if cur.fabric_rank is None or rows == 0 or rows > min(fabric.max_tokens, 4096) or shape[2] != fabric.hidden_size:
            return None
        if server_idx + 1 < len(self._server_sessions):
            nxt = self._server_sessions[server_idx + 1]
            if nxt.fabric_rank is None or nxt.fabric_rank == cur.fabric_rank:
                return None
            return dict(kind="x_in", rank=nxt.fabric_rank)
        if cur.span.end == self.num_blocks and cur.fabric_rank != fabric.rank:
            return dict(kind="y_ret", rank=fabric.rank)
        return None

    def _full_history(self, step_inputs: torch.Tensor, n_new: int) -> torch.Tensor:
        """Everything the first stage has ever received in this session, ending with the current step's inputs."""
        first = self._server_sessions[0] if self._server_sessions else None
        hist = first.history if first is not None else None
        if hist is None or self._position == 0:
            if self._position > 0:
                raise RuntimeError("cannot rebuild remote attention caches: the input history of the first stage is gone")
            return step_inputs[:, step_inputs.shape[1] - n_new:]
        if hist.shape[1] == self._position + n_new:
            return hist
        return torch.cat([hist[:, : self._position], step_inputs[:, step_inputs.shape[1] - n_new:].to(hist.device)], dim=1)

    def _update_sequence(self, server_idx: int, block_idx: int, attempt_no: int) -> int:
        """Replace the chain from ``server_idx`` on with a fresh route covering the failed span (reference :364-391).

        The failed stream's input history moves to the first replacement stream (same start block), which replays
        it on its first step and hands its full-length output to the next replacement stream, and so on."""
        n_prev_spans = len(self._server_sessions)
        update_end = self._server_sessions[server_idx].span.end if server_idx < n_prev_spans else self.num_blocks
        if attempt_no >= 1:
            logger.debug(f"Due to a server failure, remote attention caches from block {block_idx} to {update_end} will be regenerated")
        old = self._server_sessions[server_idx: server_idx + 1]
        self._exit_server_sessions(old)
        updated_spans = self._sequence_manager.make_sequence(block_idx, update_end, mode="min_latency", cache_tokens_needed=self._max_length)
        updated_spans[-1].end = min(updated_spans[-1].end, update_end)  # make_sequence() could return a longer chain
        updated_sessions = self._enter_server_sessions(updated_spans)
        logger.debug(f"Found path from block {block_idx} to {update_end} via {len(updated_spans)} servers")
        for i, new_session in enumerate(updated_sessions):
            new_session._position = self._position
            if i == 0 and old and old[0].history is not None:
                new_session.history = old[0].history[:, : self._position] if self._position > 0 else None
        if self._position > 0 and updated_sessions and updated_sessions[0].history is None and not (server_idx == 0 and block_idx == 0):
            raise RuntimeError("cannot rebuild a remote attention cache: no input history for the failed span")
        self._server_sessions[server_idx: server_idx + 1] = updated_sessions
        for a, b in zip(self._server_sessions[:-1], self._server_sessions[1:]):
            a.next_session = b
        if self._server_sessions:
            self._server_sessions[-1].next_session = None
        return len(self._server_sessions) - n_prev_spans

    def close(self, *exc_details) -> None:
        if not getattr(self, "_closed", True):  # also safe when __init__ did not finish (called from __del__)
            self._exit_server_sessions(self._server_sessions)
            self._server_sessions.clear()
            self._closed = True

    def __exit__(self, *exc_details):
        self.close(*exc_details)

    def __del__(self):
        self.close()

    @property
    def last_token_id(self) -> Optional[torch.Tensor]:  # backward compatibility with petals<=2.1
        return self.output_ids[:, -1:] if self.output_ids is not None else None

    @last_token_id.setter
    def last_token_id(self, value: torch.Tensor) -> None:
        if self.output_ids is None:
            raise RuntimeError("Can't override `last_token_id` since the session has not stepped yet")
        self.output_ids[:, -1:] = value

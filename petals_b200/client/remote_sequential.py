"""``RemoteSequential``: the transformer blocks of a model, as an ``nn.Module`` whose layers live on stage workers.

Call surface of the reference (src/petals/client/remote_sequential.py:20-111): differentiable ``forward`` outside a session,
a KV-cached step inside ``with seq.inference_session(max_length=...)``, ``seq[i]`` / ``seq[a:b]`` views, ``use_session``,
``active_session``, ``position``.  The implementation is organised around one small object, :class:`_SessionScope`, that owns
the "which session is active in this context" state, so nested / concurrent contexts (threads, asyncio tasks) each see
their own session.
"""
from __future__ import annotations

import contextlib
import contextvars
from typing import Iterator, Optional, Union

import torch
from torch import nn

from petals_b200.client.config import ClientConfig
from petals_b200.client.inference_session import InferenceSession
from petals_b200.client.routing import RemoteSequenceManager
from petals_b200.client.sequential_autograd import PipelinedRemoteFunction
from petals_b200.data_structures import make_uid
from petals_b200.parallel.swarm import Swarm
from petals_b200.utils.logging import get_logger
from petals_b200.utils.misc import DUMMY

logger = get_logger(__name__)


class _SessionScope:
    """Context-local pointer to the inference session that ``forward`` should step."""

    def __init__(self):
        self._var: contextvars.ContextVar = contextvars.ContextVar("petals_b200_active_session", default=None)

    @property
    def current(self) -> Optional[InferenceSession]:
        return self._var.get()

    @contextlib.contextmanager
    def bound_to(self, session: Optional[InferenceSession]) -> Iterator[Optional[InferenceSession]]:
        token = self._var.set(session)
        try:
            yield session
        finally:
            self._var.reset(token)


def _span_uids(config: ClientConfig, start_block: Optional[int], end_block: Optional[int]) -> tuple:
    first = 0 if start_block is None else start_block
    last = config.num_hidden_layers if end_block is None else end_block
    return tuple(make_uid(config.dht_prefix, index) for index in range(first, last))


class RemoteSequential(nn.Module):
    """A span of blocks served by the swarm.

    * no session active: ``forward(inputs[, prompts])`` pipelines micro-batches through the stages and is differentiable
      with respect to ``inputs`` and the deep ``prompts`` (``[n_blocks, batch, pre_seq_len, hidden]``);
    * inside ``inference_session`` / ``use_session``: ``forward`` is one step of a fault-tolerant session whose KV caches
      live on the stages."""

    def __init__(self, config: ClientConfig, *, sequence_manager: Optional[RemoteSequenceManager] = None, dht: Optional[Swarm] = None,
                 start_block: Optional[int] = None, end_block: Optional[int] = None, **kwargs):
        super().__init__()
        self.config = config
        if sequence_manager is not None:
            assert dht is None and start_block is None and end_block is None, \
                "`dht`, `start_block`, and `end_block` have no effect when you provide a custom `sequence_manager`"
            self.sequence_manager = sequence_manager
        else:
            self.sequence_manager = RemoteSequenceManager(config, _span_uids(config, start_block, end_block), dht=dht, **kwargs)
        self._scope = _SessionScope()

    # ---- execution ---------------------------------------------------------------------------------------------------------
    def forward(self, inputs: torch.Tensor, prompts: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
        assert inputs.ndim == 3, "inputs must be a tensor of shape [batch_size, seq_length, hidden_size]"
        session = self._scope.current
        if session is not None:
            return session.step(inputs, prompts, **kwargs)
        unsupported = {name: value for name, value in kwargs.items() if value is not None}
        assert not unsupported, f"Extra kwargs are not supported in forward: {unsupported}"
        # caller mistakes are reported here: a server's rejection would look like a failing server to the retry logic
        hidden_size = getattr(self.config, "hidden_size", None)
        if not inputs.is_floating_point() or (hidden_size is not None and inputs.shape[-1] != hidden_size):
            raise ValueError(f"inputs must be floating-point hidden states of size {hidden_size}, got {inputs.dtype} {tuple(inputs.shape)}")
        if inputs.shape[0] == 0 or inputs.shape[1] == 0:
            raise ValueError(f"inputs must contain at least one token, got shape {tuple(inputs.shape)}")
        if prompts is not None and prompts.numel() and (prompts.ndim != 4 or prompts.shape[0] != len(self) or prompts.shape[1] not in (1, inputs.shape[0])
                                                         or prompts.shape[2] > inputs.shape[1] or prompts.shape[3] != inputs.shape[2]):
            raise ValueError(f"deep prompts must be [{len(self)}, {inputs.shape[0]} or 1, <= {inputs.shape[1]}, {inputs.shape[2]}], got {tuple(prompts.shape)}")
        return PipelinedRemoteFunction.apply(inputs, DUMMY if prompts is None else prompts, self.sequence_manager)

    # ---- sessions ------------------------------------------------------------------------------------------------------------
    @property
    def active_session(self) -> Optional[InferenceSession]:
        """The session ``forward`` steps in the current context (thread / task), if any."""
        return self._scope.current

    @property
    def position(self) -> int:
        return self._scope.current.position

    def use_session(self, session: Optional[InferenceSession]):
        """``with seq.use_session(sess):`` routes ``forward`` through an existing session (``None``: outside of any)."""
        return self._scope.bound_to(session)

    @contextlib.contextmanager
    def inference_session(self, **kwargs) -> Iterator[InferenceSession]:
        """``with seq.inference_session(max_length=N) as sess:`` opens a session, makes it active, closes it on exit."""
        assert self._scope.current is None, "Already in an inference session"
        with InferenceSession(self.sequence_manager, **kwargs) as session, self._scope.bound_to(session):
            yield session

    # ---- container protocol -----------------------------------------------------------------------------------------------------
    def __len__(self) -> int:
        return len(self.sequence_manager)

    def __getitem__(self, ix: Union[int, slice]) -> "RemoteSequential":
        """A view over a sub-span; it shares the routing state (known servers, bans, pings) with its parent."""
        return type(self)(self.config, sequence_manager=self.sequence_manager[ix])

    def __iter__(self) -> Iterator["RemoteSequential"]:
        return (self[index] for index in range(len(self)))

    def extra_repr(self) -> str:
        uids = self.sequence_manager.block_uids
        return f"modules={uids[0]}..{uids[-1]}"

"""``RemoteSequential``: an ``nn.Module`` whose layers are transformer blocks served by stage workers
(reference: src/petals/client/remote_sequential.py:20-111)."""
from __future__ import annotations

from contextlib import contextmanager
from contextvars import ContextVar
from typing import Optional, Union

import torch
from torch import nn

from petals_b200.client.config import ClientConfig
from petals_b200.client.inference_session import InferenceSession
from petals_b200.client.routing import RemoteSequenceManager
from petals_b200.client.sequential_autograd import _RemoteSequentialAutogradFunction
from petals_b200.data_structures import UID_DELIMITER, make_uid
from petals_b200.parallel.swarm import Swarm
from petals_b200.utils.logging import get_logger
from petals_b200.utils.misc import DUMMY

logger = get_logger(__name__)


class RemoteSequential(nn.Module):
    """A sequence of transformer blocks hosted by the swarm of stages.

    Outside of an inference session, ``forward`` is differentiable (w.r.t. inputs and deep prompts) and pipelines
    micro-batches through the stages; inside ``with seq.inference_session(max_length=...)`` it is a step of a
    fault-tolerant server-side-KV session. ``seq[a:b]`` / ``seq[i]`` are views that share routing state."""

    def __init__(self, config: ClientConfig, *, sequence_manager: Optional[RemoteSequenceManager] = None, dht: Optional[Swarm] = None,
                 start_block: Optional[int] = None, end_block: Optional[int] = None, **kwargs):
        super().__init__()
        self.config = config
        assert sequence_manager is None or (dht is None and start_block is None and end_block is None), \
            "`dht`, `start_block`, and `end_block` have no effect when you provide a custom `sequence_manager`"
        if sequence_manager is None:
            if start_block is None:
                start_block = 0
            if end_block is None:
                end_block = self.config.num_hidden_layers
            block_uids = tuple(make_uid(config.dht_prefix, i) for i in range(start_block, end_block))
            sequence_manager = RemoteSequenceManager(config, block_uids, dht=dht, **kwargs)
        self.sequence_manager = sequence_manager
        self._active_session: ContextVar = ContextVar("active_session", default=None)

    def forward(self, inputs: torch.Tensor, prompts: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
        assert inputs.ndim == 3, "inputs must be a tensor of shape [batch_size, seq_length, hidden_size]"
        if self.active_session is None:
            assert all(v is None for v in kwargs.values()), f"Extra kwargs are not supported in forward: {kwargs}"
            return _RemoteSequentialAutogradFunction.apply(inputs, prompts if prompts is not None else DUMMY, self.sequence_manager)
        return self.active_session.step(inputs, prompts, **kwargs)

    @property
    def active_session(self) -> Optional[InferenceSession]:
        """The session used by ``forward`` (set by ``inference_session`` / ``use_session``), per context."""
        return self._active_session.get()

    @property
    def position(self) -> int:
        return self.active_session.position

    @contextmanager
    def use_session(self, session: Optional[InferenceSession]) -> InferenceSession:
        """Run ``forward`` calls through an existing session (or, with ``None``, outside of any session)."""
        token = self._active_session.set(session)
        try:
            yield session
        finally:
            self._active_session.reset(token)

    @contextmanager
    def inference_session(self, **kwargs) -> InferenceSession:
        """``with seq.inference_session(max_length=N) as sess:`` — creates a session and makes it active."""
        assert self.active_session is None, "Already in an inference session"
        with InferenceSession(self.sequence_manager, **kwargs) as session:
            token = self._active_session.set(session)
            try:
                yield session
            finally:
                self._active_session.reset(token)

    def __getitem__(self, ix: Union[int, slice]) -> "RemoteSequential":
        return RemoteSequential(self.config, sequence_manager=self.sequence_manager[ix])

    def __iter__(self):
        for block_index in range(len(self)):
            yield self[block_index]

    def __len__(self) -> int:
        return len(self.sequence_manager)

    def extra_repr(self) -> str:
        return f"modules={self.sequence_manager.block_uids[0]}..{self.sequence_manager.block_uids[-1]}"

"""Snapshot of the block -> servers map (reference: src/petals/client/routing/sequence_info.py:13-67)."""
from __future__ import annotations

import dataclasses
import time
from typing import Iterable, List, Optional, Sequence, Tuple, Type, TypeVar

from petals_b200.data_structures import ModuleUID, RemoteModuleInfo, RemoteSpanInfo, ServerState
from petals_b200.utils.dht import compute_spans

T = TypeVar("T")


@dataclasses.dataclass(frozen=True)
class RemoteSequenceInfo:
    """Mutable-in-place (via ``update_``) view shared by a sequence manager and its slices."""

    block_uids: Tuple[ModuleUID, ...]
    block_infos: Tuple[RemoteModuleInfo, ...]
    spans_by_priority: List[RemoteSpanInfo]  # longest spans first
    spans_containing_block: Tuple[List[RemoteSpanInfo], ...]
    last_updated_time: Optional[float]

    @classmethod
    def make_empty(cls: Type[T], block_uids: Iterable[ModuleUID]) -> T:
        block_uids = tuple(block_uids)
        empty_infos = tuple(RemoteModuleInfo(uid, {}) for uid in block_uids)
        return cls(block_uids, empty_infos, [], tuple([] for _ in block_uids), last_updated_time=None)

    def __getitem__(self, ix: slice) -> "RemoteSequenceInfo":
        assert isinstance(ix, slice)
        block_uids, block_infos = self.block_uids[ix], self.block_infos[ix]
        spans_by_priority, spans_containing_block = self.compute_spans(block_infos)
        return RemoteSequenceInfo(block_uids, block_infos, spans_by_priority, spans_containing_block, self.last_updated_time)

    def __len__(self) -> int:
        return len(self.block_uids)

    def update_(self, new_block_infos: Sequence[RemoteModuleInfo]) -> None:
        assert len(new_block_infos) == len(self.block_uids)
        for i, (uid, info) in enumerate(zip(self.block_uids, new_block_infos)):
            assert info.uid == uid, f"block {i}: expected {uid}, got {info.uid}"
            self.block_infos[i].servers = info.servers
        spans_by_priority, spans_containing_block = self.compute_spans(self.block_infos)
        object.__setattr__(self, "spans_by_priority", spans_by_priority)
        object.__setattr__(self, "spans_containing_block", spans_containing_block)
        object.__setattr__(self, "last_updated_time", time.perf_counter())

    @staticmethod
    def compute_spans(block_infos: Sequence[RemoteModuleInfo]):
        spans = compute_spans(block_infos, min_state=ServerState.ONLINE)
        by_priority = sorted(spans.values(), key=lambda s: s.length, reverse=True)
        containing = tuple([] for _ in block_infos)
        for span in by_priority:
            for i in range(span.start, span.end):
                containing[i].append(span)
        return by_priority, containing

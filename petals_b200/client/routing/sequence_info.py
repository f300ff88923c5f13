"""The routing table of a client: which peers serve which blocks right now
(reference: src/petals/client/routing/sequence_info.py:13-67 — same attribute names, since the sequence manager and user code
read them: ``block_uids``, ``block_infos``, ``spans_by_priority``, ``spans_containing_block``, ``last_updated_time``).

One table is shared by a :class:`RemoteSequenceManager` and all views sliced from it; a refresh replaces the per-block server
dictionaries in place and re-derives the two span indexes, so every holder sees the new state without re-linking.
"""
from __future__ import annotations

import time
from typing import Iterable, List, Optional, Sequence, Tuple

from petals_b200.data_structures import ModuleUID, RemoteModuleInfo, RemoteSpanInfo, ServerState
from petals_b200.utils.dht import compute_spans


def _index_spans(block_infos: Sequence[RemoteModuleInfo]) -> Tuple[List[RemoteSpanInfo], Tuple[List[RemoteSpanInfo], ...]]:
    """-> (ONLINE spans, longest first; for every block the spans that cover it, longest first)."""
    online = compute_spans(block_infos, min_state=ServerState.ONLINE)
    ranked = sorted(online.values(), key=lambda span: -span.length)
    covering: Tuple[List[RemoteSpanInfo], ...] = tuple([] for _ in block_infos)
    for span in ranked:
        for block in range(span.start, span.end):
            covering[block].append(span)
    return ranked, covering


class RemoteSequenceInfo:
    __slots__ = ("block_uids", "block_infos", "spans_by_priority", "spans_containing_block", "last_updated_time")

    def __init__(self, block_uids: Iterable[ModuleUID], block_infos: Optional[Sequence[RemoteModuleInfo]] = None,
                 last_updated_time: Optional[float] = None):
        self.block_uids: Tuple[ModuleUID, ...] = tuple(block_uids)
        if block_infos is None:
            block_infos = [RemoteModuleInfo(uid, {}) for uid in self.block_uids]
        if len(block_infos) != len(self.block_uids):
            raise ValueError("one RemoteModuleInfo per block uid is required")
        self.block_infos: Tuple[RemoteModuleInfo, ...] = tuple(block_infos)
        self.last_updated_time = last_updated_time
        self.spans_by_priority, self.spans_containing_block = _index_spans(self.block_infos)

    @classmethod
    def make_empty(cls, block_uids: Iterable[ModuleUID]) -> "RemoteSequenceInfo":
        """A table that knows the uids but no servers yet (``last_updated_time`` is None until the first refresh)."""
        return cls(block_uids)

    def __len__(self) -> int:
        return len(self.block_uids)

    def __getitem__(self, ix: slice) -> "RemoteSequenceInfo":
        """The table of a sub-sequence: it shares the per-block records (a refresh of the parent shows through) and has its own
        span indexes, numbered from the slice's first block."""
        if not isinstance(ix, slice):
            raise TypeError("RemoteSequenceInfo can only be sliced")
        return type(self)(self.block_uids[ix], self.block_infos[ix], self.last_updated_time)

    def update_(self, new_block_infos: Sequence[RemoteModuleInfo]) -> None:
        """Install fresh records (same blocks, same order) and re-derive the span indexes."""
        if len(new_block_infos) != len(self.block_uids):
            raise ValueError(f"expected {len(self.block_uids)} records, got {len(new_block_infos)}")
        for position, (mine, fresh) in enumerate(zip(self.block_infos, new_block_infos)):
            if fresh.uid != mine.uid:
                raise ValueError(f"record {position} is for {fresh.uid}, expected {mine.uid}")
            mine.servers = fresh.servers
        self.spans_by_priority, self.spans_containing_block = _index_spans(self.block_infos)
        self.last_updated_time = time.perf_counter()

    @staticmethod
    def compute_spans(block_infos: Sequence[RemoteModuleInfo]):
        return _index_spans(block_infos)

    def __repr__(self) -> str:
        served = sum(1 for spans in self.spans_containing_block if spans)
        return f"RemoteSequenceInfo({len(self)} blocks, {served} served, {len(self.spans_by_priority)} spans)"

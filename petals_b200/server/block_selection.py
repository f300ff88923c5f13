"""Which blocks should a joining / idle stage serve? (reference: src/petals/server/block_selection.py:1-95).

The policy is the reference's, because mixed swarms must agree on it:

* **join**: take the contiguous window whose blocks are currently worst served — compare windows by their sorted per-block
  throughputs (so the weakest block decides, then the second weakest, ...), earliest window on ties;
* **re-balance**: a stage moves when, after it re-joins at its best window and every other stage is allowed to react the same way
  until nobody wants to move, the swarm's bottleneck throughput would rise by more than a factor ``1 / balance_quality``.

The implementation keeps the aggregate in a small :class:`_Coverage` object (per-block throughput with stages lifted out and
dropped back in) instead of threading arrays through helper functions.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from petals_b200.data_structures import PeerID, RemoteModuleInfo, RemoteSpanInfo, ServerState
from petals_b200.utils.dht import compute_spans
from petals_b200.utils.logging import get_logger

logger = get_logger(__name__)

_EPS = 1e-3  # a lifted stage leaves a trace of this relative size so that an empty window still prefers its old place


class _Coverage:
    """Aggregate advertised throughput of every block (JOINING stages count, so simultaneous joiners spread out)."""

    def __init__(self, spans: Dict[PeerID, RemoteSpanInfo], total_blocks: int):
        self.per_block = np.zeros(total_blocks, dtype=np.float64)
        for span in spans.values():
            self.add(span)

    def add(self, span: RemoteSpanInfo, weight: float = 1.0) -> None:
        self.per_block[span.start: span.end] += weight * span.throughput

    def bottleneck(self) -> float:
        return float(self.per_block.min())

    def weakest_window(self, length: int) -> int:
        """Start of the window of ``length`` blocks that is lexicographically worst served."""
        last_start = len(self.per_block) - length
        keyed = [(sorted(self.per_block[start: start + length]), start) for start in range(last_start + 1)]
        return min(keyed)[1]

    def relocate(self, span: RemoteSpanInfo) -> bool:
        """Lift ``span`` out (leaving an epsilon trace), drop it onto the weakest window; True if it ended up elsewhere."""
        self.add(span, -(1.0 + _EPS))
        target = self.weakest_window(span.length)
        self.add(span, _EPS)  # remove the trace at the old place
        moved = target != span.start
        if moved:
            span.start, span.end = target, target + span.length
        self.add(span)
        return moved


def compute_throughputs(spans: Dict[PeerID, RemoteSpanInfo], *, total_blocks: int) -> np.ndarray:
    return _Coverage(spans, total_blocks).per_block


def choose_best_blocks(num_blocks: int, module_infos: List[RemoteModuleInfo]) -> List[int]:
    spans = compute_spans(module_infos, min_state=ServerState.JOINING)
    start = _Coverage(spans, len(module_infos)).weakest_window(num_blocks)
    return list(range(start, start + num_blocks))


def should_choose_other_blocks(local_peer_id: PeerID, module_infos: List[RemoteModuleInfo], balance_quality: float) -> bool:
    if balance_quality > 1.0:
        return True  # debugging aid, same as the reference: force a move at every check
    spans = compute_spans(module_infos, min_state=ServerState.JOINING)
    mine = spans.get(local_peer_id)
    if mine is None:
        return False
    coverage = _Coverage(spans, len(module_infos))
    before = coverage.bottleneck()

    # would leaving cut the chain? (only matters if the chain is whole now)
    without_me = coverage.per_block.copy()
    without_me[mine.start: mine.end] -= mine.throughput * (1.0 + _EPS)
    if before > _EPS and without_me.min() <= 0:
        return False
    if not coverage.relocate(mine):
        return False  # already at the weakest window
    # everybody else reacts greedily (shortest spans first) until a fixed point
    while any([coverage.relocate(spans[peer]) for peer in sorted(spans, key=lambda p: spans[p].length)]):
        pass
    after = coverage.bottleneck()
    if after < before or after < _EPS:
        return False
    quality = before / after
    logger.info(f"Swarm balance quality: {quality * 100:.1f}%")
    return quality < balance_quality - _EPS

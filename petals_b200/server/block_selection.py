"""Which blocks should a joining / idle stage serve? (reference: src/petals/server/block_selection.py:1-95).

Same policy as the reference — serve the contiguous window whose blocks currently have the least aggregate
throughput, and move when a greedy re-placement of all servers would improve the swarm's bottleneck throughput by
more than ``1 / balance_quality`` — computed on the static in-box registry instead of DHT snapshots."""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from petals_b200.data_structures import PeerID, RemoteModuleInfo, RemoteSpanInfo, ServerState
from petals_b200.utils.dht import compute_spans
from petals_b200.utils.logging import get_logger

logger = get_logger(__name__)


def compute_throughputs(spans: Dict[PeerID, RemoteSpanInfo], *, total_blocks: int) -> np.ndarray:
    """Aggregate advertised throughput per block. JOINING servers count so that simultaneous joiners spread out."""
    throughputs = np.zeros(total_blocks)
    for span in spans.values():
        throughputs[span.start: span.end] += span.throughput
    return throughputs


def _choose_best_start(throughputs: np.ndarray, num_blocks: int) -> int:
    options = [(sorted(throughputs[i: i + num_blocks]), i) for i in range(0, len(throughputs) - num_blocks + 1)]
    return min(options)[-1]


def choose_best_blocks(num_blocks: int, module_infos: List[RemoteModuleInfo]) -> List[int]:
    spans = compute_spans(module_infos, min_state=ServerState.JOINING)
    throughputs = compute_throughputs(spans, total_blocks=len(module_infos))
    start = _choose_best_start(throughputs, num_blocks)
    return list(range(start, start + num_blocks))


def _move_span(span: RemoteSpanInfo, new_start: int) -> None:
    span.start, span.end = new_start, new_start + span.length


def should_choose_other_blocks(local_peer_id: PeerID, module_infos: List[RemoteModuleInfo], balance_quality: float) -> bool:
    if balance_quality > 1.0:
        return True  # debugging aid, same as the reference: force a move at every check
    spans = compute_spans(module_infos, min_state=ServerState.JOINING)
    throughputs = compute_throughputs(spans, total_blocks=len(module_infos))
    initial = throughputs.min()
    eps = 1e-3
    if local_peer_id not in spans:
        return False
    local = spans[local_peer_id]
    throughputs[local.start: local.end] -= local.throughput * (1 + eps)
    if initial > eps and throughputs.min() <= 0:
        return False  # moving away would disconnect the pipeline
    new_start = _choose_best_start(throughputs, local.length)
    if local.start == new_start:
        return False
    throughputs[local.start: local.end] += local.throughput * eps
    _move_span(local, new_start)
    throughputs[local.start: local.end] += local.throughput
    moved = True
    while moved:  # let every other server react greedily until a fixed point
        moved = False
        for peer_id in sorted(spans, key=lambda p: spans[p].length):
            span = spans[peer_id]
            throughputs[span.start: span.end] -= span.throughput * (1 + eps)
            best = _choose_best_start(throughputs, span.length)
            throughputs[span.start: span.end] += span.throughput * eps
            if span.start != best:
                _move_span(span, best)
                moved = True
            throughputs[span.start: span.end] += span.throughput
    new = throughputs.min()
    if new < initial or new < eps:
        return False
    actual_quality = initial / new
    logger.info(f"Swarm balance quality: {actual_quality * 100:.1f}%")
    return actual_quality < balance_quality - eps

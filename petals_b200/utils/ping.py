"""Round-trip-time bookkeeping (reference: src/petals/utils/ping.py:1-64): EMA (alpha 0.2) of measured RTTs
with expiry. On one box an RTT is the control-channel ping of a peer (micro- to milli-seconds); the data-plane
latency of the NVLink hop is measured separately by parallel/symmetric.py."""
from __future__ import annotations

import math
import threading
import time
from typing import Dict, Sequence

from petals_b200.parallel.swarm import Swarm
from petals_b200.utils.logging import get_logger

logger = get_logger(__name__)


def ping(peer_id: str, swarm: Swarm, *, wait_timeout: float = 5.0) -> float:
    try:
        stub = swarm.connect(peer_id, connect_timeout=wait_timeout, request_timeout=wait_timeout)
        t0 = time.perf_counter()
        if hasattr(stub, "rpc_ping"):
            stub.rpc_ping()
        else:
            stub.rpc_info()
        return time.perf_counter() - t0
    except Exception as e:  # noqa: BLE001
        logger.debug(f"ping of {peer_id} failed: {e}")
        return math.inf


def ping_parallel(peer_ids: Sequence[str], swarm: Swarm, **kwargs) -> Dict[str, float]:
    return {p: ping(p, swarm, **kwargs) for p in peer_ids}


class PingAggregator:
    def __init__(self, swarm: Swarm, *, ema_alpha: float = 0.2, expiration: float = 300):
        self.swarm, self.ema_alpha, self.expiration = swarm, ema_alpha, expiration
        self._rtts: Dict[str, tuple] = {}
        self._lock = threading.Lock()

    def ping(self, peer_ids: Sequence[str], **kwargs) -> None:
        current = ping_parallel(peer_ids, self.swarm, **kwargs)
        now = time.time()
        with self._lock:
            for peer, rtt in current.items():
                if not math.isfinite(rtt):
                    continue
                prev = self._rtts.get(peer)
                if prev is not None and prev[1] > now:
                    rtt = self.ema_alpha * rtt + (1 - self.ema_alpha) * prev[0]
                self._rtts[peer] = (rtt, now + self.expiration)

    def to_dict(self) -> Dict[str, float]:
        now = time.time()
        with self._lock:
            return {p: r for p, (r, exp) in self._rtts.items() if exp > now}

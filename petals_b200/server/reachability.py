"""Reachability checks (reference: src/petals/server/reachability.py:1-164).

The reference verifies that a server is reachable from the public Internet (health API + peer-assisted
``rpc_check``). Inside one box the equivalent question is "can every other stage reach my control endpoint and,
on GPUs, does peer-to-peer memory access work over NVLink?"."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from petals_b200.parallel.swarm import Swarm
from petals_b200.utils.logging import get_logger
from petals_b200.utils.ping import ping

logger = get_logger(__name__)


def validate_reachability(peer_id: str, swarm: Swarm, wait_time: float = 5.0) -> None:
    rtt = ping(peer_id, swarm, wait_timeout=wait_time)
    if rtt == float("inf"):
        raise RuntimeError(f"stage {peer_id} registered itself but its control endpoint does not answer")


def check_direct_reachability(max_peers: int = 5, threshold: float = 0.5, **kwargs) -> Optional[bool]:
    """Always directly reachable: there are no NATs or relays between GPUs of one box."""
    return True


def check_p2p_access(devices=None) -> Dict[str, bool]:
    """NVLink peer access matrix between visible GPUs (the data-plane analogue of reachability)."""
    out: Dict[str, bool] = {}
    if not torch.cuda.is_available():
        return out
    n = torch.cuda.device_count()
    devices = list(range(n)) if devices is None else list(devices)
    for a in devices:
        for b in devices:
            if a != b:
                out[f"{a}->{b}"] = bool(torch.cuda.can_device_access_peer(a, b))
    return out


class ReachabilityProtocol:
    """Kept for API parity: answers whether a given peer can be reached from this process."""

    def __init__(self, swarm: Swarm, *, wait_timeout: float = 5.0):
        self.swarm, self.wait_timeout = swarm, wait_timeout

    def call_check(self, remote_peer: str, *, check_peer: str) -> Optional[bool]:
        return ping(check_peer, self.swarm, wait_timeout=self.wait_timeout) != float("inf")

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


def check_direct_reachability(swarm: Optional[Swarm] = None, peer_id: Optional[str] = None, *, max_peers: int = 5,
                              threshold: float = 0.5, wait_timeout: float = 5.0, **kwargs) -> Optional[bool]:
    """Ask up to ``max_peers`` other peers of a network swarm to dial ``peer_id`` back (``rpc_check``).

    True / False = at least / fewer than ``threshold`` of those that answered could reach us; None = nobody to ask
    (the first server of a swarm).  Swarms that live inside one box (in-process or rendezvous directory) have no NATs,
    firewalls or wrong ``--public_ip`` to detect: always True.  (reference reachability.py:55-84)"""
    if swarm is None or peer_id is None or not hasattr(swarm, "registry_address"):
        return True
    import random

    others = [p for p in swarm.peers() if p != peer_id]
    random.shuffle(others)
    verdicts = []
    for other in others:
        if len(verdicts) >= max_peers:
            break
        try:
            stub = swarm.connect(other, connect_timeout=wait_timeout, request_timeout=2 * wait_timeout + 1)
            verdicts.append(bool(stub.rpc_check(peer_id, wait_timeout)))
        except Exception as e:  # noqa: BLE001 - an unreachable *helper* says nothing about us
            logger.debug(f"reachability helper {other} did not answer: {e}")
    if not verdicts:
        return None
    return sum(verdicts) / len(verdicts) >= threshold


def validate_direct_reachability(swarm: Swarm, peer_id: str, **kwargs) -> None:
    """Raise with the fix spelled out when other peers cannot reach this server's announced address."""
    if check_direct_reachability(swarm, peer_id, **kwargs) is False:
        raise RuntimeError(
            f"Server {peer_id} announced an address that other peers of the swarm cannot connect to. Pass the address they should "
            f"use with --public_ip <ip> (or --announce_maddrs /ip4/<ip>/tcp/0), check firewalls between the boxes, or start the "
            f"server with --skip_reachability_check if you know better.")


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
        """Ask ``remote_peer`` whether it can reach ``check_peer`` (None when the helper itself does not answer)."""
        try:
            stub = self.swarm.connect(remote_peer, connect_timeout=self.wait_timeout, request_timeout=2 * self.wait_timeout + 1)
            if hasattr(stub, "rpc_check"):
                return bool(stub.rpc_check(check_peer, self.wait_timeout))
        except Exception as e:  # noqa: BLE001
            logger.debug(f"rpc_check via {remote_peer} failed: {e}")
            return None
        return ping(check_peer, self.swarm, wait_timeout=self.wait_timeout) != float("inf")

"""Client-side connection/routing options, mixed into every distributed model config
(reference: src/petals/client/config.py:13-35). Field names are kept so that
``from_pretrained(..., initial_peers=..., request_timeout=...)`` keeps working; the semantics of the
networking fields are re-mapped to the single-box swarm registry (``initial_peers`` = rendezvous
location(s) or an in-process swarm object)."""
from __future__ import annotations

import dataclasses
import os
from typing import Optional, Sequence, Union

from petals_b200.constants import PUBLIC_INITIAL_PEERS

_max_retries = os.getenv("PETALS_MAX_RETRIES")
DEFAULT_MAX_RETRIES = int(_max_retries) if isinstance(_max_retries, str) else None


@dataclasses.dataclass
class ClientConfig:
    initial_peers: Sequence[str] = tuple(PUBLIC_INITIAL_PEERS)  # rendezvous paths / swarm names
    dht_prefix: Optional[str] = None  # a prefix for all uids of this model
    daemon_startup_timeout: int = 60  # kept for CLI compatibility (no daemon is started)

    show_route: Union[str, bool] = "inference"  # log the chosen chain for "inference", or always (True)
    allowed_servers: Optional[Sequence[str]] = None  # whitelist of peer ids
    blocked_servers: Optional[Sequence[str]] = None  # blacklist of peer ids
    use_server_to_server: bool = True  # stage i pushes activations straight into stage i+1 (fused NVLink hop)

    connect_timeout: float = 5
    request_timeout: float = 3 * 60
    update_period: float = 60  # how often the block -> servers map is refreshed

    max_retries: Optional[int] = DEFAULT_MAX_RETRIES  # None = retry forever
    min_backoff: float = 1
    max_backoff: float = 60
    ban_timeout: float = 15

    active_adapter: Optional[str] = None  # LoRA adapter to activate server-side

    max_pinged: int = 3
    ping_timeout: float = 2

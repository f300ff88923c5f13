"""Membership records on top of the swarm registry (reference: src/petals/utils/dht.py:28-153).
Function names and semantics are the reference's; the ``dht`` argument is a :class:`Swarm`."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

from petals_b200.data_structures import ModuleUID, PeerID, RemoteModuleInfo, RemoteSpanInfo, ServerInfo, ServerState
from petals_b200.parallel.swarm import Swarm, get_dht_time
from petals_b200.utils.logging import get_logger

logger = get_logger(__name__)


def declare_active_modules(dht: Swarm, uids: Sequence[ModuleUID], server_info: ServerInfo, expiration_time: float,
                           peer_id: Optional[PeerID] = None, wait: bool = True) -> Dict[ModuleUID, bool]:
    """Publish that ``peer_id`` serves ``uids`` (one record per block uid, subkey = peer id)."""
    if isinstance(uids, str):
        uids = [uids]
    value = list(server_info.to_tuple())
    return {uid: dht.store(uid, peer_id, value, expiration_time) for uid in uids}


def get_remote_module_infos(dht: Swarm, uids: Sequence[ModuleUID], expiration_time: Optional[float] = None, *,
                            active_adapter: Optional[str] = None, latest: bool = False) -> List[RemoteModuleInfo]:
    """For every uid: the servers currently announcing it (optionally only those holding ``active_adapter``)."""
    now = get_dht_time() if expiration_time is None else expiration_time
    infos = []
    # a network registry answers for all uids in one round trip (parallel/registry.py); local ones are dict lookups
    records = dht.get_many(uids) if hasattr(dht, "get_many") else None
    for uid in uids:
        servers: Dict[PeerID, ServerInfo] = {}
        for peer_id, (value, exp) in (records.get(uid, {}) if records is not None else dht.get(uid)).items():
            if exp < now:
                continue
            try:
                info = ServerInfo.from_tuple(tuple(value))
            except Exception as e:  # noqa: BLE001 - a malformed record must not break routing
                logger.warning(f"malformed record for {uid} from {peer_id}: {e}")
                continue
            if active_adapter and active_adapter not in info.adapters:
                logger.debug(f"skipping {peer_id}: adapter {active_adapter} not in {info.adapters}")
                continue
            servers[peer_id] = info
        infos.append(RemoteModuleInfo(uid=uid, servers=servers))
    return infos


def compute_spans(module_infos: Sequence[RemoteModuleInfo], *, min_state: ServerState) -> Dict[PeerID, RemoteSpanInfo]:
    """peer -> the contiguous span it serves, as indices into ``module_infos`` (at ``min_state`` or better).

    Like the reference, a peer owns at most one span: if its records are not contiguous the last run wins."""
    spans: Dict[PeerID, RemoteSpanInfo] = {}
    for idx, info in enumerate(module_infos):
        for peer_id, server_info in sorted(info.servers.items()):
            if server_info.state.value < min_state.value:
                continue
            span = spans.get(peer_id)
            if span is not None and span.end == idx:
                span.end = idx + 1
            else:
                spans[peer_id] = RemoteSpanInfo(peer_id=peer_id, start=idx, end=idx + 1, server_info=server_info)
    return spans

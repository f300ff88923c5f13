"""The swarm registry: who serves which blocks (replaces the reference's Kademlia DHT).

Petals publishes ``ServerInfo`` records under block-uid keys in a hivemind DHT with TTLs
(src/petals/utils/dht.py:28-131). On one NVLink box membership is a small key-value store with the same data
model — key = block uid, subkey = peer id, value = ``ServerInfo.to_tuple()``, expiration time — in two forms:

* ``Swarm``      — in-process (default for single-process serving, benchmarks and most tests);
* ``FileSwarm``  — a rendezvous directory shared by several OS processes (``run_server`` + clients); peers
  are reached through Unix-socket RPC (parallel/transport.py);
* ``TcpSwarm``   — (parallel/registry.py) a registry process on the network + TCP endpoints, for swarms that span
  several boxes, addressed like the reference's bootstrap peers (``/ip4/<host>/tcp/<port>`` or ``tcp://host:port``).

``resolve_swarm(initial_peers)`` maps the reference's ``initial_peers`` argument onto these.
"""
from __future__ import annotations

import json
import os
import tempfile
import threading
import time
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

from petals_b200.utils.logging import get_logger

logger = get_logger(__name__)
INPROC_SCHEME = "inproc://"

_named: Dict[str, "Swarm"] = {}
_named_lock = threading.Lock()


def get_dht_time() -> float:
    return time.time()


class Swarm:
    """In-process registry + endpoint table."""

    def __init__(self, name: Optional[str] = None):
        self.name = name or f"swarm-{id(self):x}"
        self._lock = threading.RLock()
        self._records: Dict[str, Dict[str, Tuple[Any, float]]] = {}
        self._endpoints: Dict[str, Any] = {}
        self._meta: Dict[str, Dict[str, Tuple[Any, float]]] = {}
        with _named_lock:
            _named[self.name] = self

    # ---- addressing --------------------------------------------------------------------------------------
    @property
    def address(self) -> str:
        return INPROC_SCHEME + self.name

    @classmethod
    def named(cls, name: str) -> "Swarm":
        with _named_lock:
            if name in _named:
                return _named[name]
        return cls(name)

    # ---- key-value store (DHT data model) ---------------------------------------------------------------------
    def store(self, key: str, subkey: str, value: Any, expiration_time: float) -> bool:
        with self._lock:
            self._records.setdefault(key, {})[subkey] = (value, expiration_time)
        return True

    def get(self, key: str) -> Dict[str, Tuple[Any, float]]:
        now = get_dht_time()
        with self._lock:
            entries = self._records.get(key, {})
            for sub in [s for s, (_, exp) in entries.items() if exp < now]:
                del entries[sub]
            return dict(entries)

    def delete_subkey(self, subkey: str) -> None:
        with self._lock:
            for entries in self._records.values():
                entries.pop(subkey, None)

    # ---- endpoints ----------------------------------------------------------------------------------------------
    def register_endpoint(self, peer_id: str, handler: Any) -> None:
        with self._lock:
            self._endpoints[peer_id] = handler

    def unregister_endpoint(self, peer_id: str) -> None:
        with self._lock:
            self._endpoints.pop(peer_id, None)
        self.delete_subkey(peer_id)

    def connect(self, peer_id: str, **kwargs):
        with self._lock:
            if peer_id not in self._endpoints:
                raise ConnectionError(f"peer {peer_id} is not reachable in swarm {self.name}")
            return self._endpoints[peer_id]

    def peers(self) -> List[str]:
        with self._lock:
            return list(self._endpoints)


class FileSwarm(Swarm):
    """Rendezvous directory: ``records/<key>/<peer>.json`` + ``peers/<peer>.sock`` (Unix socket RPC)."""

    def __init__(self, path: str):
        self.path = os.path.abspath(path)
        os.makedirs(os.path.join(self.path, "records"), exist_ok=True)
        os.makedirs(os.path.join(self.path, "peers"), exist_ok=True)
        super().__init__(name="file:" + self.path)
        self._servers: Dict[str, Any] = {}

    @property
    def address(self) -> str:
        return self.path

    @staticmethod
    def _safe(key: str) -> str:
        return key.replace("/", "_")

    def store(self, key: str, subkey: str, value: Any, expiration_time: float) -> bool:
        d = os.path.join(self.path, "records", self._safe(key))
        os.makedirs(d, exist_ok=True)
        fd, tmp = tempfile.mkstemp(dir=d, suffix=".tmp")
        with os.fdopen(fd, "w") as f:
            json.dump({"value": value, "expiration": expiration_time}, f)
        os.replace(tmp, os.path.join(d, self._safe(subkey) + ".json"))  # atomic publish
        return True

    def get(self, key: str) -> Dict[str, Tuple[Any, float]]:
        d = os.path.join(self.path, "records", self._safe(key))
        out: Dict[str, Tuple[Any, float]] = {}
        if not os.path.isdir(d):
            return out
        now = get_dht_time()
        for fn in os.listdir(d):
            if not fn.endswith(".json"):
                continue
            try:
                with open(os.path.join(d, fn)) as f:
                    rec = json.load(f)
            except (OSError, ValueError):
                continue
            if rec["expiration"] >= now:
                out[fn[: -len(".json")]] = (rec["value"], rec["expiration"])
        return out

    def delete_subkey(self, subkey: str) -> None:
        root = os.path.join(self.path, "records")
        for key in os.listdir(root):
            p = os.path.join(root, key, self._safe(subkey) + ".json")
            if os.path.exists(p):
                try:
                    os.unlink(p)
                except OSError:
                    pass

    def register_endpoint(self, peer_id: str, handler: Any) -> None:
        from petals_b200.parallel.transport import RpcServer

        super().register_endpoint(peer_id, handler)
        server = RpcServer(handler, os.path.join(self.path, "peers", self._safe(peer_id) + ".sock"))
        server.start()
        self._servers[peer_id] = server

    def unregister_endpoint(self, peer_id: str) -> None:
        server = self._servers.pop(peer_id, None)
        if server is not None:
            server.shutdown()
        super().unregister_endpoint(peer_id)

    def connect(self, peer_id: str, connect_timeout: float = 5.0, request_timeout: float = 180.0, **kwargs):
        with self._lock:
            if peer_id in self._endpoints:  # same process: skip the socket
                return self._endpoints[peer_id]
        from petals_b200.parallel.transport import RemoteHandlerProxy

        sock = os.path.join(self.path, "peers", self._safe(peer_id) + ".sock")
        if not os.path.exists(sock):
            raise ConnectionError(f"peer {peer_id} has no endpoint in {self.path}")
        return RemoteHandlerProxy(sock, connect_timeout, request_timeout)

    def peers(self) -> List[str]:
        d = os.path.join(self.path, "peers")
        return [fn[: -len(".sock")] for fn in os.listdir(d) if fn.endswith(".sock")]


def resolve_swarm(initial_peers: Union[None, str, Swarm, Sequence[Union[str, Swarm]]]) -> Swarm:
    """Map the reference's ``initial_peers`` onto a registry object."""
    if isinstance(initial_peers, Swarm):
        return initial_peers
    if isinstance(initial_peers, str):
        initial_peers = [initial_peers]
    peers = list(initial_peers or [])
    if not peers:
        return Swarm.named("default")
    first = peers[0]
    if isinstance(first, Swarm):
        return first
    if first.startswith(INPROC_SCHEME):
        return Swarm.named(first[len(INPROC_SCHEME):])
    from petals_b200.parallel.transport import format_address, is_network_address, parse_address

    if is_network_address(first):  # tcp://host:port or /ip4/<host>/tcp/<port>[/p2p/..]: a registry on the network
        from petals_b200.parallel.registry import TcpSwarm

        _, host, port = parse_address(first)
        key = "tcp:" + format_address(host, port)
        with _named_lock:
            if key in _named:
                return _named[key]
        return TcpSwarm(first)
    key = "file:" + os.path.abspath(first)
    with _named_lock:
        if key in _named:
            return _named[key]
    return FileSwarm(first)

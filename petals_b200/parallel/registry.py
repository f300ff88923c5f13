"""A network swarm registry: the multi-host form of ``parallel/swarm.py``.

The reference bootstraps every swarm from a DHT node that ``petals.cli.run_dht`` starts and prints as a multiaddr
(src/petals/cli/run_dht.py:37-102); servers and clients pass it as ``--initial_peers`` and from then on talk TCP to
each other.  Inside one NVLink box none of that is needed (``Swarm`` / ``FileSwarm``), but a
deployment of several boxes needs the same two things the DHT gives the reference:

* a **membership store** with the DHT's data model — key = block uid, subkey = peer id, value, expiration — served by
  one small process (``RegistryServer``, started by ``python -m petals.cli.run_dht --host_maddrs /ip4/0.0.0.0/tcp/31337``);
* **peer endpoints reachable over TCP** — every stage worker serves its RPCs on ``tcp://host:port`` (same framing and
  the same C++ send/receive loops as the unix-socket transport) and announces that address through the registry.

``TcpSwarm`` is the client of both; it is what ``resolve_swarm`` returns for ``tcp://host:port`` or libp2p-style
``/ip4/<host>/tcp/<port>[/p2p/<id>]`` initial peers.  Hidden states between boxes travel host-staged over these
sockets exactly like the reference's (optionally in a wire codec, utils/compression.py); stages that share a box
keep their NVLink hops because the fabric is negotiated separately (parallel/fabric.py).

A single registry process is a deliberate simplification of Kademlia: record TTLs make a restarted registry
repopulate itself within one ``update_period`` (servers re-announce, exactly as they do for the DHT), and clients
keep their last routing table while it is away.
"""
from __future__ import annotations

import socket
import threading
import time
from typing import Any, Dict, List, Optional, Tuple

from petals_b200.parallel.swarm import Swarm, get_dht_time
from petals_b200.parallel.transport import (RemoteHandlerProxy, RpcServer, TrackedConn, format_address, open_connection, parse_address,
                                            recv_message, send_message)
from petals_b200.utils.logging import get_logger

logger = get_logger(__name__)

ENDPOINT_TTL = 3600.0  # an announced address lives this long without a refresh (servers refresh with every announce)


class _RegistryState:
    def __init__(self):
        self.lock = threading.Lock()
        self.records: Dict[str, Dict[str, Tuple[Any, float]]] = {}
        self.endpoints: Dict[str, Tuple[str, float]] = {}

    # every method returns a msgpack-able reply
    def store(self, key: str, subkey: str, value: Any, expiration: float) -> bool:
        with self.lock:
            self.records.setdefault(key, {})[subkey] = (value, expiration)  # last writer wins, like Swarm.store
        return True

    def get(self, key: str) -> Dict[str, list]:
        now = get_dht_time()
        with self.lock:
            entries = self.records.get(key, {})
            for sub in [s for s, (_, exp) in entries.items() if exp < now]:
                del entries[sub]
            if not entries:
                self.records.pop(key, None)
            return {s: [v, exp] for s, (v, exp) in entries.items()}

    def delete_subkey(self, subkey: str) -> None:
        with self.lock:
            for entries in self.records.values():
                entries.pop(subkey, None)

    def register(self, peer_id: str, address: str, ttl: float) -> None:
        with self.lock:
            self.endpoints[peer_id] = (address, get_dht_time() + ttl)

    def unregister(self, peer_id: str) -> None:
        with self.lock:
            self.endpoints.pop(peer_id, None)
        self.delete_subkey(peer_id)

    def lookup(self, peer_id: str) -> Optional[str]:
        with self.lock:
            rec = self.endpoints.get(peer_id)
            if rec is None or rec[1] < get_dht_time():
                self.endpoints.pop(peer_id, None)
                return None
            return rec[0]

    def peers(self) -> List[str]:
        now = get_dht_time()
        with self.lock:
            return sorted(p for p, (_, exp) in self.endpoints.items() if exp >= now)


class _RegistryConn(TrackedConn):
    """One connection = a sequence of {"method": "reg_*", "meta": {...}} requests (kept open by TcpSwarm)."""

    def handle(self) -> None:
        state: _RegistryState = self.server.rpc_handler  # type: ignore[attr-defined]
        sock: socket.socket = self.request
        peer_ip = self.client_address[0] if isinstance(self.client_address, tuple) else "127.0.0.1"
        while True:
            try:
                header, _ = recv_message(sock)
            except (ConnectionError, OSError):
                return
            m, a = header.get("method"), header.get("meta") or {}
            try:
                if m == "reg_store":
                    reply = state.store(a["key"], a["subkey"], a["value"], float(a["expiration"]))
                elif m == "reg_get":
                    reply = state.get(a["key"])
                elif m == "reg_get_many":
                    reply = {k: state.get(k) for k in a["keys"]}
                elif m == "reg_delete_subkey":
                    reply = state.delete_subkey(a["subkey"])
                elif m == "reg_register":
                    reply = state.register(a["peer_id"], a["address"], float(a.get("ttl", ENDPOINT_TTL)))
                elif m == "reg_unregister":
                    reply = state.unregister(a["peer_id"])
                elif m == "reg_lookup":
                    reply = state.lookup(a["peer_id"])
                elif m == "reg_peers":
                    reply = state.peers()
                elif m == "reg_whoami":  # the address this client is seen from: what it should announce by default
                    reply = peer_ip
                elif m == "reg_time":
                    reply = get_dht_time()
                else:
                    raise ValueError(f"unknown registry method {m!r}")
                send_message(sock, {"ok": True, "meta": reply})
            except Exception as e:  # noqa: BLE001 - report, keep the connection
                try:
                    send_message(sock, {"ok": False, "error": str(e), "etype": type(e).__name__})
                except OSError:
                    return


class RegistryServer:
    """``RegistryServer("tcp://0.0.0.0:31337").start()``; ``.address`` is what peers pass as ``--initial_peers``."""

    def __init__(self, address: str = "tcp://127.0.0.1:0"):
        self.state = _RegistryState()
        self._rpc = RpcServer(self.state, address, conn_class=_RegistryConn)
        self.address = self._rpc.address

    def start(self) -> "RegistryServer":
        self._rpc.start()
        return self

    def shutdown(self) -> None:
        self._rpc.shutdown()

    def peers(self) -> List[str]:
        return self.state.peers()


class TcpSwarm(Swarm):
    """Registry client + TCP endpoint table (one per process and registry address)."""

    def __init__(self, address: str, *, bind_host: str = "0.0.0.0", announce_host: Optional[str] = None,
                 connect_timeout: float = 5.0, request_timeout: float = 30.0):
        kind, host, port = parse_address(address)
        assert kind == "tcp", address
        self.registry_address = format_address(host, port)
        self.bind_host, self.announce_host = bind_host, announce_host
        self._connect_timeout, self._request_timeout = connect_timeout, request_timeout
        self._sock: Optional[socket.socket] = None
        self._io_lock = threading.Lock()
        self._servers: Dict[str, RpcServer] = {}
        self._addr_cache: Dict[str, Tuple[str, float]] = {}
        super().__init__(name="tcp:" + self.registry_address)

    @property
    def address(self) -> str:
        return self.registry_address

    # ---- registry RPC (one persistent connection, re-opened on failure) -------------------------------------------------
    def _call(self, method: str, **meta) -> Any:
        last: Optional[Exception] = None
        for attempt in range(2):
            with self._io_lock:
                try:
                    if self._sock is None:
                        self._sock = open_connection(self.registry_address, self._connect_timeout, self._request_timeout)
                    send_message(self._sock, {"method": method, "meta": meta})
                    reply, _ = recv_message(self._sock)
                except (OSError, ConnectionError) as e:
                    last = e
                    if self._sock is not None:
                        try:
                            self._sock.close()
                        finally:
                            self._sock = None
                    continue
            if not reply.get("ok"):
                raise RuntimeError(f"registry {self.registry_address}: {reply.get('error')}")
            return reply.get("meta")
        raise ConnectionError(f"swarm registry {self.registry_address} is unreachable: {last}")

    # ---- DHT data model -------------------------------------------------------------------------------------------------------
    def store(self, key: str, subkey: str, value: Any, expiration_time: float) -> bool:
        return bool(self._call("reg_store", key=key, subkey=subkey, value=value, expiration=expiration_time))

    def get(self, key: str) -> Dict[str, Tuple[Any, float]]:
        return {s: (v, exp) for s, (v, exp) in (self._call("reg_get", key=key) or {}).items()}

    def get_many(self, keys) -> Dict[str, Dict[str, Tuple[Any, float]]]:
        """One round trip for a whole model's block uids (what a routing refresh asks for)."""
        reply = self._call("reg_get_many", keys=list(keys)) or {}
        return {k: {s: (v, exp) for s, (v, exp) in entries.items()} for k, entries in reply.items()}

    def delete_subkey(self, subkey: str) -> None:
        self._call("reg_delete_subkey", subkey=subkey)

    # ---- endpoints --------------------------------------------------------------------------------------------------------------
    def _announce_host(self) -> str:
        if self.announce_host:
            return self.announce_host
        if self.bind_host not in ("0.0.0.0", "::", ""):
            return self.bind_host
        return str(self._call("reg_whoami"))

    def register_endpoint(self, peer_id: str, handler: Any) -> None:
        with self._lock:
            self._endpoints[peer_id] = handler
        server = RpcServer(handler, format_address(self.bind_host, 0))
        server.start()
        self._servers[peer_id] = server
        address = format_address(self._announce_host(), server.port)
        self._call("reg_register", peer_id=peer_id, address=address, ttl=ENDPOINT_TTL)
        logger.info(f"peer {peer_id} serves RPCs on {address}")

    def refresh_endpoint(self, peer_id: str) -> None:
        server = self._servers.get(peer_id)
        if server is not None:
            self._call("reg_register", peer_id=peer_id, address=format_address(self._announce_host(), server.port), ttl=ENDPOINT_TTL)

    def unregister_endpoint(self, peer_id: str) -> None:
        server = self._servers.pop(peer_id, None)
        if server is not None:
            server.shutdown()
        with self._lock:
            self._endpoints.pop(peer_id, None)
        try:
            self._call("reg_unregister", peer_id=peer_id)
        except ConnectionError:
            pass  # the registry is gone: its records die with it

    def connect(self, peer_id: str, connect_timeout: float = 5.0, request_timeout: float = 180.0, **kwargs):
        with self._lock:
            if peer_id in self._endpoints:  # same process: skip the socket
                return self._endpoints[peer_id]
        cached = self._addr_cache.get(peer_id)
        if cached is None or cached[1] < time.monotonic():
            address = self._call("reg_lookup", peer_id=peer_id)
            if address is None:
                raise ConnectionError(f"peer {peer_id} has no endpoint in swarm {self.registry_address}")
            self._addr_cache[peer_id] = cached = (address, time.monotonic() + 30.0)
        return RemoteHandlerProxy(cached[0], connect_timeout, request_timeout)

    def forget(self, peer_id: str) -> None:
        self._addr_cache.pop(peer_id, None)

    def peers(self) -> List[str]:
        return list(self._call("reg_peers") or [])

    def close(self) -> None:
        for peer_id in list(self._servers):
            self.unregister_endpoint(peer_id)
        with self._io_lock:
            if self._sock is not None:
                self._sock.close()
                self._sock = None

"""RemoteSequenceManager: keeps the block -> stages map fresh and turns a block range into a chain of spans
(reference: src/petals/client/routing/sequence_manager.py:59-519).

Kept: the public surface (``make_sequence(start, end, mode=...)``, slicing, ``rpc_info``,
``on_request_failure/success``, ``get_retry_delay``, ``get_request_metadata``, ``MissingBlocksError``), the two
routing modes — ``min_latency`` = shortest path over (peer, block) nodes with compute cost ``blocks / inference_rps``
plus per-hop cost and a 10 s penalty for servers short on KV cache, ``max_throughput`` = random choice weighted by
span length — allow/block lists and the temporary ban list with exponential back-off.

Changed: membership comes from the in-box swarm registry (no DHT process); the per-hop cost is the measured
control-channel RTT (micro-seconds) instead of the reference's 18 ms serialisation constant; the shortest path
is a small in-file Dijkstra (no Dijkstar dependency); the first pinged-servers sample is not thrown away (Q3).
"""
from __future__ import annotations

import dataclasses
import heapq
import itertools
import logging
import random
import threading
import time
import weakref
from typing import Any, Dict, List, Optional, Sequence, Set, Tuple, Union

import numpy as np

from petals_b200.client.config import ClientConfig
from petals_b200.client.routing.sequence_info import RemoteSequenceInfo
from petals_b200.client.routing.spending_policy import NoSpendingPolicy
from petals_b200.data_structures import ModuleUID, RemoteSpanInfo, ServerState
from petals_b200.parallel.swarm import Swarm, resolve_swarm
from petals_b200.utils.dht import get_remote_module_infos
from petals_b200.utils.logging import get_logger
from petals_b200.utils.ping import PingAggregator
from petals_b200.utils.random import sample_up_to

logger = get_logger(__name__)


class MissingBlocksError(RuntimeError):
    def __init__(self, block_indices: Union[int, Sequence[int]]):
        super().__init__(f"No servers holding blocks {block_indices} are online. "
                         "Start a stage that serves them (python -m petals.cli.run_server ... --block_indices a:b) "
                         "or check the rendezvous location passed as initial_peers.")


def maybe_log_traceback(exc: Exception) -> None:
    logger.log(logging.DEBUG if str(exc) or isinstance(exc, TimeoutError) else logging.WARNING, "See detailed traceback below:", exc_info=True)


class _Blacklist:
    """Peers banned until a deadline; the ban duration doubles with consecutive failures."""

    def __init__(self, base_time: float, backoff_rate: float = 2.0):
        self.base_time, self.backoff_rate = base_time, backoff_rate
        self._until: Dict[str, float] = {}
        self._fails: Dict[str, int] = {}

    def register_failure(self, peer: str) -> None:
        n = self._fails.get(peer, 0)
        self._fails[peer] = n + 1
        self._until[peer] = time.monotonic() + self.base_time * self.backoff_rate ** n

    def register_success(self, peer: str) -> None:
        self._fails.pop(peer, None)
        self._until.pop(peer, None)

    def __contains__(self, peer: str) -> bool:
        return self._until.get(peer, 0) > time.monotonic()


@dataclasses.dataclass
class SequenceManagerState:
    p2p: Any = None
    sequence_info: Optional[RemoteSequenceInfo] = None
    rpc_info: Optional[dict] = None
    banned_peers: Optional[_Blacklist] = None
    blocked_servers: Optional[Set[str]] = None
    allowed_servers: Optional[Set[str]] = None
    ping_aggregator: Optional[PingAggregator] = None


class RemoteSequenceManager:
    def __init__(self, config: ClientConfig, block_uids: Sequence[ModuleUID], *, dht: Optional[Swarm] = None,
                 state: Optional[SequenceManagerState] = None):
        assert config.dht_prefix, "Could not find dht_prefix in config, please create the model with dht_prefix=..."
        assert len(block_uids) > 0, "Sequences must contain at least one block"
        self.config = config
        if state is None:
            state = SequenceManagerState()
        self.state = state
        self.dht: Swarm = dht if dht is not None else resolve_swarm(config.initial_peers)
        self.lock_changes = threading.Lock()
        self.policy = NoSpendingPolicy()
        if state.banned_peers is None:
            state.banned_peers = _Blacklist(base_time=config.ban_timeout, backoff_rate=2.0)
        if state.sequence_info is None:
            state.sequence_info = RemoteSequenceInfo.make_empty(block_uids)
        if state.allowed_servers is None and config.allowed_servers is not None:
            state.allowed_servers = set(config.allowed_servers)
        if state.blocked_servers is None and config.blocked_servers is not None:
            state.blocked_servers = set(config.blocked_servers)
        if state.ping_aggregator is None:
            state.ping_aggregator = PingAggregator(self.dht)
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self._need_update = threading.Event()
        self.ready = threading.Event()
        if state.sequence_info.last_updated_time is not None:
            assert block_uids == state.sequence_info.block_uids
            self.ready.set()

    # ---- background refresh (reference :493-519) ---------------------------------------------------------------
    def _ensure_thread(self) -> None:
        if self._thread is None or not self._thread.is_alive():
            # the thread only holds a weak reference: a manager (and the model that owns it) that nobody uses any more is
            # collected and its refresh loop ends, instead of polling the registry for the rest of the process's life
            self._thread = threading.Thread(target=_update_loop, args=(weakref.ref(self), self._stop, self._need_update),
                                            name="sequence-manager", daemon=True)
            self._thread.start()

    def __del__(self):
        stop = getattr(self, "_stop", None)
        if stop is not None:
            stop.set()
            self._need_update.set()

    def update(self, *, wait: bool = True) -> None:
        """Refresh the block -> servers map now."""
        self._update()
        self._ensure_thread()

    def _update(self, background: bool = False) -> None:
        for attempt_no in itertools.count():
            try:
                self._update_once()
                if any(not info.servers for info in self.state.sequence_info.block_infos):
                    missing = [i for i, info in enumerate(self.state.sequence_info.block_infos) if not info.servers]
                    raise MissingBlocksError(missing)
                self.ready.set()
                return
            except MissingBlocksError as e:
                if background or (self.config.max_retries is not None and attempt_no >= self.config.max_retries):
                    raise  # the refresh loop retries on its own schedule; callers that wait for a route retry here
                delay = self.get_retry_delay(attempt_no)
                logger.warning(f"Could not find route through the model: {e!r} (retry in {delay:.0f} sec)")
                if self._stop.wait(delay):
                    raise

    def _update_once(self) -> None:
        new_infos = get_remote_module_infos(self.dht, self.state.sequence_info.block_uids, active_adapter=self.config.active_adapter, latest=True)
        for info in new_infos:
            servers = {p: s for p, s in info.servers.items() if s.state == ServerState.ONLINE}
            if self.state.allowed_servers is not None:
                servers = {p: s for p, s in servers.items() if p in self.state.allowed_servers}
            if self.state.blocked_servers is not None:
                servers = {p: s for p, s in servers.items() if p not in self.state.blocked_servers}
            unbanned = {p: s for p, s in servers.items() if p not in self.state.banned_peers}
            # if every candidate is banned, ignore the ban list rather than give up (reference :362-370)
            info.servers = unbanned if unbanned or not servers else servers
        with self.lock_changes:
            self.state.sequence_info.update_(new_infos)
        first = self.state.sequence_info.spans_containing_block[0] if len(self.state.sequence_info) else []
        candidates = [s.peer_id for s in first]
        if candidates:
            self.state.ping_aggregator.ping(sample_up_to(candidates, self.config.max_pinged), wait_timeout=self.config.ping_timeout)

    # ---- slicing -----------------------------------------------------------------------------------------------------
    def __getitem__(self, ix: Union[int, slice]) -> "RemoteSequenceManager":
        assert isinstance(ix, (int, slice))
        if not isinstance(ix, slice):
            ix = slice(int(ix), int(ix) + 1, 1)
        return type(self)(self.config, self.block_uids[ix], dht=self.dht, state=self.state_slice(ix))

    def state_slice(self, ix: slice) -> SequenceManagerState:
        # shares bans, lists and pings; rpc_info is the same for all blocks of one model
        return dataclasses.replace(self.state, sequence_info=self.state.sequence_info[ix])

    def __len__(self) -> int:
        return len(self.block_uids)

    @property
    def block_uids(self) -> Tuple[ModuleUID, ...]:
        return self.state.sequence_info.block_uids

    @property
    def is_alive(self) -> bool:
        return self._thread is not None and self._thread.is_alive()

    # ---- routing -------------------------------------------------------------------------------------------------------
    def make_sequence(self, start_index: int = 0, end_index: Optional[int] = None, *, mode: str, cache_tokens_needed: Optional[int] = None) -> List[RemoteSpanInfo]:
        """A chain of spans that covers blocks [start_index, end_index) exactly once each."""
        if not self.ready.is_set() or self.state.sequence_info.last_updated_time is None:
            self.update(wait=True)
        end_index = end_index if end_index is not None else len(self)
        if mode == "min_latency":
            span_sequence = self._make_sequence_with_min_latency(start_index, end_index, cache_tokens_needed=cache_tokens_needed)
        elif mode == "max_throughput":
            span_sequence = self._make_sequence_with_max_throughput(start_index, end_index)
        else:
            raise RuntimeError(f"Unexpected mode {mode}")
        if self.config.show_route is True or (mode == "min_latency" and self.config.show_route == "inference"):
            route = " => ".join(f"{s.start}:{s.end} via {s.peer_id}" for s in span_sequence)
            logger.debug(f"Route found: {route}")
        return span_sequence

    def _make_sequence_with_min_latency(self, start_index: int, end_index: int, *, cache_tokens_needed: Optional[int]) -> List[RemoteSpanInfo]:
        if start_index == end_index:
            return []
        with self.lock_changes:
            missing = [i for i in range(start_index, end_index) if not self.state.sequence_info.spans_containing_block[i]]
            if missing:
                raise MissingBlocksError(missing)
            graph = self._build_inference_graph(start_index, end_index, cache_tokens_needed=cache_tokens_needed)
        path = _dijkstra(graph, "start", "end")
        if path is None:
            raise MissingBlocksError(list(range(start_index, end_index)))
        # nodes look like (peer, block); consecutive nodes with the same peer form one span
        span_sequence: List[RemoteSpanInfo] = []
        nodes = [n for n in path if isinstance(n, tuple)]
        for (peer, blk), (next_peer, next_blk) in zip(nodes[:-1], nodes[1:]):
            if peer == next_peer and next_blk > blk:
                info = self.state.sequence_info.block_infos[blk].servers[peer]
                if span_sequence and span_sequence[-1].peer_id == peer and span_sequence[-1].end == blk:
                    span_sequence[-1].end = next_blk
                else:
                    span_sequence.append(RemoteSpanInfo(peer_id=peer, start=blk, end=next_blk, server_info=info))
        return span_sequence

    def _build_inference_graph(self, start_index: int, end_index: int, *, cache_tokens_needed: Optional[int],
                               overhead_delay: float = 20e-6, default_inference_rps: float = 300, alloc_delay: float = 10) -> Dict[Any, Dict[Any, float]]:
        """Edges: start -> (peer, b) entering, (peer, b) -> (peer, b') compute, (peer, b') -> (peer2, b') hop, -> end."""
        missing = [b for b in range(start_index, end_index) if not self.state.sequence_info.spans_containing_block[b]]
        if missing:
            raise MissingBlocksError(missing)
        client_pings = self.state.ping_aggregator.to_dict()
        graph: Dict[Any, Dict[Any, float]] = {"start": {}, "end": {}}

        def add(u, v, w):
            graph.setdefault(u, {})[v] = min(w, graph.get(u, {}).get(v, float("inf")))
            graph.setdefault(v, {})

        spans = [s for s in self.state.sequence_info.spans_by_priority if s.end > start_index and s.start < end_index]
        for span in spans:
            lo, hi = max(span.start, start_index), min(span.end, end_index)
            rps = span.server_info.inference_rps or default_inference_rps
            penalty = 0.0
            if cache_tokens_needed is not None and span.server_info.cache_tokens_left is not None:
                # tokens x blocks needed if the whole usable part of the span is taken
                if span.server_info.cache_tokens_left < cache_tokens_needed * 2 * (hi - lo):
                    penalty = alloc_delay
            rtt = client_pings.get(span.peer_id, 0.0)
            for b in range(lo, hi):
                # entering this span at block b (from the client at the very start, or after a hop)
                if b == start_index:
                    add("start", (span.peer_id, b), rtt / 2 + overhead_delay + penalty)
                for b2 in range(b + 1, hi + 1):
                    add((span.peer_id, b), (span.peer_id, b2), (b2 - b) / rps)
            add((span.peer_id, end_index), "end", rtt / 2) if hi == end_index else None
        # hops between spans at every block boundary
        for b in range(start_index + 1, end_index):
            here = [s for s in spans if s.start < b <= s.end]  # can finish at b
            there = [s for s in spans if s.start <= b < s.end]  # can start at b
            for a in here:
                for c in there:
                    if a.peer_id == c.peer_id:
                        continue
                    next_pings = a.server_info.next_pings or {}
                    delay = next_pings.get(c.peer_id, client_pings.get(c.peer_id, 0.0)) / 2 + overhead_delay
                    pen = 0.0
                    if cache_tokens_needed is not None and c.server_info.cache_tokens_left is not None:
                        if c.server_info.cache_tokens_left < cache_tokens_needed * 2 * (min(c.end, end_index) - b):
                            pen = alloc_delay
                    add((a.peer_id, b), (c.peer_id, b), delay + pen)
        return graph

    def _make_sequence_with_max_throughput(self, start_index: int, end_index: int) -> List[RemoteSpanInfo]:
        span_sequence: List[RemoteSpanInfo] = []
        current = start_index
        while current < end_index:
            candidates = self.state.sequence_info.spans_containing_block[current]
            if not candidates:
                raise MissingBlocksError(current)
            weights = np.array([s.end - current for s in candidates], dtype=np.float64)
            chosen = candidates[int(np.random.choice(len(candidates), p=weights / weights.sum()))]
            assert chosen.start <= current < chosen.end
            span_sequence.append(dataclasses.replace(chosen, start=current, end=min(chosen.end, end_index)))
            current = span_sequence[-1].end
        return span_sequence

    # ---- failure handling --------------------------------------------------------------------------------------------
    def on_request_failure(self, peer_id: Optional[str]) -> None:
        """Ban the peer for a while and drop it from the current map; the next refresh may bring it back."""
        if peer_id is not None:
            logger.debug(f"Peer {peer_id} did not respond, banning it temporarily")
            self.state.banned_peers.register_failure(peer_id)
            if hasattr(self.dht, "forget"):  # network swarms cache peer addresses: a restarted server announces a new port
                self.dht.forget(peer_id)
        with self.lock_changes:
            should_update = False
            for info in self.state.sequence_info.block_infos:
                if peer_id in info.servers and len(info.servers) > 1:
                    info.servers.pop(peer_id)  # keep the last candidate: better a retry than MissingBlocksError
                    should_update = True
            if should_update:
                self.state.sequence_info.update_(self.state.sequence_info.block_infos)
        self._need_update.set()

    def on_request_success(self, peer_id: str) -> None:
        self.state.banned_peers.register_success(peer_id)

    def get_retry_delay(self, attempt_no: int) -> float:
        if attempt_no == 0:
            return 0.0
        return min(self.config.min_backoff * 2 ** (attempt_no - 1), self.config.max_backoff)

    # ---- per-request data ----------------------------------------------------------------------------------------------
    @property
    def rpc_info(self) -> dict:
        """Schema / version info of the served model, fetched lazily from any stage holding the first block."""
        if self.state.rpc_info is not None:
            return self.state.rpc_info
        if not self.ready.is_set():
            self.update(wait=True)
        for attempt_no in itertools.count():
            peer_id = None
            try:
                candidates = [s.peer_id for s in self.state.sequence_info.spans_containing_block[0]]
                if not candidates:
                    raise MissingBlocksError(0)
                peer_id = random.choice(candidates)
                self.state.rpc_info = self.dht.connect(peer_id, connect_timeout=self.config.connect_timeout).rpc_info(self.block_uids[0])
                self.on_request_success(peer_id)
                return self.state.rpc_info
            except Exception as e:  # noqa: BLE001
                self.on_request_failure(peer_id)
                if self.config.max_retries is not None and attempt_no + 1 >= self.config.max_retries:
                    raise
                delay = self.get_retry_delay(attempt_no)
                logger.warning(f"Caught exception when gathering information from peer {peer_id} (retry in {delay:.0f} sec): {e!r}")
                maybe_log_traceback(e)
                time.sleep(delay)

    def get_request_metadata(self, protocol: str, args_structure: Any = None, *args, **kwargs) -> Dict[str, Any]:
        meta = dict(points=self.policy.get_points(protocol, *args, **kwargs), active_adapter=self.config.active_adapter,
                    args_structure=args_structure)
        codec = getattr(self.config, "output_compression", None)
        if codec and protocol in ("rpc_inference", "rpc_forward"):  # one output tensor each; backward keeps the server's default per gradient
            meta["output_compression"] = [codec]
        return meta

    def connect(self, peer_id: str):
        stub = self.dht.connect(peer_id, connect_timeout=self.config.connect_timeout, request_timeout=self.config.request_timeout)
        codec = getattr(self.config, "wire_compression", None)
        if codec and hasattr(stub, "socket_path"):  # a socket proxy: compress what this client sends through it
            stub.compression = codec
        return stub

    def shutdown(self) -> None:
        self._stop.set()
        self._need_update.set()
        if self._thread is not None and self._thread.is_alive():
            self._thread.join(timeout=2)


def _update_loop(manager_ref, stop: threading.Event, need_update: threading.Event) -> None:
    """Background refresh (reference sequence_manager.py:493-519). Holds the manager only while refreshing."""
    failures = 0
    while not stop.is_set():
        manager = manager_ref()
        if manager is None:
            return
        try:
            manager._update(background=True)
            failures, delay = 0, manager.config.update_period
        except Exception as e:  # noqa: BLE001 - keep refreshing
            delay = min(manager.config.update_period, manager.get_retry_delay(failures))
            failures += 1
            logger.debug(f"sequence info update failed: {e!r} (next attempt in {delay:.0f} sec)")
        del manager
        need_update.wait(delay)
        need_update.clear()


def _dijkstra(graph: Dict[Any, Dict[Any, float]], src: Any, dst: Any) -> Optional[List[Any]]:
    dist = {src: 0.0}
    prev: Dict[Any, Any] = {}
    counter = itertools.count()
    heap = [(0.0, next(counter), src)]
    done = set()
    while heap:
        d, _, u = heapq.heappop(heap)
        if u in done:
            continue
        done.add(u)
        if u == dst:
            break
        for v, w in graph.get(u, {}).items():
            nd = d + w
            if nd < dist.get(v, float("inf")):
                dist[v], prev[v] = nd, u
                heapq.heappush(heap, (nd, next(counter), v))
    if dst not in dist:
        return None
    path = [dst]
    while path[-1] != src:
        path.append(prev[path[-1]])
    return path[::-1]

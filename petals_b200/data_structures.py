"""Shared value types of the control plane (reference: src/petals/data_structures.py:1-117).

Two things are contractual and therefore kept: the UID grammar (``"<dht_prefix>.<block_idx>"``, chains joined by one space —
the only "wire" convention user scripts ever see) and the *field names* of ``ServerInfo`` (they travel in registry records and
``rpc_info``).  Everything else is this repo's own: plain dataclasses validated in ``__post_init__`` (no pydantic on the
announce path), records serialised as ``(state, throughput, {non-default extras})``, and a peer is a GPU worker named by any
string (``"gpu3"``, ``"stage0"``) instead of a libp2p identity.
"""
from __future__ import annotations

import dataclasses
import math
from enum import Enum
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

ModuleUID = str
PeerID = str
UID_DELIMITER = "."  # <prefix>.<block index>
CHAIN_DELIMITER = " "  # several block uids in one request


# ---- uid grammar ---------------------------------------------------------------------------------------------------------
def make_uid(prefix: str, index: int) -> ModuleUID:
    return prefix + UID_DELIMITER + str(int(index))


def parse_uid(uid: ModuleUID) -> Tuple[str, int]:
    prefix, sep, index = uid.rpartition(UID_DELIMITER)
    if not sep or CHAIN_DELIMITER in uid or not index.isdigit():
        raise ValueError(f"not a single-block uid: {uid!r}")
    return prefix, int(index)


def join_uids(uids: Iterable[ModuleUID]) -> str:
    return CHAIN_DELIMITER.join(uids)


def split_uids(chain: str) -> List[ModuleUID]:
    return [uid for uid in chain.split(CHAIN_DELIMITER) if uid]  # only the delimiter separates: prefixes may contain other whitespace


# ---- validation helpers (what pydantic's conint / confloat enforce in the reference) ---------------------------------------
def _require_count(name: str, value: Any, minimum: int = 0, optional: bool = True) -> None:
    if value is None and optional:
        return
    if isinstance(value, bool) or not isinstance(value, int) or value < minimum:
        raise ValueError(f"{name} must be an integer >= {minimum}, got {value!r}")


def _require_rate(name: str, value: Any, optional: bool = True) -> None:
    if value is None and optional:
        return
    if isinstance(value, bool) or not isinstance(value, (int, float)) or not math.isfinite(value) or value < 0:
        raise ValueError(f"{name} must be a finite number >= 0, got {value!r}")


class ServerState(Enum):
    OFFLINE = 0
    JOINING = 1
    ONLINE = 2


@dataclasses.dataclass
class ModelInfo:
    """Registry entry of a model (key ``_petals.models``): how many blocks a complete chain has."""

    num_blocks: int
    repository: Optional[str] = None

    def __post_init__(self):
        _require_count("num_blocks", self.num_blocks, minimum=1, optional=False)

    def to_dict(self) -> dict:
        return {"num_blocks": self.num_blocks, "repository": self.repository}

    @classmethod
    def from_dict(cls, source: dict) -> "ModelInfo":
        return cls(num_blocks=source["num_blocks"], repository=source.get("repository"))


@dataclasses.dataclass
class ServerInfo:
    """What a stage advertises about itself under every block uid it serves."""

    state: ServerState
    throughput: float

    start_block: Optional[int] = None
    end_block: Optional[int] = None

    public_name: Optional[str] = None
    version: Optional[str] = None

    network_rps: Optional[float] = None
    forward_rps: Optional[float] = None
    inference_rps: Optional[float] = None

    adapters: Sequence[str] = ()
    torch_dtype: Optional[str] = None
    quant_type: Optional[str] = None
    using_relay: Optional[bool] = None
    cache_tokens_left: Optional[int] = None
    next_pings: Optional[Dict[str, float]] = None

    _RATES = ("network_rps", "forward_rps", "inference_rps")
    _COUNTS = ("start_block", "end_block", "cache_tokens_left")

    def __post_init__(self):
        if not isinstance(self.state, ServerState):
            self.state = ServerState(self.state)
        _require_rate("throughput", self.throughput, optional=False)
        for name in self._RATES:
            _require_rate(name, getattr(self, name))
        for name in self._COUNTS:
            _require_count(name, getattr(self, name))
        self.adapters = tuple(self.adapters)
        if self.next_pings is not None:
            for peer, rtt in self.next_pings.items():
                _require_rate(f"next_pings[{peer!r}]", rtt, optional=False)

    # wire form: (state value, throughput, extras); extras only carry what differs from the defaults
    def to_tuple(self) -> Tuple[int, float, dict]:
        extras = {}
        for f in dataclasses.fields(self):
            if f.name in ("state", "throughput"):
                continue
            value = getattr(self, f.name)
            if value is None or (f.name == "adapters" and not value):
                continue
            extras[f.name] = list(value) if f.name == "adapters" else value
        return self.state.value, self.throughput, extras

    @classmethod
    def from_tuple(cls, source: Sequence[Any]) -> "ServerInfo":
        if not isinstance(source, (tuple, list)) or len(source) < 2:
            raise TypeError(f"expected (state, throughput[, extras]), got {type(source).__name__}")
        extras = dict(source[2]) if len(source) > 2 and source[2] else {}
        accepted = {f.name for f in dataclasses.fields(cls)} - {"state", "throughput"}
        return cls(ServerState(source[0]), float(source[1]), **{k: v for k, v in extras.items() if k in accepted})  # unknown keys: newer peers


@dataclasses.dataclass
class RemoteModuleInfo:
    """A block uid and every peer that currently serves it."""

    uid: ModuleUID
    servers: Dict[PeerID, ServerInfo]


@dataclasses.dataclass
class RemoteSpanInfo:
    """A contiguous span of blocks ``[start, end)`` held by one peer."""

    peer_id: PeerID
    start: int
    end: int
    server_info: ServerInfo

    @property
    def length(self) -> int:
        return self.end - self.start

    @property
    def state(self) -> ServerState:
        return self.server_info.state

    @property
    def throughput(self) -> float:
        return self.server_info.throughput


RPCInfo = Dict[str, Any]
Handle = int


@dataclasses.dataclass(frozen=True)
class InferenceMetadata:
    """Per-block bookkeeping of one inference step (kept for API parity; sessions here carry page tables instead of handles)."""

    uid: ModuleUID
    prefix_length: int
    cache_handles: Tuple[Handle, ...]
    active_adapter: Optional[str]

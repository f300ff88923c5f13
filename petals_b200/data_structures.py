"""Shared value types of the control plane (reference: src/petals/data_structures.py:1-117).

UID grammar is kept (``"<dht_prefix>.<block_idx>"``, chains joined by a space) because it is the only
"wire" convention that user scripts ever see; everything else is re-thought for a single NVLink box:
a peer is a GPU worker (``PeerID`` = ``"gpu<rank>"`` or any string), not a libp2p identity.
"""
from __future__ import annotations

import dataclasses
from enum import Enum
from typing import Any, Dict, Optional, Sequence, Tuple

import pydantic

ModuleUID = str
PeerID = str
UID_DELIMITER = "."  # <prefix>.<block index>
CHAIN_DELIMITER = " "  # several block uids in one request


def make_uid(prefix: str, index: int) -> ModuleUID:
    return f"{prefix}{UID_DELIMITER}{index}"


def parse_uid(uid: ModuleUID) -> Tuple[str, int]:
    if CHAIN_DELIMITER in uid or UID_DELIMITER not in uid:
        raise ValueError(f"not a single-block uid: {uid!r}")
    prefix, index = uid.rsplit(UID_DELIMITER, 1)
    return prefix, int(index)


def join_uids(uids: Sequence[ModuleUID]) -> str:
    return CHAIN_DELIMITER.join(uids)


def split_uids(chain: str) -> list:
    return [u for u in chain.split(CHAIN_DELIMITER) if u]


@pydantic.dataclasses.dataclass
class ModelInfo:
    num_blocks: pydantic.conint(ge=1, strict=True)
    repository: Optional[str] = None

    def to_dict(self) -> dict:
        return dataclasses.asdict(self)

    @classmethod
    def from_dict(cls, source: dict) -> "ModelInfo":
        return cls(**source)


class ServerState(Enum):
    OFFLINE = 0
    JOINING = 1
    ONLINE = 2


RPS = pydantic.confloat(ge=0, allow_inf_nan=False, strict=True)


@pydantic.dataclasses.dataclass
class ServerInfo:
    """What a stage advertises about itself (published through the swarm registry instead of a DHT)."""

    state: ServerState
    throughput: RPS

    start_block: Optional[pydantic.conint(ge=0, strict=True)] = None
    end_block: Optional[pydantic.conint(ge=0, strict=True)] = None

    public_name: Optional[str] = None
    version: Optional[str] = None

    network_rps: Optional[RPS] = None
    forward_rps: Optional[RPS] = None
    inference_rps: Optional[RPS] = None

    adapters: Sequence[str] = ()
    torch_dtype: Optional[str] = None
    quant_type: Optional[str] = None
    using_relay: Optional[bool] = None
    cache_tokens_left: Optional[pydantic.conint(ge=0, strict=True)] = None
    next_pings: Optional[Dict[str, pydantic.confloat(ge=0, strict=True)]] = None

    def to_tuple(self) -> Tuple[int, float, dict]:
        extra = dataclasses.asdict(self)
        del extra["state"], extra["throughput"]
        return (self.state.value, self.throughput, extra)

    @classmethod
    def from_tuple(cls, source: tuple) -> "ServerInfo":
        if not isinstance(source, (tuple, list)):
            raise TypeError(f"expected a tuple, got {type(source)}")
        state, throughput = source[:2]
        extra = dict(source[2]) if len(source) > 2 else {}
        known = {f.name for f in dataclasses.fields(cls)}
        extra = {k: v for k, v in extra.items() if k in known}  # forward compatibility
        if "adapters" in extra:
            extra["adapters"] = tuple(extra["adapters"])
        return cls(state=ServerState(state), throughput=float(throughput), **extra)


@dataclasses.dataclass
class RemoteModuleInfo:
    """A block uid and every peer that currently serves it."""

    uid: ModuleUID
    servers: Dict[PeerID, ServerInfo]


@dataclasses.dataclass
class RemoteSpanInfo:
    """A contiguous span of blocks held by one peer."""

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
    uid: ModuleUID
    prefix_length: int
    cache_handles: Tuple[Handle, ...]
    active_adapter: Optional[str]

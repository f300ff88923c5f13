"""Env / API driven fault injector (the reference has no fault injection; its fail-over paths are only exercised
by real outages).

A *plan* is a ``;``-separated list of rules ``rpc=<name>[,peer=<peer_id>][,after=<n>][,times=<k>][,error=<text>]``:
the ``after+1``-th call of ``rpc`` on ``peer`` (any peer if omitted) raises ``InjectedFault`` — ``times`` times in a row
(default 1). Example: ``PETALS_B200_FAULTS="rpc=rpc_inference,peer=stage1,after=3;rpc=rpc_backward,times=2"``.

Handlers call :func:`maybe_fail` at the top of every RPC; with no plan configured this is a dictionary lookup."""
from __future__ import annotations

import os
import threading
from dataclasses import dataclass
from typing import List, Optional


class InjectedFault(RuntimeError):
    """Raised in place of serving an RPC."""


@dataclass
class _Rule:
    rpc: str
    peer: Optional[str] = None
    after: int = 0
    times: int = 1
    error: str = "injected fault"
    seen: int = 0
    fired: int = 0


_lock = threading.Lock()
_rules: List[_Rule] = []


def set_fault_plan(plan: Optional[str]) -> None:
    """Replace the active plan (``None`` / empty clears it)."""
    rules: List[_Rule] = []
    for part in (plan or "").split(";"):
        part = part.strip()
        if not part:
            continue
        kv = dict(item.split("=", 1) for item in part.split(",") if "=" in item)
        if "rpc" not in kv:
            raise ValueError(f"fault rule without rpc=: {part!r}")
        rules.append(_Rule(rpc=kv["rpc"], peer=kv.get("peer"), after=int(kv.get("after", 0)), times=int(kv.get("times", 1)),
                           error=kv.get("error", "injected fault")))
    with _lock:
        _rules[:] = rules


def maybe_fail(rpc: str, peer_id: Optional[str] = None) -> None:
    if not _rules:
        return
    with _lock:
        for r in _rules:
            if r.rpc != rpc or (r.peer is not None and r.peer != peer_id):
                continue
            r.seen += 1
            if r.seen > r.after and r.fired < r.times:
                r.fired += 1
                raise InjectedFault(f"{r.error} ({rpc} on {peer_id}, call #{r.seen})")


def fired_count() -> int:
    with _lock:
        return sum(r.fired for r in _rules)


set_fault_plan(os.environ.get("PETALS_B200_FAULTS"))

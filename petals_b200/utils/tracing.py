"""Lightweight tracing hooks (the reference has none).

* ``nvtx_range(name)`` — NVTX ranges around stage steps / client steps when ``PETALS_B200_NVTX=1`` (visible in Nsight Systems
  and as range names in ``ncu``); a no-op otherwise, on CPU, or when NVTX is unavailable.
* ``StepTimer`` — wall-clock accumulation per label for the periodic stats line of a stage (``--stats_report_interval``)."""
from __future__ import annotations

import contextlib
import os
import threading
import time
from collections import defaultdict
from typing import Dict

_ENABLED = os.environ.get("PETALS_B200_NVTX", "0") not in ("", "0")


def nvtx_enabled() -> bool:
    return _ENABLED


@contextlib.contextmanager
def nvtx_range(name: str):
    if not _ENABLED:
        yield
        return
    try:
        import torch

        if not torch.cuda.is_available():
            yield
            return
        torch.cuda.nvtx.range_push(name)
    except Exception:  # noqa: BLE001 - tracing must never break serving
        yield
        return
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


class StepTimer:
    """Thread-safe (label -> count, seconds) accumulator."""

    def __init__(self):
        self._lock = threading.Lock()
        self._n: Dict[str, int] = defaultdict(int)
        self._t: Dict[str, float] = defaultdict(float)

    @contextlib.contextmanager
    def measure(self, label: str):
        t0 = time.perf_counter()
        try:
            yield
        finally:
            dt = time.perf_counter() - t0
            with self._lock:
                self._n[label] += 1
                self._t[label] += dt

    def snapshot(self, reset: bool = True) -> Dict[str, Dict[str, float]]:
        with self._lock:
            out = {k: {"count": self._n[k], "seconds": self._t[k]} for k in self._n}
            if reset:
                self._n.clear()
                self._t.clear()
        return out

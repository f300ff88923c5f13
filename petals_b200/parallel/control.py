"""Leader -> followers command ring in POSIX shared memory (the control plane of a worker group).

The reference pays a protobuf + libp2p round trip per step per server. Inside one box the control
messages of a tensor-parallel group ("open session", "step B x T at position p", "close") are a few dozen bytes:
the leader writes them into a single-producer / multi-consumer ring and followers spin on a sequence number
(micro-second latency, no syscalls). Payloads are msgpack. Data never travels here — only on NVLink."""
from __future__ import annotations

import struct
import time
from multiprocessing import shared_memory
from typing import Any, Optional

import msgpack
import numpy as np


def attach_shared_memory(name: str) -> shared_memory.SharedMemory:
    """Attach to a segment another process created WITHOUT adopting it: before Python 3.13 every attach registers the segment
    with this process's resource tracker, which then unlinks it a second time at exit ("leaked shared_memory objects",
    "No such file '/psm_...'"). Only the creator owns (and unlinks) the segment."""
    shm = shared_memory.SharedMemory(name=name, create=False)
    try:
        from multiprocessing import resource_tracker

        resource_tracker.unregister(shm._name, "shared_memory")  # noqa: SLF001 - the public `track=False` only exists from 3.13 on
    except Exception:  # noqa: BLE001
        pass
    return shm


class CommandRing:
    HEADER = 64  # bytes reserved for the write index

    def __init__(self, name: Optional[str], *, create: bool, n_consumers: int = 0, slots: int = 512, slot_bytes: int = 1024):
        if create:
            size = self.HEADER + 64 * n_consumers + slots * slot_bytes
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=size)
            self.shm.buf[:size] = b"\x00" * size
            self.shm.buf[8:32] = struct.pack("<QQQ", n_consumers, slots, slot_bytes)  # geometry travels with the ring
        else:
            self.shm = attach_shared_memory(name)
            n_consumers, slots, slot_bytes = struct.unpack("<QQQ", bytes(self.shm.buf[8:32]))
            size = self.HEADER + 64 * n_consumers + slots * slot_bytes
        self.n_consumers, self.slots, self.slot_bytes = int(n_consumers), int(slots), int(slot_bytes)
        self.name = self.shm.name
        self._u64 = np.ndarray((size // 8,), dtype=np.uint64, buffer=self.shm.buf)
        self._created = create
        self._data_off = self.HEADER + 64 * n_consumers

    # index helpers (u64 slots): write index at 0, consumer i's read index at (HEADER + 64*i) / 8
    def _widx(self) -> int:
        return int(self._u64[0])

    def _ridx(self, i: int) -> int:
        return int(self._u64[(self.HEADER + 64 * i) // 8])

    def send(self, obj: Any, timeout: float = 60.0) -> None:
        payload = msgpack.packb(obj, use_bin_type=True)
        if len(payload) + 4 > self.slot_bytes:
            raise ValueError(f"command of {len(payload)} bytes exceeds the slot size {self.slot_bytes}")
        w = self._widx()
        deadline = time.monotonic() + timeout
        while self.n_consumers and w - min(self._ridx(i) for i in range(self.n_consumers)) >= self.slots:
            if time.monotonic() > deadline:  # back-pressure: a follower fell a whole ring behind
                raise TimeoutError("command ring is full: a follower stopped consuming")
            time.sleep(0)
        off = self._data_off + (w % self.slots) * self.slot_bytes
        self.shm.buf[off + 4: off + 4 + len(payload)] = payload
        self.shm.buf[off: off + 4] = struct.pack("<I", len(payload))
        self._u64[0] = w + 1  # publish (x86 TSO keeps the payload stores before this one)

    def recv(self, consumer: int, timeout: Optional[float] = None) -> Any:
        r = self._ridx(consumer)
        deadline = None if timeout is None else time.monotonic() + timeout
        spins = 0
        while self._widx() <= r:
            spins += 1
            if spins > 2000:
                time.sleep(0)  # yield; the hot path (back-to-back decode steps) never gets here
            if deadline is not None and time.monotonic() > deadline:
                raise TimeoutError("no command from the leader")
        off = self._data_off + (r % self.slots) * self.slot_bytes
        (n,) = struct.unpack("<I", bytes(self.shm.buf[off: off + 4]))
        obj = msgpack.unpackb(bytes(self.shm.buf[off + 4: off + 4 + n]), raw=False)
        self._u64[(self.HEADER + 64 * consumer) // 8] = r + 1
        return obj

    def close(self) -> None:
        shm, self.shm, self._u64 = getattr(self, "shm", None), None, None
        if shm is None:  # idempotent: shutdown paths may reach here twice
            return
        try:
            shm.close()
            if self._created:
                shm.unlink()
        except Exception:  # noqa: BLE001
            pass

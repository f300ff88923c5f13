"""The NVLink activation fabric between pipeline stages (one process per GPU).

In the reference every hop between servers is an RPC that drags the activation through host memory, protobuf and the
libp2p daemon, and the router budgets 18 ms for it (src/petals/client/routing/sequence_manager.py:223).
Here each worker process owns a *landing zone* in a CUDA-IPC symmetric heap:

* ``x_in``  — where the previous stage's **last kernel** (down-projection GEMV or tcgen05 GEMM + residual) stores its
  output tiles directly over NVLink (``push_out`` epilogue), followed by one release-increment of ``in_flag``;
* ``y_ret`` — the same for the last stage returning the final hidden states to the client's GPU.

The consuming stage's **first kernel** waits on the flag (``ld.acquire.sys``) — so the hop costs one NVLink store stream
overlapped with the producer's math plus a flag latency, and the control RPC between processes carries *no tensor
bytes* (only "your input is in your landing zone"). Flags are monotonic; every consumer keeps a device-resident count
of consumed transfers (the "epoch" passed to the kernels), so nothing is ever reset and CUDA graphs stay valid.

Created once per process by :func:`init_fabric` after ``torch.distributed`` is initialised; absent (``get_fabric()`` is
None) in single-process / CPU runs, where stages exchange tensors by reference or over the Unix-socket transport."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from petals_b200.ops import native
from petals_b200.parallel.symmetric import SymmetricHeap, host_barrier, ptr_array, tensor_from_ptr
from petals_b200.utils.logging import get_logger

logger = get_logger(__name__)
_fabric: Optional["Fabric"] = None


class Fabric:
    def __init__(self, hidden_size: int, max_tokens: int = 8192, group=None, extra_bytes: int = (256 << 20) + (1 << 20)):
        self.hidden_size, self.max_tokens = hidden_size, max_tokens
        zone = max_tokens * hidden_size * 2
        self.heap = SymmetricHeap(2 * zone + extra_bytes, group=group)
        self.rank, self.world, self.device = self.heap.rank, self.heap.world, self.heap.device
        self.off_x_in = self.heap.alloc(zone)
        self.off_y_ret = self.heap.alloc(zone)
        self.off_flags = self.heap.alloc(64)
        self.x_in = self.heap.tensor(self.off_x_in, (max_tokens, hidden_size), torch.bfloat16)
        self.y_ret = self.heap.tensor(self.off_y_ret, (max_tokens, hidden_size), torch.bfloat16)
        # device-resident counts of consumed transfers (one per landing zone) + producer-side election counter
        self.in_epoch = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.ret_epoch = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.done_counter = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._flag_src = torch.zeros(16, dtype=torch.uint8, device=self.device)
        self._scratch_off = self.heap.alloc(64)
        # back-pressure: a producer may not overwrite a landing zone before the consumer copied the previous transfer out.
        # Every consumer acknowledges to the producer's ack flag; the pushing kernel's prologue waits for
        # ack >= number of pushes issued so far (the flag starts at 1, so the first push never waits).
        self.push_epoch = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.heap.tensor(self.off_flags + 16, (1,), torch.int64).fill_(1)
        torch.cuda.synchronize(self.device)
        host_barrier(group)

    # ---- addresses ---------------------------------------------------------------------------------------------
    def x_in_addr(self, rank: int) -> int:
        return self.heap.addr(rank, self.off_x_in)

    def y_ret_addr(self, rank: int) -> int:
        return self.heap.addr(rank, self.off_y_ret)

    def in_flag_addr(self, rank: int) -> int:
        return self.heap.addr(rank, self.off_flags)

    def ret_flag_addr(self, rank: int) -> int:
        return self.heap.addr(rank, self.off_flags + 8)

    def ack_flag_addr(self, rank: int) -> int:
        return self.heap.addr(rank, self.off_flags + 16)

    def begin_push(self) -> dict:
        """Call right before launching a kernel whose epilogue pushes into a peer's landing zone: returns the prologue
        wait arguments that implement the back-pressure described above."""
        native.check(native.lib().pb_bump_epoch(self.push_epoch.data_ptr(), native.stream_ptr()), "bump_epoch")
        return dict(wait_flag=self.ack_flag_addr(self.rank), wait_per_epoch=1, epoch=self.push_epoch.data_ptr(), error_flag=self.err.data_ptr())

    def take(self, M: int, kind: str, src_rank: int, out: torch.Tensor) -> torch.Tensor:
        """Consume the next transfer into this rank's ``kind`` zone: wait, copy the rows out, acknowledge to ``src_rank``."""
        self.wait(kind)
        buf = self.y_ret if kind == "y_ret" else self.x_in
        out.copy_(buf[:M])
        self._signal(self.ack_flag_addr(src_rank), src_rank)
        return out

    def _signal(self, flag_addr: int, rank: int) -> None:
        scratch = ptr_array([self.heap.addr(rank, self._scratch_off)])
        native.check(native.lib().pb_push_rows(self._flag_src.data_ptr(), scratch, ptr_array([flag_addr]), 1, 16, native.stream_ptr()), "fabric signal")

    def zone(self, kind: str, rank: int):
        """(data address, flag address) of a landing zone on ``rank``."""
        if kind == "x_in":
            return self.x_in_addr(rank), self.in_flag_addr(rank)
        if kind == "y_ret":
            return self.y_ret_addr(rank), self.ret_flag_addr(rank)
        raise ValueError(kind)

    # ---- host-issued transfers (client -> first stage; anything not produced by a fused epilogue) ---------------------
    def send(self, rows: torch.Tensor, rank: int, kind: str = "x_in") -> None:
        """Copy ``rows`` [M, H] into a landing zone of ``rank`` and publish it (stream ordered)."""
        M = rows.shape[0]
        if M > self.max_tokens:
            raise ValueError(f"{M} rows exceed the fabric landing zone ({self.max_tokens})")
        data, flag = self.zone(kind, rank)
        dst = tensor_from_ptr(data, (M, self.hidden_size), torch.bfloat16, self.device)
        kw = self.begin_push()  # honour back-pressure like the fused pushes do
        native.check(native.lib().pb_wait_flag(kw["wait_flag"], kw["epoch"], 1, 0, kw["error_flag"], native.stream_ptr()), "wait_flag")
        dst.copy_(rows.reshape(M, self.hidden_size).to(torch.bfloat16))  # P2P memcpy over NVLink
        self._signal(flag, rank)

    def wait(self, kind: str = "y_ret") -> None:
        """Enqueue a wait for the next transfer into this rank's zone (advances the zone's epoch)."""
        epoch = self.ret_epoch if kind == "y_ret" else self.in_epoch
        _, flag = self.zone(kind, self.rank)
        lib = native.lib()
        native.check(lib.pb_bump_epoch(epoch.data_ptr(), native.stream_ptr()), "bump_epoch")
        native.check(lib.pb_wait_flag(flag, epoch.data_ptr(), 1, 0, self.err.data_ptr(), native.stream_ptr()), "wait_flag")

    def recv(self, M: int, kind: str, src_rank: int) -> torch.Tensor:
        out = torch.empty(M, self.hidden_size, dtype=torch.bfloat16, device=self.device)
        return self.take(M, kind, src_rank, out)

    def check_errors(self) -> None:
        if int(self.err.item()):
            self.err.zero_()
            raise RuntimeError(f"rank {self.rank}: a fabric flag wait timed out (a peer stage is gone or stalled)")

    def close(self) -> None:
        self.heap.close()


class HostFabric:
    """The same protocol over POSIX shared memory, for CPU processes (gloo): lets the multi-process plumbing tests exercise
    the fabric code paths of the client and the handlers (landing zones, flags, acknowledgements) without GPUs."""

    def __init__(self, hidden_size: int, max_tokens: int = 1024, group=None, dtype: torch.dtype = torch.float32):
        import time
        from multiprocessing import shared_memory

        import numpy as np

        self.hidden_size, self.max_tokens, self.dtype = hidden_size, max_tokens, dtype
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device("cpu")
        self._time, self._np = time, np
        itemsize = torch.empty(0, dtype=dtype).element_size()
        self._zone_bytes = max_tokens * hidden_size * itemsize
        per_rank = 2 * self._zone_bytes + 64
        names = [None]
        if self.rank == 0:
            self._shm = shared_memory.SharedMemory(create=True, size=per_rank * self.world)
            self._shm.buf[: per_rank * self.world] = b"\x00" * (per_rank * self.world)
            names[0] = self._shm.name
        dist.broadcast_object_list(names, src=0, group=group)
        if self.rank != 0:
            from petals_b200.parallel.control import attach_shared_memory

            self._shm = attach_shared_memory(names[0])
        self._per_rank = per_rank
        self._consumed = {"x_in": 0, "y_ret": 0}
        self._pushes = 0
        for r in range(self.world):
            self._flags(r)[2] = 1  # ack flags start at 1 (first push never waits)
        host_barrier(group)

    def _flags(self, rank: int):
        off = rank * self._per_rank + 2 * self._zone_bytes
        return self._np.ndarray((8,), dtype=self._np.uint64, buffer=self._shm.buf, offset=off)

    def _zone(self, kind: str, rank: int, rows: int) -> torch.Tensor:
        off = rank * self._per_rank + (0 if kind == "x_in" else self._zone_bytes)
        itemsize = torch.empty(0, dtype=self.dtype).element_size()
        flat = torch.frombuffer(self._shm.buf, dtype=self.dtype, count=rows * self.hidden_size, offset=off)
        return flat.view(rows, self.hidden_size)

    def _spin(self, rank: int, idx: int, target: int, what: str, timeout: float = 30.0) -> None:
        deadline = self._time.monotonic() + timeout
        while int(self._flags(rank)[idx]) < target:
            if self._time.monotonic() > deadline:
                raise TimeoutError(f"rank {self.rank}: timed out waiting for {what}")
            self._time.sleep(0)

    def send(self, rows: torch.Tensor, rank: int, kind: str = "x_in") -> None:
        M = rows.shape[0]
        if M > self.max_tokens:
            raise ValueError(f"{M} rows exceed the fabric landing zone ({self.max_tokens})")
        self._pushes += 1
        self._spin(self.rank, 2, self._pushes, "the consumer's acknowledgement")
        self._zone(kind, rank, M).copy_(rows.reshape(M, self.hidden_size).to(self.dtype))
        self._flags(rank)[0 if kind == "x_in" else 1] += 1

    def take(self, M: int, kind: str, src_rank: int, out: torch.Tensor) -> torch.Tensor:
        self._consumed[kind] += 1
        self._spin(self.rank, 0 if kind == "x_in" else 1, self._consumed[kind], f"a transfer into {kind}")
        out.copy_(self._zone(kind, self.rank, M))
        self._flags(src_rank)[2] += 1
        return out

    def recv(self, M: int, kind: str, src_rank: int) -> torch.Tensor:
        return self.take(M, kind, src_rank, torch.empty(M, self.hidden_size, dtype=self.dtype))

    def check_errors(self) -> None:
        pass

    def close(self) -> None:
        shm, self._shm = getattr(self, "_shm", None), None
        if shm is None:
            return
        try:
            shm.close()
            if self.rank == 0:
                shm.unlink()
        except Exception:  # noqa: BLE001
            pass


def init_fabric(hidden_size: int, max_tokens: int = 8192, group=None, host_dtype: torch.dtype = torch.float32):
    """Collective over ``group``. Returns None when there is nothing to connect (single process)."""
    global _fabric
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2:
        return None
    if torch.cuda.is_available():
        _fabric = Fabric(hidden_size, max_tokens, group)
    else:
        _fabric = HostFabric(hidden_size, min(max_tokens, 1024), group, host_dtype)
    return _fabric


def get_fabric() -> Optional[Fabric]:
    return _fabric

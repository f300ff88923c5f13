"""NVLink symmetric memory for one-process-per-GPU worker groups.

Every rank allocates one heap with ``cudaMalloc``, exports it with CUDA IPC, and maps all peers' heaps
(csrc/ipc.cu). Allocation is a bump allocator executed identically on every rank, so an allocation is an *offset*
valid in every heap: ``heap.addr(rank, off)`` is a raw device pointer that this rank's kernels can store to (peer
pushes fused into GEMV/GEMM epilogues) and ``heap.tensor(off, ...)`` is a torch view of the local copy.

``torch.distributed`` (NCCL or gloo) is only the *plumbing* — it carries the 64-byte IPC handles once at start-up
and barriers; no collective is issued on the token path."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from petals_b200.ops import native
from petals_b200.utils.logging import get_logger

logger = get_logger(__name__)


def host_barrier(group=None) -> None:
    """A barrier that never touches the GPU: an all-reduce of one CPU scalar (gloo).

    ``dist.barrier()`` on a NCCL-backed group parks a spinning kernel on every waiting rank's GPU. A stage that keeps serving
    from background threads while its main thread sits in such a barrier dead-locks on its first device-wide
    synchronisation (CUDA-graph capture, ``cudaFree``): the sync waits for the barrier kernel, the barrier waits for the
    client, the client waits for the stage. Host barriers order *processes*; callers synchronise their own device first
    where GPU work must be finished."""
    try:
        dist.all_reduce(torch.zeros(1), group=group)
    except RuntimeError:  # group without a CPU backend
        dist.barrier(group=group)


class _RawCudaBuffer:
    """Adapter exposing a raw device pointer through __cuda_array_interface__ (uint8)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


def tensor_from_ptr(ptr: int, shape: Sequence[int], dtype: torch.dtype, device) -> torch.Tensor:
    n = 1
    for s in shape:
        n *= s
    nbytes = n * torch.empty(0, dtype=dtype).element_size()
    raw = torch.as_tensor(_RawCudaBuffer(ptr, nbytes), device=device)
    return raw.view(dtype).view(*shape)


class SymmetricHeap:
    """``multicast_ptr`` (0 when unavailable): an NVSwitch multicast mapping of the same heap — a ``multimem.st`` to
    ``multicast_ptr + off`` lands at offset ``off`` of EVERY rank's heap in one store (NVLS), a ``multimem.ld_reduce`` reads the
    switch-side sum of all ranks' copies. The CUDA VMM / multicast-object / file-descriptor plumbing is torch's symmetric-memory
    allocator (``torch.distributed._symmetric_memory``); when it cannot be set up (no NVSwitch, old driver, CPU build) the heap
    is a plain ``cudaMalloc`` exported with CUDA IPC and ``multicast_ptr`` stays 0."""

    def _try_symm_mem(self, group) -> bool:
        import os

        if os.environ.get("PETALS_B200_SYMM_MEM", "1") == "0":
            return False
        try:
            import torch.distributed._symmetric_memory as symm_mem

            with torch.cuda.device(self.device):
                buf = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=self.device)
                hdl = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
            ptrs = [int(x) for x in hdl.buffer_ptrs]
            if len(ptrs) != self.world or not all(ptrs):
                return False
            self._symm_buf, self._symm_hdl = buf, hdl  # keep the mapping alive
            self.ptrs, self.local_ptr = ptrs, ptrs[self.rank]
            self.multicast_ptr = int(getattr(hdl, "multicast_ptr", 0) or 0)
            return True
        except Exception as e:  # noqa: BLE001 - any failure means: use the IPC heap
            logger.info(f"rank {dist.get_rank(group)}: symmetric-memory allocator unavailable ({type(e).__name__}: {str(e)[:120]}); using CUDA IPC")
            return False

    def __init__(self, nbytes: int, group: Optional[dist.ProcessGroup] = None, device: Optional[torch.device] = None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.nbytes = (nbytes + 4095) // 4096 * 4096
        self._lib = native.lib()
        self.multicast_ptr = 0
        self._symm_buf = self._symm_hdl = None
        # every rank must take the same branch: agree on the outcome
        ok = torch.tensor([1 if self._try_symm_mem(group) else 0])
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 1:
            mc = torch.tensor([1 if self.multicast_ptr else 0])
            dist.all_reduce(mc, op=dist.ReduceOp.MIN, group=group)
            if int(mc.item()) == 0:
                self.multicast_ptr = 0
            self._top = 0
            self.tensor(0, (self.nbytes,), torch.uint8).zero_()
            torch.cuda.synchronize(self.device)
            host_barrier(group)
            logger.info(f"rank {self.rank}: symmetric heap of {self.nbytes >> 20} MiB mapped on {self.world} ranks (VMM, multicast {'on' if self.multicast_ptr else 'off'})")
            return
        self._symm_buf = self._symm_hdl = None
        self.multicast_ptr = 0
        base = C.c_void_p()
        with torch.cuda.device(self.device):
            native.check(self._lib.pb_ipc_malloc(C.byref(base), self.nbytes), "ipc_malloc", 0)
            self.local_ptr = int(base.value)
            hsize = self._lib.pb_ipc_handle_size()
            buf = C.create_string_buffer(hsize)
            native.check(self._lib.pb_ipc_get_handle(self.local_ptr, buf), "ipc_get_handle", 0)
            handles: List[Optional[bytes]] = [None] * self.world
            dist.all_gather_object(handles, bytes(buf.raw), group=group)
            self.ptrs: List[int] = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    self.ptrs.append(self.local_ptr)
                    continue
                out = C.c_void_p()
                native.check(self._lib.pb_ipc_open_handle(h, C.byref(out)), f"ipc_open_handle(rank {r})", 0)
                self.ptrs.append(int(out.value))
        self._top = 0
        host_barrier(group)
        logger.info(f"rank {self.rank}: symmetric heap of {self.nbytes >> 20} MiB mapped on {self.world} ranks")

    # ---- allocation (must be called in the same order on every rank) ---------------------------------------------
    def alloc(self, nbytes: int, align: int = 256) -> int:
        off = (self._top + align - 1) // align * align
        if off + nbytes > self.nbytes:
            raise MemoryError(f"symmetric heap exhausted ({off + nbytes} > {self.nbytes})")
        self._top = off + nbytes
        return off

    def alloc_tensor(self, shape: Sequence[int], dtype: torch.dtype) -> Tuple[int, torch.Tensor]:
        n = 1
        for s in shape:
            n *= s
        off = self.alloc(n * torch.empty(0, dtype=dtype).element_size())
        return off, self.tensor(off, shape, dtype)

    def addr(self, rank: int, off: int) -> int:
        return self.ptrs[rank] + off

    def mc_addr(self, off: int) -> int:
        """Multicast address of offset ``off`` (0 when the heap has no multicast mapping)."""
        return self.multicast_ptr + off if self.multicast_ptr else 0

    def tensor(self, off: int, shape: Sequence[int], dtype: torch.dtype) -> torch.Tensor:
        return tensor_from_ptr(self.local_ptr + off, shape, dtype, self.device)

    def zero_(self) -> None:
        self.tensor(0, (self.nbytes,), torch.uint8).zero_()
        torch.cuda.synchronize(self.device)
        host_barrier(self.group)

    def close(self) -> None:
        if self._symm_hdl is not None:  # released with the tensor / handle
            self._symm_hdl = self._symm_buf = None
            self.local_ptr = 0
            return
        for r, p in enumerate(self.ptrs):
            if r != self.rank and p:
                self._lib.pb_ipc_close_handle(p)
        if self.local_ptr:
            self._lib.pb_ipc_free(self.local_ptr)
            self.local_ptr = 0


def ptr_array(ptrs: Sequence[Optional[int]]):
    arr = (C.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


# ---- probes: hop latency and peer bandwidth (roofline denominators for the fused paths) -------------------------------
def measure_peer_bandwidth(heap: SymmetricHeap, src_rank: int = 0, dst_rank: int = 1, nbytes: int = 256 << 20, iters: int = 10) -> Optional[float]:
    """GB/s of SM-issued stores from ``src_rank`` into ``dst_rank``'s heap (what a fused epilogue push can reach).

    Collective: allocates its own scratch region in the symmetric heap (never touches live buffers or flags)."""
    nbytes = min(nbytes, (heap.nbytes - heap._top - 4096)) // 4096 * 4096
    if nbytes < (1 << 20):
        host_barrier(heap.group)
        return None
    region = heap.alloc(nbytes, align=4096)
    result = None
    if heap.rank == src_rank:
        src = torch.empty(nbytes, dtype=torch.uint8, device=heap.device)
        lib = native.lib()
        for _ in range(3):
            lib.pb_peer_copy(src.data_ptr(), heap.addr(dst_rank, region), nbytes, 0, native.stream_ptr())
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            lib.pb_peer_copy(src.data_ptr(), heap.addr(dst_rank, region), nbytes, 0, native.stream_ptr())
        e.record()
        torch.cuda.synchronize()
        result = nbytes * iters / (s.elapsed_time(e) * 1e-3) / 1e9
    host_barrier(heap.group)
    return result


def measure_hop_latency(heap: SymmetricHeap, flag_off: int, a: int = 0, b: int = 1, iters: int = 200) -> Optional[float]:
    """Median one-way flag latency (us) between ranks a and b: release-store + acquire-spin over NVLink."""
    lib = native.lib()
    heap.tensor(flag_off, (1,), torch.int64).zero_()
    torch.cuda.synchronize()
    host_barrier(heap.group)
    result = None
    if heap.rank in (a, b):
        peer = b if heap.rank == a else a
        rtt = torch.zeros(iters, dtype=torch.int64, device=heap.device)
        err = torch.zeros(1, dtype=torch.int32, device=heap.device)
        native.check(lib.pb_pingpong(heap.addr(heap.rank, flag_off), heap.addr(peer, flag_off), iters, int(heap.rank == a), rtt.data_ptr(),
                                     err.data_ptr(), native.stream_ptr()), "pingpong")
        torch.cuda.synchronize()
        if int(err.item()):
            raise RuntimeError("ping-pong watchdog expired")
        if heap.rank == a:
            result = float(rtt[iters // 10:].float().median().item()) / 2e3
    host_barrier(heap.group)
    return result

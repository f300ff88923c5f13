"""KV-cache memory manager of one stage (reference: src/petals/server/memory_cache.py:26-225).

The reference reserves *bytes* per session under multiprocessing locks and materialises dense
``[B, H, D, Lmax]`` tensors lazily in the runtime process. Here the cache is **paged**: one pool tensor
``[n_blocks, 2, num_pages, Hkv, 64, D]`` per stage, sessions *reserve* pages up front (admission control
with ``alloc_timeout``, FIFO-fair, implemented natively in csrc/runtime/kv_allocator.cpp) and *bind*
physical pages lazily as tokens arrive. A page id is valid for every block of the span, so a session has
one block table shared by all its layers. Rollback (speculative decoding) only moves the position,
beam-search reordering permutes block-table rows with refcounted page sharing and copy-on-write of the
page being written (replacing the whole-cache gather at src/petals/server/backend.py:154-158).

On CPU (plumbing tests) the same interface is backed by dense per-block tensors.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import threading
import time
from typing import Dict, List, Optional, Sequence

import torch

from petals_b200.ops import native
from petals_b200.ops.functional import PAGE
from petals_b200.utils.logging import get_logger

logger = get_logger(__name__)


class AllocationFailed(Exception):
    pass


class SessionCache:
    """KV state of one inference session on one stage."""

    def __init__(self, owner: "MemoryCache", batch_size: int, max_length: int, reserved_pages: int):
        self.owner, self.batch_size, self.max_length = owner, batch_size, max_length
        self.reserved_pages = reserved_pages
        self.position = 0
        self.closed = False
        self.on_close = []  # callbacks(session): e.g. a TP leader tells its followers to drop their mirror session
        self.pages_per_seq = (max_length + PAGE - 1) // PAGE
        if owner.paged:
            self.tables: List[List[int]] = [[] for _ in range(batch_size)]
            dev = owner.device
            with torch.inference_mode(False):  # sessions are opened and stepped from different threads
                self.table_dev = torch.zeros(batch_size, owner.max_pages_per_seq, dtype=torch.int32, device=dev)
            self._table_dirty = False
            self._version, self._synced_version = 0, -1
        else:
            self.dense: Dict[int, tuple] = {}  # block slot -> (k, v) [B, max_length, Hkv, D]

    # ---- paged (GPU) -----------------------------------------------------------------------------------
    def _alloc(self, n: int) -> List[int]:
        buf = (C.c_int * n)()
        if self.owner._rt.pb_kv_alloc(self.owner._alloc, n, buf) != 0:
            raise AllocationFailed("KV page pool exhausted despite reservation (internal accounting error)")
        return list(buf)

    def _free(self, pages: Sequence[int]) -> None:
        if pages:
            arr = (C.c_int * len(pages))(*pages)
            self.owner._rt.pb_kv_free(self.owner._alloc, arr, len(pages))

    def prepare_write(self, n_new_tokens: int) -> None:
        """Make positions [position, position + n) writable: bind pages, copy-on-write shared ones."""
        new_len = self.position + n_new_tokens
        if new_len > self.max_length:
            raise ValueError(f"Maximum length exceeded: prefix {self.position} + current {n_new_tokens} exceeds pre-allocated maximum {self.max_length}")
        if not self.owner.paged or n_new_tokens == 0:
            return
        need = (new_len + PAGE - 1) // PAGE
        first_written = self.position // PAGE
        cow_src, cow_dst = [], []
        for b, table in enumerate(self.tables):
            for idx in range(first_written, min(len(table), need)):
                pg = table[idx]
                if self.owner._rt.pb_kv_refcount(self.owner._alloc, pg) > 1:  # shared with another hypothesis
                    fresh = self._alloc(1)[0]
                    cow_src.append(pg), cow_dst.append(fresh)
                    self._free([pg])
                    table[idx] = fresh
                    self._table_dirty = True
            if len(table) < need:
                table.extend(self._alloc(need - len(table)))
                self._table_dirty = True
        if cow_src:
            self.owner.copy_pages(cow_src, cow_dst)
        self.sync_table()

    def sync_table(self) -> None:
        if self.owner.paged and self._table_dirty:
            host = torch.zeros(self.table_dev.shape, dtype=torch.int32)
            for b, table in enumerate(self.tables):
                if table:
                    host[b, : len(table)] = torch.tensor(table, dtype=torch.int32)
            self.table_dev.copy_(host, non_blocking=False)
            self._table_dirty = False
            self._version += 1

    def set_position(self, pos: int, sync_device: bool = True) -> None:
        if pos < 0 or pos > self.max_length:
            raise ValueError(f"position {pos} outside [0, {self.max_length}]")
        self.position = pos

    def reorder(self, hypo_ids: torch.Tensor) -> None:
        """Beam search: sequence b continues hypothesis hypo_ids[b] (reference backend.py:154-158)."""
        ids = [int(i) for i in hypo_ids.tolist()]
        if ids == list(range(self.batch_size)):
            return
        if self.owner.paged:
            new_tables = [list(self.tables[i]) for i in ids]
            flat = [p for t in new_tables for p in t]
            if flat:
                arr = (C.c_int * len(flat))(*flat)
                self.owner._rt.pb_kv_incref(self.owner._alloc, arr, len(flat))
            for t in self.tables:
                self._free(t)
            self.tables = new_tables
            self._table_dirty = True
            self.sync_table()
        else:
            index = torch.tensor(ids)
            for slot, (k, v) in self.dense.items():
                self.dense[slot] = (k.index_select(0, index.to(k.device)).contiguous(), v.index_select(0, index.to(v.device)).contiguous())

    # ---- dense (CPU oracle) ---------------------------------------------------------------------------
    def dense_kv(self, slot: int, spec, dtype: torch.dtype, device) -> tuple:
        if slot not in self.dense:
            shape = (self.batch_size, self.max_length, spec.num_kv_heads, spec.head_dim)
            with torch.inference_mode(False):
                self.dense[slot] = (torch.zeros(shape, dtype=dtype, device=device), torch.zeros(shape, dtype=dtype, device=device))
        return self.dense[slot]

    def close(self) -> None:
        if self.closed:
            return
        self.closed = True
        for cb in self.on_close:
            try:
                cb(self)
            except Exception:  # noqa: BLE001
                pass
        if self.owner.paged:
            for t in self.tables:
                self._free(t)
            self.tables = []
        else:
            self.dense.clear()
        self.owner._release(self)


class MemoryCache:
    """Page pool + admission control for one stage."""

    def __init__(self, max_size_tokens: int, max_alloc_timeout: Optional[float] = None, *, n_blocks: int = 1,
                 spec=None, dtype: torch.dtype = torch.bfloat16, device="cpu", paged: Optional[bool] = None,
                 max_length: int = 8192):
        self.device = torch.device(device)
        self.paged = self.device.type == "cuda" if paged is None else paged
        self.max_pages_per_seq = (max_length + PAGE - 1) // PAGE + 1
        self.n_blocks, self.spec, self.dtype = n_blocks, spec, dtype
        self.num_pages = max(1, (int(max_size_tokens) + PAGE - 1) // PAGE)
        self.max_alloc_timeout = max_alloc_timeout
        self._rt = native.rt()
        self._alloc = self._rt.pb_kv_create(self.num_pages)
        self._lock = threading.Lock()
        self._sessions: List[SessionCache] = []
        self.pool: Optional[torch.Tensor] = None
        if self.paged:
            assert spec is not None
            self.pool = torch.zeros(n_blocks, 2, self.num_pages, spec.num_kv_heads, PAGE, spec.head_dim, dtype=dtype, device=self.device)
            logger.info(f"KV pool: {self.num_pages} pages x {PAGE} tokens x {n_blocks} blocks = {self.pool.numel() * self.pool.element_size() / 2**30:.2f} GiB")

    # ---- accounting (reference: memory_cache.py:59-61, handler.py:582) -----------------------------------
    @property
    def tokens_left(self) -> int:
        return max(0, self.num_pages - int(self._rt.pb_kv_reserved(self._alloc))) * PAGE

    @property
    def max_size_tokens(self) -> int:
        return self.num_pages * PAGE

    @property
    def bytes_per_token(self) -> int:
        if self.spec is None:
            return 1
        return self.spec.kv_bytes_per_token(self.dtype) * self.n_blocks

    @property
    def bytes_left(self) -> int:
        return self.tokens_left * self.bytes_per_token

    @property
    def current_size_bytes(self) -> int:
        return int(self._rt.pb_kv_reserved(self._alloc)) * PAGE * self.bytes_per_token

    @staticmethod
    def pages_needed(batch_size: int, max_length: int) -> int:
        # +1 page per sequence: slack for copy-on-write of the page under the write head
        return batch_size * ((max_length + PAGE - 1) // PAGE + 1)

    @contextlib.contextmanager
    def allocate_cache(self, batch_size: int, max_length: int, timeout: Optional[float] = 0.0):
        """Reserve KV memory for a session; waits up to ``timeout`` s (bounded by max_alloc_timeout).

        ``timeout=0`` fails fast, like the reference's default client ``alloc_timeout`` (memory_cache.py:89-90)."""
        session = self.open_session(batch_size, max_length, timeout)
        try:
            yield session
        finally:
            session.close()

    def open_session(self, batch_size: int, max_length: int, timeout: Optional[float] = 0.0) -> SessionCache:
        pages = self.pages_needed(batch_size, max_length)
        if max_length > (self.max_pages_per_seq - 1) * PAGE:
            raise AllocationFailed(f"max_length={max_length} exceeds this stage's inference_max_length={(self.max_pages_per_seq - 1) * PAGE}")
        if pages > self.num_pages:
            raise AllocationFailed(f"Could not allocate {pages * PAGE} tokens of KV cache: the stage only has {self.num_pages * PAGE}")
        if timeout is None:
            timeout = self.max_alloc_timeout if self.max_alloc_timeout is not None else 1e9
        if self.max_alloc_timeout is not None:
            timeout = min(timeout, self.max_alloc_timeout)
        t0 = time.perf_counter()
        if self._rt.pb_kv_reserve(self._alloc, pages, float(timeout)) != 0:
            raise AllocationFailed(f"Could not allocate {pages * PAGE} cache tokens within {timeout} seconds "
                                   f"({self.tokens_left} left of {self.max_size_tokens})")
        waited = time.perf_counter() - t0
        if waited > 0.1:
            logger.info(f"KV reservation of {pages} pages waited {waited:.2f}s")
        session = SessionCache(self, batch_size, max_length, pages)
        with self._lock:
            self._sessions.append(session)
        return session

    def _release(self, session: SessionCache) -> None:
        with self._lock:
            if session in self._sessions:
                self._sessions.remove(session)
        self._rt.pb_kv_unreserve(self._alloc, session.reserved_pages)

    def copy_pages(self, src: Sequence[int], dst: Sequence[int]) -> None:
        if self.device.type != "cuda":  # host pools (tests of the page bookkeeping): same semantics with tensor indexing
            self.pool[:, :, list(dst)] = self.pool[:, :, list(src)]
            return
        s = torch.tensor(src, dtype=torch.int32, device=self.device)
        d = torch.tensor(dst, dtype=torch.int32, device=self.device)
        page_elems = self.pool.shape[3] * self.pool.shape[4] * self.pool.shape[5]
        slab_stride = self.pool.shape[2] * page_elems
        native.check(native.lib().pb_kv_copy_pages(self.pool.data_ptr(), s.data_ptr(), d.data_ptr(), len(src), page_elems,
                                                   slab_stride, self.pool.shape[0] * 2, native.stream_ptr()), "kv_copy_pages")

    def layer_pools(self, slot: int) -> tuple:
        return self.pool[slot, 0], self.pool[slot, 1]

    def __del__(self):
        try:
            self._rt.pb_kv_destroy(self._alloc)
        except Exception:
            pass

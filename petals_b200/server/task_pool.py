"""Prioritised task pools and the per-GPU runtime that drains them
(reference: src/petals/server/task_pool.py:17-177 + hivemind's Runtime loop).

The reference moves tasks between handler *processes* and the runtime thread through mp queues,
shared-memory tensors and MPFutures, and copies every batch host->device->host. Here a stage is one process
per GPU: tasks carry device tensors by reference, ordering lives in a native priority queue
(csrc/runtime/task_queue.cpp) and an uncontended submit executes inline in the caller's thread (no thread
hop on the single-stream latency path). Semantics kept: smallest (priority, submission order) first, no
batching across requests, tasks above ``max_batch_size`` tokens are rejected."""
from __future__ import annotations

import ctypes as C
import threading
import time
from concurrent.futures import Future
from typing import Any, Callable, Dict, Optional, Sequence

import torch

from petals_b200.ops import native
from petals_b200.utils.logging import get_logger

logger = get_logger(__name__)


class Runtime(threading.Thread):
    """Executes tasks of all pools of one stage, one at a time, in priority order."""

    def __init__(self, name: str = "runtime", stats_report_interval: Optional[float] = None, device=None):
        super().__init__(name=name, daemon=True)
        self._rt = native.rt()
        self._queue = self._rt.pb_tq_create()
        self._tasks: Dict[int, tuple] = {}
        self._tasks_lock = threading.Lock()
        self._exec_lock = threading.Lock()
        self._next_id = 0
        self._shutdown = threading.Event()
        self.ready = threading.Event()
        self.device = device
        self.stats_report_interval = stats_report_interval
        self.stats: Dict[str, list] = {}
        self.allow_inline = True

    # -- submission ------------------------------------------------------------------------------------
    def submit(self, pool: "PrioritizedTaskPool", priority: float, args: Sequence[Any], in_caller_thread: bool = False) -> Future:
        fut: Future = Future()
        if self._shutdown.is_set():
            fut.set_exception(RuntimeError("runtime is shut down"))
            return fut
        if in_caller_thread:
            # Tasks that use torch.autograd (span backward) must run in the submitting thread: when that thread is
            # a CUDA autograd worker, the nested autograd call is re-entrant only from the same thread.
            with self._exec_lock:
                self._execute(pool, args, fut)
            return fut
        if self.allow_inline and self._rt.pb_tq_size(self._queue) == 0 and self._exec_lock.acquire(blocking=False):
            try:
                self._execute(pool, args, fut)
            finally:
                self._exec_lock.release()
            return fut
        with self._tasks_lock:
            task_id = self._next_id
            self._next_id += 1
            self._tasks[task_id] = (pool, args, fut)
        self._rt.pb_tq_push(self._queue, float(priority), task_id)
        return fut

    def _execute(self, pool: "PrioritizedTaskPool", args: Sequence[Any], fut: Future) -> None:
        if not fut.set_running_or_notify_cancel():
            return
        t0 = time.perf_counter()
        try:
            if self.device is not None and torch.device(self.device).type == "cuda":
                with torch.cuda.device(self.device):
                    result = pool.process_func(*args)
            else:
                result = pool.process_func(*args)
            fut.set_result(result)
        except BaseException as e:  # noqa: BLE001 - propagate everything to the requester
            fut.set_exception(e)
        finally:
            if self.stats_report_interval is not None:
                self.stats.setdefault(pool.name, []).append(time.perf_counter() - t0)

    # -- thread loop -------------------------------------------------------------------------------------
    def run(self) -> None:
        self.ready.set()
        task_id, prio = C.c_int64(), C.c_double()
        last_report = time.perf_counter()
        while not self._shutdown.is_set():
            rc = self._rt.pb_tq_pop(self._queue, 0.1, C.byref(task_id), C.byref(prio))
            if rc == -2:
                break
            if rc == 0:
                with self._tasks_lock:
                    pool, args, fut = self._tasks.pop(task_id.value)
                with self._exec_lock:
                    self._execute(pool, args, fut)
            if self.stats_report_interval and time.perf_counter() - last_report > self.stats_report_interval:
                last_report = time.perf_counter()
                for name, times in self.stats.items():
                    if times:
                        logger.info(f"{name}: {len(times)} tasks, mean {1e3 * sum(times) / len(times):.2f} ms")
                self.stats = {}

    def shutdown(self) -> None:
        self._shutdown.set()
        self._rt.pb_tq_close(self._queue)
        with self._tasks_lock:
            pending = list(self._tasks.values())
            self._tasks.clear()
        for _, _, fut in pending:
            if not fut.done():
                fut.set_exception(RuntimeError("runtime shut down before the task ran"))
        if self.is_alive() and threading.current_thread() is not self:
            self.join(timeout=5)

    @property
    def queue_size(self) -> int:
        return self._rt.pb_tq_size(self._queue)


class PrioritizedTaskPool:
    """A named entry point (inference / forward / backward of a block or span) into the runtime."""

    def __init__(self, process_func: Callable[..., Any], max_batch_size: int, name: str, runtime: Optional[Runtime] = None,
                 min_batch_size: int = 1, device=None, in_caller_thread: bool = False):
        self.in_caller_thread = in_caller_thread
        if min_batch_size != 1:
            raise ValueError("batching across requests is not supported (min_batch_size must be 1)")
        self.process_func, self.max_batch_size, self.name = process_func, max_batch_size, name
        self.runtime = runtime
        self.device = device

    def attach(self, runtime: Runtime) -> None:
        self.runtime = runtime

    @staticmethod
    def get_task_size(*args: Any) -> int:
        """Size of a task in tokens (batch x sequence of the first tensor), reference task_pool.py:113-117."""
        for a in args:
            if isinstance(a, torch.Tensor) and a.dim() >= 2:
                return int(a.shape[0] * a.shape[1])
        return 1

    def submit_task(self, *args: Any, priority: float = 0.0, size: Optional[int] = None) -> Future:
        task_size = self.get_task_size(*args) if size is None else size
        if task_size > self.max_batch_size:
            fut: Future = Future()
            fut.set_exception(ValueError(f"Task size ({task_size}) exceeds max_batch_size ({self.max_batch_size})"))
            return fut
        if self.runtime is None:
            raise RuntimeError(f"pool {self.name} is not attached to a runtime")
        return self.runtime.submit(self, priority, args, in_caller_thread=self.in_caller_thread)

"""Task ordering across pools (reference tests/test_priority_pool.py:31-82): smaller priority first, FIFO among equals,
size limit, errors propagate, inline fast path when idle."""
import threading
import time

import pytest
import torch

from petals_b200.server.task_pool import PrioritizedTaskPool, Runtime
from petals_b200.server.task_prioritizer import DummyTaskPrioritizer


def test_priority_order_across_pools():
    runtime = Runtime()
    runtime.allow_inline = False
    done = []

    def work(tag, delay=0.01):
        time.sleep(delay)
        done.append(tag)
        return tag

    inference = PrioritizedTaskPool(work, max_batch_size=100, name="inference", runtime=runtime)
    forward = PrioritizedTaskPool(work, max_batch_size=100, name="forward", runtime=runtime)
    prio = DummyTaskPrioritizer()
    futs = []
    for i in range(3):
        futs.append(forward.submit_task(f"fwd{i}", priority=prio.prioritize(points=0, type="forward")))
    for i in range(3):
        futs.append(inference.submit_task(f"inf{i}", priority=prio.prioritize(points=0, type="inference")))
    futs.append(forward.submit_task("urgent", priority=0.5))
    runtime.start()
    assert [f.result(timeout=5) for f in futs] == ["fwd0", "fwd1", "fwd2", "inf0", "inf1", "inf2", "urgent"]
    assert done == ["urgent", "inf0", "inf1", "inf2", "fwd0", "fwd1", "fwd2"]
    runtime.shutdown()


def test_limits_errors_and_inline():
    runtime = Runtime()
    runtime.start()
    pool = PrioritizedTaskPool(lambda x: x * 2, max_batch_size=8, name="p", runtime=runtime)
    assert pool.submit_task(torch.ones(2, 3, 4)).result(timeout=5).sum() == 48  # 6 tokens <= 8
    with pytest.raises(ValueError, match="exceeds max_batch_size"):
        pool.submit_task(torch.ones(3, 3, 4)).result(timeout=5)
    boom = PrioritizedTaskPool(lambda: 1 / 0, max_batch_size=8, name="boom", runtime=runtime)
    with pytest.raises(ZeroDivisionError):
        boom.submit_task().result(timeout=5)
    # idle runtime => the task runs in the caller's thread (no hop on the latency path)
    ident = PrioritizedTaskPool(lambda: threading.get_ident(), max_batch_size=8, name="ident", runtime=runtime)
    assert ident.submit_task().result(timeout=5) == threading.get_ident()
    runtime.shutdown()
    with pytest.raises(RuntimeError):
        pool.submit_task(torch.ones(1, 1, 1)).result(timeout=5)

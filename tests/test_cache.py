"""MemoryCache admission control (reference tests/test_cache.py:24-184): budgets, timeouts, fail-fast, FIFO queueing,
release on close; plus paged-table behaviours the reference does not have (rollback, beam reorder bookkeeping)."""
import threading
import time

import pytest
import torch

from petals_b200.ops.functional import PAGE
from petals_b200.server.memory_cache import AllocationFailed, MemoryCache


def test_budget_and_release():
    cache = MemoryCache(max_size_tokens=8 * PAGE, max_alloc_timeout=0.5, device="cpu")
    assert cache.tokens_left == 8 * PAGE
    with cache.allocate_cache(1, 2 * PAGE, timeout=0) as a:  # 2 pages + 1 slack
        assert cache.tokens_left == 5 * PAGE
        with cache.allocate_cache(2, PAGE, timeout=0):  # 2 * (1 + 1) pages
            assert cache.tokens_left == PAGE
            with pytest.raises(AllocationFailed):
                cache.open_session(1, 2 * PAGE, timeout=0)  # fail fast, like the reference's alloc_timeout=0
        assert cache.tokens_left == 5 * PAGE
    assert cache.tokens_left == 8 * PAGE
    with pytest.raises(AllocationFailed):
        cache.open_session(1, 100 * PAGE, timeout=0)  # can never fit


def test_timeout_and_queueing():
    cache = MemoryCache(max_size_tokens=4 * PAGE, max_alloc_timeout=5.0, device="cpu")
    big = cache.open_session(1, 3 * PAGE, timeout=0)  # takes all 4 pages
    t0 = time.perf_counter()
    with pytest.raises(AllocationFailed):
        cache.open_session(1, PAGE, timeout=0.3)
    assert 0.25 < time.perf_counter() - t0 < 2.0
    order = []

    def waiter(name, delay):
        time.sleep(delay)
        s = cache.open_session(1, PAGE, timeout=4.0)
        order.append(name)
        time.sleep(0.05)
        s.close()

    threads = [threading.Thread(target=waiter, args=("first", 0.0)), threading.Thread(target=waiter, args=("second", 0.15))]
    for t in threads:
        t.start()
    time.sleep(0.4)
    assert order == []  # both are queued behind the big session
    big.close()
    for t in threads:
        t.join(timeout=5)
    assert order == ["first", "second"]  # FIFO among waiters
    assert cache.tokens_left == 4 * PAGE


def test_dense_session_semantics():
    from petals_b200.models.spec import BlockSpec

    spec = BlockSpec(family="llama", hidden_size=32, num_heads=4, num_kv_heads=2, head_dim=8, intermediate_size=64)
    cache = MemoryCache(max_size_tokens=16 * PAGE, device="cpu", spec=spec, dtype=torch.float32)
    with cache.allocate_cache(3, 10, timeout=0) as sess:
        k, v = sess.dense_kv(0, spec, torch.float32, "cpu")
        assert k.shape == (3, 10, 2, 8)
        k[:] = torch.arange(3.0).view(3, 1, 1, 1)
        sess.prepare_write(4)
        sess.set_position(4)
        with pytest.raises(ValueError, match="Maximum length exceeded"):
            sess.prepare_write(7)
        sess.reorder(torch.tensor([2, 0, 0]))
        k2, _ = sess.dense_kv(0, spec, torch.float32, "cpu")
        assert k2[:, 0, 0, 0].tolist() == [2.0, 0.0, 0.0]
        sess.set_position(2)  # rollback
        assert sess.position == 2
        with pytest.raises(ValueError):
            sess.set_position(11)


def test_native_allocator_and_queue_tolerate_misuse():
    """The C ABI of the host runtime (csrc/runtime) is called through ctypes: nothing a caller passes may corrupt memory or let a C++
    exception escape (found by fuzzing: out-of-range incref / refcount, negative sizes, NaN priorities)."""
    import ctypes as C

    from petals_b200.ops import native

    rt = native.rt()
    assert not rt.pb_kv_create(-3)  # no std::length_error through the ABI
    h = rt.pb_kv_create(8)
    pages = (C.c_int * 16)()
    assert rt.pb_kv_alloc(h, 3, pages) == 0 and rt.pb_kv_num_free(h) == 5
    assert rt.pb_kv_alloc(h, 100, pages) == -1 and rt.pb_kv_num_free(h) == 5  # all or nothing
    wild = (C.c_int * 4)(-5, 99, 1 << 30, 7)  # 7 is a valid but free page: it cannot gain an owner either
    rt.pb_kv_incref(h, wild, 4)
    rt.pb_kv_free(h, wild, 4)
    assert rt.pb_kv_num_free(h) == 5 and rt.pb_kv_refcount(h, 7) == 0
    assert rt.pb_kv_refcount(h, 1000) == -1 and rt.pb_kv_refcount(h, -1) == -1
    live = (C.c_int * 3)(0, 1, 2)
    rt.pb_kv_free(h, live, 3)
    rt.pb_kv_free(h, live, 3)  # double free is ignored, pages are not duplicated in the free list
    assert rt.pb_kv_num_free(h) == 8
    assert rt.pb_kv_alloc(h, 8, pages) == 0 and sorted(pages[:8]) == list(range(8))
    assert rt.pb_kv_reserve(h, -4, 0.01) == -1 and rt.pb_kv_reserve(h, 1000, 0.01) == -1 and rt.pb_kv_reserved(h) == 0
    rt.pb_kv_unreserve(h, 50)
    rt.pb_kv_unreserve(h, -50)
    assert rt.pb_kv_reserved(h) == 0
    rt.pb_kv_destroy(h)

    q = rt.pb_tq_create()
    tid, prio = C.c_int64(), C.c_double()
    for priority, task in ((float("nan"), 1), (1.0, 2), (float("-inf"), 3), (1.0, 4)):
        rt.pb_tq_push(q, priority, task)
    order = []
    while rt.pb_tq_pop(q, 0.0, C.byref(tid), C.byref(prio)) == 0:
        order.append(tid.value)
    assert order == [3, 2, 4, 1]  # -inf first, FIFO among equals, the NaN task last instead of scrambling the heap
    rt.pb_tq_close(q)
    assert rt.pb_tq_pop(q, 0.0, C.byref(tid), C.byref(prio)) == -2
    rt.pb_tq_destroy(q)


def test_paged_sessions_survive_random_beam_reorders_and_rollbacks():
    """The page-table bookkeeping of the GPU path (bind pages, copy-on-write of pages shared between hypotheses, reorder, rollback), run
    against a host pool: every sequence must always read back exactly the history it is supposed to have. Tokens are written into the
    pool the way the kernels address it — page = table[pos // PAGE], slot = pos % PAGE."""
    import random

    from petals_b200.models.spec import BlockSpec
    from petals_b200.ops.functional import PAGE

    spec = BlockSpec(family="llama", hidden_size=8, num_heads=1, num_kv_heads=1, head_dim=2, intermediate_size=8)
    rng = random.Random(0)
    for trial in range(25):
        cache = MemoryCache(64 * PAGE, None, n_blocks=1, spec=spec, dtype=torch.float32, device="cpu", paged=True, max_length=6 * PAGE)
        B = rng.choice([1, 2, 3, 4])
        session = cache.open_session(B, 6 * PAGE, timeout=0)
        histories = [[] for _ in range(B)]  # what each sequence should contain
        stamp = 0

        def read(b):
            table = session.tables[b]
            return [int(cache.pool[0, 0, table[p // PAGE], 0, p % PAGE, 0]) for p in range(session.position)]

        for _ in range(40):
            op = rng.choice(["write", "write", "write", "reorder", "rollback"])
            if op == "write":
                n = rng.randint(1, PAGE + 5)
                if session.position + n > session.max_length:
                    continue
                session.prepare_write(n)
                for b in range(B):
                    for p in range(session.position, session.position + n):
                        stamp += 1
                        cache.pool[0, :, session.tables[b][p // PAGE], 0, p % PAGE, :] = stamp
                        histories[b].append(stamp)
                session.set_position(session.position + n)
            elif op == "reorder":
                ids = [rng.randrange(B) for _ in range(B)]
                session.reorder(torch.tensor(ids))
                histories = [list(histories[i]) for i in ids]
            else:
                back = rng.randint(0, session.position)
                session.set_position(back)
                histories = [h[:back] for h in histories]
            for b in range(B):
                assert read(b) == histories[b], (trial, op, b)
            # no page is referenced by more sequences than its reference count says, none is both free and in use
            in_use = [p for t in session.tables for p in t]
            for p in set(in_use):
                assert cache._rt.pb_kv_refcount(cache._alloc, p) == in_use.count(p)
        used = len({p for t in session.tables for p in t})
        assert cache._rt.pb_kv_num_free(cache._alloc) == cache.num_pages - used
        session.close()
        assert cache._rt.pb_kv_num_free(cache._alloc) == cache.num_pages and cache.tokens_left == cache.num_pages * PAGE


def test_fail_fast_reservation_does_not_queue_behind_a_waiter():
    """A reservation with timeout 0 (the client's default alloc_timeout) returns at once even when another reservation is
    blocked ahead of it, and the queue keeps moving after out-of-order give-ups (reference: memory_cache.py:103-131)."""
    import threading
    import time

    from petals_b200.ops import native

    rt = native.rt()
    h = rt.pb_kv_create(8)
    assert rt.pb_kv_reserve(h, 8, 0.0) == 0  # the pool is fully promised
    results = {}

    def waiter():
        t0 = time.perf_counter()
        results["code"] = rt.pb_kv_reserve(h, 4, 1.5)
        results["waited"] = time.perf_counter() - t0

    th = threading.Thread(target=waiter)
    th.start()
    time.sleep(0.2)  # the waiter is now queued at the head
    t0 = time.perf_counter()
    assert rt.pb_kv_reserve(h, 1, 0.0) == -1
    assert rt.pb_kv_reserve(h, 1, 0.05) == -1
    assert time.perf_counter() - t0 < 0.5, "fail-fast reservations must not wait for the queue head's timeout"
    rt.pb_kv_unreserve(h, 8)
    th.join()
    assert results["code"] == 0 and results["waited"] < 1.4  # the head got its pages as soon as they were released
    assert rt.pb_kv_reserve(h, 4, 0.0) == 0  # abandoned tickets were skipped: the queue is not stuck
    assert rt.pb_kv_reserved(h) == 8
    rt.pb_kv_destroy(h)

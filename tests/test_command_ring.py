"""The shared-memory command ring of a tensor-parallel worker group (parallel/control.py): one producer, several consumer processes,
many wrap-arounds of a deliberately tiny ring, back-pressure and size limits. (The NVLink side of a TP group needs GPUs; its control
plane does not.)"""
import multiprocessing as mp
import time

import pytest

from petals_b200.parallel.control import CommandRing


def _consume(name: str, consumer: int, n: int, slow_every: int, out):
    ring = CommandRing(name, create=False)
    got = []
    for i in range(n):
        got.append(ring.recv(consumer, timeout=60))
        if slow_every and i % slow_every == 0:
            time.sleep(0.002)  # a follower that lags: the leader must wait instead of overwriting unread slots
    out.put((consumer, got))
    ring.close()


def test_every_consumer_sees_every_command_in_order_across_wraparounds():
    n, consumers = 1500, 3
    ring = CommandRing(None, create=True, n_consumers=consumers, slots=8, slot_bytes=256)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_consume, args=(ring.name, c, n, (0, 7, 50)[c], out)) for c in range(consumers)]
    try:
        for p in procs:
            p.start()
        sent = [{"op": "step", "sid": i % 5, "B": 1 + i % 3, "T": 1, "pos": i, "hypo": list(range(i % 4))} if i % 10 else {"op": "open", "sid": i, "note": "x" * (i % 100)}
                for i in range(n)]
        for cmd in sent:
            ring.send(cmd, timeout=60)
        results = dict(out.get(timeout=120) for _ in range(consumers))
        for p in procs:
            p.join(timeout=30)
        assert all(p.exitcode == 0 for p in procs)
        for c in range(consumers):
            assert results[c] == sent, f"consumer {c} saw a different command stream"
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
        ring.close()


def test_back_pressure_and_limits():
    ring = CommandRing(None, create=True, n_consumers=1, slots=4, slot_bytes=64)
    try:
        for i in range(4):
            ring.send({"i": i})
        t0 = time.monotonic()
        with pytest.raises(TimeoutError, match="full"):  # nobody consumes: the fifth command may not overwrite the first
            ring.send({"i": 4}, timeout=0.2)
        assert 0.15 < time.monotonic() - t0 < 5
        assert ring.recv(0, timeout=1) == {"i": 0}
        ring.send({"i": 4}, timeout=1)  # room again
        assert [ring.recv(0, timeout=1)["i"] for _ in range(4)] == [1, 2, 3, 4]
        with pytest.raises(TimeoutError, match="no command"):
            ring.recv(0, timeout=0.05)
        with pytest.raises(ValueError, match="slot size"):
            ring.send({"blob": "x" * 200})
        # a second handle on the same memory sees the same geometry (what a follower does)
        other = CommandRing(ring.name, create=False)
        assert (other.n_consumers, other.slots, other.slot_bytes) == (1, 4, 64)
        ring.send({"late": True})
        assert other.recv(0, timeout=1) == {"late": True}
        other.close()
    finally:
        ring.close()

"""Cancellation-safe waiting for user code that drives the client from asyncio (reference: src/petals/utils/asyncio.py:4-21,
used there so that a cancelled handler does not leak the KV allocation lock).  The engine itself is thread-based."""
import asyncio
from typing import Any, Awaitable


async def shield_and_wait(task: Awaitable[Any]) -> Any:
    """Run ``task`` to completion no matter how often the caller is cancelled meanwhile; then deliver the cancellation
    (if there was one) or the task's result / exception."""
    inner = asyncio.ensure_future(task)
    loop = asyncio.get_running_loop()
    was_cancelled = False
    while not inner.done():
        wake = loop.create_future()
        inner.add_done_callback(lambda _done, wake=wake: wake.cancelled() or wake.done() or wake.set_result(None))
        try:
            await wake
        except asyncio.CancelledError:
            was_cancelled = True  # remember it; the critical section is not finished yet
    if was_cancelled:
        inner.exception() if not inner.cancelled() else None  # mark a possible exception as retrieved
        raise asyncio.CancelledError()
    return inner.result()

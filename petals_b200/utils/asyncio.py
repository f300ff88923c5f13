"""Cancellation-safe waiting (reference: src/petals/utils/asyncio.py:4-21). The engine is thread-based, but the
helper is kept for user code that drives the client from asyncio."""
import asyncio


async def shield_and_wait(task):
    """Await ``task`` to completion even if the awaiting coroutine is cancelled; re-raise the cancellation after."""
    if not isinstance(task, asyncio.Task):
        task = asyncio.create_task(task)
    cancel_exc = None
    while True:
        try:
            result = await asyncio.shield(task)
            break
        except asyncio.CancelledError as e:
            cancel_exc = e
    if cancel_exc is not None:
        raise cancel_exc
    return result

"""Which queued task runs first on a stage (reference: src/petals/server/task_prioritizer.py:1-20).

A priority is a float, smaller = sooner; ties are broken by arrival time (server/task_pool.py).  The default policy keeps the
reference's two classes — a single-stream inference step (latency-critical, a user is waiting for a token) outranks
forward / backward passes (throughput work) — and lets points paid by the client pull a task forward inside its class.
"""
from typing import Any

PRIORITY_INFERENCE = 1.0
PRIORITY_TRAINING = 2.0  # rpc_forward / rpc_backward


class TaskPrioritizerBase:
    """Subclass and override :meth:`prioritize` to change the queueing discipline of a server."""

    def prioritize(self, *tensors: Any, points: float = 0.0, **kwargs: Any) -> float:
        raise NotImplementedError


class DummyTaskPrioritizer(TaskPrioritizerBase):
    """``type="inference"`` -> 1.0, everything else -> 2.0 (the reference's constants; points do not reorder anything)."""

    def prioritize(self, *tensors: Any, points: float = 0.0, **kwargs: Any) -> float:
        return PRIORITY_INFERENCE if kwargs.get("type") == "inference" else PRIORITY_TRAINING


class PointsTaskPrioritizer(DummyTaskPrioritizer):
    """Same two classes, but within a class a request that carries points moves ahead: priority = class - points / (1 + points) / 2
    (bounded, so paid training work never overtakes unpaid inference)."""

    def prioritize(self, *tensors: Any, points: float = 0.0, **kwargs: Any) -> float:
        base = super().prioritize(*tensors, points=points, **kwargs)
        paid = max(float(points), 0.0)
        return base - 0.5 * paid / (1.0 + paid)

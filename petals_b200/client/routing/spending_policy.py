"""Request "points": what a client is willing to spend on a call, used by servers to order their queues
(reference: src/petals/client/routing/spending_policy.py:1-17 — a placeholder there as well: everything costs 0 points, so
ordering falls back to task type and arrival time, server/task_prioritizer.py)."""
from typing import Any


class SpendingPolicyBase:
    """Decides how many points accompany an RPC (``protocol`` is ``"rpc_inference"``, ``"rpc_forward"`` or ``"rpc_backward"``)."""

    def get_points(self, protocol: str, *args: Any, **kwargs: Any) -> float:
        raise NotImplementedError(f"{type(self).__name__} does not define get_points")


class ConstantSpendingPolicy(SpendingPolicyBase):
    """The same number of points for every call."""

    def __init__(self, points: float = 0.0):
        if points < 0:
            raise ValueError("points must be non-negative")
        self.points = float(points)

    def get_points(self, protocol: str, *args: Any, **kwargs: Any) -> float:
        return self.points


class NoSpendingPolicy(ConstantSpendingPolicy):
    """Spend nothing: the default."""

    def __init__(self):
        super().__init__(0.0)

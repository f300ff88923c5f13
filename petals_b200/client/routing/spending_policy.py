"""Request "points" for server-side prioritisation (reference: src/petals/client/routing/spending_policy.py:1-17).
A placeholder policy: every request carries 0 points, so ordering is decided by task type and arrival time."""
from abc import ABC, abstractmethod


class SpendingPolicyBase(ABC):
    @abstractmethod
    def get_points(self, protocol: str, *args, **kwargs) -> float:
        pass


class NoSpendingPolicy(SpendingPolicyBase):
    def get_points(self, protocol: str, *args, **kwargs) -> float:
        return 0.0

"""Task priorities (reference: src/petals/server/task_prioritizer.py:1-20): lower runs first; latency-critical
single-stream inference pre-empts throughput work (forward / backward)."""
from abc import ABC, abstractmethod


class TaskPrioritizerBase(ABC):
    @abstractmethod
    def prioritize(self, *input, points: float, **kwargs) -> float:
        """Priority of a task; smaller = sooner."""


class DummyTaskPrioritizer(TaskPrioritizerBase):
    def prioritize(self, *input, points: float, **kwargs) -> float:
        if kwargs.get("type") == "inference":
            return 1.0
        return 2.0

"""Sampling helper (reference: src/petals/utils/random.py): at most ``k`` distinct elements, all of them when there are fewer."""
import random
from typing import Iterable, List, Optional, TypeVar

T = TypeVar("T")


def sample_up_to(population: Iterable[T], k: int, rng: Optional[random.Random] = None) -> List[T]:
    """Uniform sample without replacement that keeps the population's own order (routing logs stay readable)."""
    items = list(population)
    if k >= len(items):
        return items
    chosen = set((rng or random).sample(range(len(items)), max(k, 0)))
    return [item for i, item in enumerate(items) if i in chosen]

"""Deprecated alias kept for import compatibility (reference: src/petals/dht_utils.py:1-9)."""
import warnings

warnings.warn("petals_b200.dht_utils has moved to petals_b200.utils.dht", DeprecationWarning, stacklevel=2)
from petals_b200.utils.dht import *  # noqa: F401,F403,E402

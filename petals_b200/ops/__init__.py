"""Hand-written sm_100a ops and their PyTorch oracles."""
from petals_b200.ops import functional, native  # noqa: F401

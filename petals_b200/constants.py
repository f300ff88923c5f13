"""Shared constants (reference: src/petals/constants.py:1-18).

The public-swarm bootstrap peers of the reference have no meaning inside a single NVLink box; the
"swarm" is addressed by a rendezvous directory (or lives in-process), see petals_b200.parallel.swarm.
"""
import torch

PUBLIC_INITIAL_PEERS: list = []  # no Internet swarm: rank -> block-span map is static
REACHABILITY_API_URL = None

DTYPE_MAP = dict(bfloat16=torch.bfloat16, float16=torch.float16, float32=torch.float32, auto="auto")

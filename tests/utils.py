"""Shared helpers for the swarm tests: tiny synthetic checkpoints + in-process multi-stage swarms."""
from __future__ import annotations

import contextlib
import itertools
import os
import tempfile
from typing import Dict, List, Optional, Sequence

import torch

from petals_b200.parallel.swarm import Swarm
from petals_b200.server.server import Server
from petals_b200.utils.checkpoints import make_random_checkpoint

_counter = itertools.count()
_ckpt_cache: Dict[tuple, str] = {}


def checkpoint(model_type: str = "llama", **overrides) -> str:
    key = (model_type, tuple(sorted(overrides.items())))
    if key not in _ckpt_cache:
        d = tempfile.mkdtemp(prefix=f"pb200-{model_type}-")
        _ckpt_cache[key] = make_random_checkpoint(d, model_type, **overrides)
    return _ckpt_cache[key]


@contextlib.contextmanager
def swarm_of(path: str, spans: Sequence[str], *, swarm: Optional[Swarm] = None, device: str = "cpu", torch_dtype: str = "float32", **kwargs):
    """Start one Server per span ("a:b") on `device`; yields (swarm, servers)."""
    swarm = swarm or Swarm(f"test-{next(_counter)}")
    servers: List[Server] = []
    try:
        for span in spans:
            s = Server(initial_peers=swarm, converted_model_name_or_path=path, block_indices=span, torch_dtype=torch_dtype, device=device,
                       throughput=1.0, update_period=0.5, mean_balance_check_period=1000, **kwargs)
            s.run_in_background(timeout=120)
            servers.append(s)
        yield swarm, servers
    finally:
        for s in servers:
            s.shutdown()


def local_blocks(path: str, n: int, dtype=torch.float32):
    from petals_b200.server.from_pretrained import load_pretrained_block

    return [load_pretrained_block(path, i, torch_dtype=dtype) for i in range(n)]

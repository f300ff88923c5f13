"""KV-cache accounting visible through rpc_info while sessions open and close (reference: tests/test_server_stats.py).

Accounting here is page-granular: a session of ``max_length`` tokens reserves ``pages_needed(batch, max_length)`` pages of
``PAGE`` tokens in every block of the stage (memory_cache.py), instead of the reference's exact byte count."""
import time

import torch

from petals_b200.client.remote_sequential import RemoteSequential
from petals_b200.ops.functional import PAGE
from petals_b200.server.handler import CACHE_TOKENS_AVAILABLE
from petals_b200.server.memory_cache import MemoryCache
from petals_b200.utils.auto_config import AutoDistributedConfig
from tests.utils import checkpoint, swarm_of


def test_server_info(block_from: int = 1, block_to: int = 4, max_length: int = 100, max_length2: int = 50):
    path = checkpoint("llama")
    with swarm_of(path, ["0:4"], attn_cache_tokens=4096, inference_max_length=512) as (swarm, servers):
        config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm)
        config.allowed_servers = [servers[0].peer_id]
        n_stage_blocks = 4
        blocks1 = RemoteSequential(config, dht=swarm, start_block=block_from, end_block=block_to)
        blocks2 = RemoteSequential(config, dht=swarm, start_block=block_to - 1, end_block=block_to)

        info_before = blocks1.sequence_manager.rpc_info
        with blocks1.inference_session(max_length=max_length) as sess:
            sess.step(torch.randn(1, 1, config.hidden_size))
            blocks1.sequence_manager.state.rpc_info = None  # invalidate the cached copy
            info_inside = blocks1.sequence_manager.rpc_info
            with blocks2.inference_session(max_length=max_length2) as sess2:
                sess2.step(torch.randn(1, 1, config.hidden_size))
                blocks2.sequence_manager.state.rpc_info = None
                info_inside2 = blocks2.sequence_manager.rpc_info
        time.sleep(0.1)
        blocks1.sequence_manager.state.rpc_info = None
        info_after = blocks1.sequence_manager.rpc_info

        assert info_before[CACHE_TOKENS_AVAILABLE] == info_after[CACHE_TOKENS_AVAILABLE]
        reserved1 = MemoryCache.pages_needed(1, max_length) * PAGE * n_stage_blocks
        reserved2 = MemoryCache.pages_needed(1, max_length2) * PAGE * n_stage_blocks
        assert info_before[CACHE_TOKENS_AVAILABLE] - info_inside[CACHE_TOKENS_AVAILABLE] == reserved1
        assert info_inside[CACHE_TOKENS_AVAILABLE] - info_inside2[CACHE_TOKENS_AVAILABLE] == reserved2

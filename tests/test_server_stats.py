"""KV budget as seen through ``rpc_info`` while sessions come and go (reference: tests/test_server_stats.py:12-39).

Here the budget is counted in pages: a session of ``max_length`` tokens reserves ``pages_needed(batch, max_length)`` pages of
``PAGE`` tokens in *every* block of the stage that serves it (server/memory_cache.py), whichever sub-span the client asked for."""
import torch

from petals_b200.client.remote_sequential import RemoteSequential
from petals_b200.ops.functional import PAGE
from petals_b200.server.handler import CACHE_TOKENS_AVAILABLE
from petals_b200.server.memory_cache import MemoryCache
from petals_b200.utils.auto_config import AutoDistributedConfig
from tests.utils import checkpoint, swarm_of

STAGE_BLOCKS = 4


def _tokens_available(seq: RemoteSequential) -> int:
    seq.sequence_manager.state.rpc_info = None  # drop the client's cached copy: ask the server again
    return seq.sequence_manager.rpc_info[CACHE_TOKENS_AVAILABLE]


def _reserved(max_length: int, batch: int = 1) -> int:
    return MemoryCache.pages_needed(batch, max_length) * PAGE * STAGE_BLOCKS


def test_cache_tokens_available_tracks_open_sessions():
    path = checkpoint("llama")
    with swarm_of(path, [f"0:{STAGE_BLOCKS}"], attn_cache_tokens=4096, inference_max_length=512) as (swarm, servers):
        config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm)
        config.allowed_servers = [servers[0].peer_id]
        wide = RemoteSequential(config, dht=swarm, start_block=1, end_block=4)
        narrow = RemoteSequential(config, dht=swarm, start_block=3, end_block=4)
        token = torch.randn(1, 1, config.hidden_size)

        idle = _tokens_available(wide)
        assert idle > 0
        with wide.inference_session(max_length=100) as first:
            first.step(token)
            one_open = _tokens_available(wide)
            with narrow.inference_session(max_length=50) as second:
                second.step(token)
                two_open = _tokens_available(narrow)
            after_inner_close = _tokens_available(wide)
        after_all_closed = _tokens_available(wide)

        assert idle - one_open == _reserved(100)
        assert one_open - two_open == _reserved(50)
        assert after_inner_close == one_open and after_all_closed == idle  # everything is returned, nothing leaks

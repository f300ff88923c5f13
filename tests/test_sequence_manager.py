"""Routing (reference tests/test_sequence_manager.py, test_server_stats.py): both modes, spans via different peers,
allow/block lists, bans and retries, rpc_info cache accounting, background thread shutdown."""
import time

import pytest
import torch

from petals_b200.client.remote_sequential import RemoteSequential
from petals_b200.client.routing import MissingBlocksError, RemoteSequenceManager
from petals_b200.data_structures import make_uid
from petals_b200.utils.auto_config import AutoDistributedConfig
from tests.utils import checkpoint, swarm_of


@pytest.fixture(scope="module")
def served():
    path = checkpoint("llama")
    with swarm_of(path, ["0:2", "2:4", "0:4"]) as (swarm, servers):
        yield path, swarm, servers


@pytest.mark.parametrize("mode", ["min_latency", "max_throughput"])
def test_make_sequence(served, mode):
    path, swarm, servers = served
    config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm)
    uids = [make_uid(config.dht_prefix, i) for i in range(config.num_hidden_layers)]
    manager = RemoteSequenceManager(config, uids, dht=swarm)
    for _ in range(5):
        chain = manager.make_sequence(0, 4, mode=mode, cache_tokens_needed=64 if mode == "min_latency" else None)
        assert chain[0].start == 0 and chain[-1].end == 4
        assert all(a.end == b.start for a, b in zip(chain[:-1], chain[1:]))
        assert all(a.peer_id != b.peer_id for a, b in zip(chain[:-1], chain[1:]))  # consecutive spans use different peers
    sub = manager[1:3]
    assert len(sub) == 2 and sub.block_uids == tuple(uids[1:3])
    assert sub.make_sequence(mode=mode)[0].start == 0
    assert manager.is_alive
    manager.shutdown()
    assert not manager.is_alive


def test_allow_block_lists_and_missing_blocks(served):
    path, swarm, servers = served
    only_full = servers[2].peer_id
    config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm, allowed_servers=[only_full])
    seq = RemoteSequential(config, dht=swarm)
    chain = seq.sequence_manager.make_sequence(mode="min_latency")
    assert [s.peer_id for s in chain] == [only_full]
    config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm, blocked_servers=[only_full])
    chain = RemoteSequential(config, dht=swarm).sequence_manager.make_sequence(mode="max_throughput")
    assert only_full not in [s.peer_id for s in chain] and len(chain) == 2
    config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm, allowed_servers=["nobody"], max_retries=1)
    with pytest.raises(MissingBlocksError):
        RemoteSequential(config, dht=swarm).sequence_manager.make_sequence(mode="min_latency")


def test_retry_delay_and_ban(served):
    path, swarm, servers = served
    config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm, min_backoff=1, max_backoff=8, ban_timeout=0.2)
    manager = RemoteSequential(config, dht=swarm).sequence_manager
    assert [manager.get_retry_delay(i) for i in range(6)] == [0, 1, 2, 4, 8, 8]
    manager.make_sequence(mode="min_latency")
    victim = servers[2].peer_id
    manager.on_request_failure(victim)
    assert victim in manager.state.banned_peers
    chain = manager.make_sequence(mode="min_latency")
    assert victim not in [s.peer_id for s in chain]
    time.sleep(0.25)
    assert victim not in manager.state.banned_peers  # ban expired
    manager.on_request_success(victim)


def test_server_stats_cache_accounting(served):
    path, swarm, servers = served
    config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm, allowed_servers=[servers[2].peer_id])
    seq = RemoteSequential(config, dht=swarm)
    info = seq.sequence_manager.rpc_info
    assert {"version", "cache_tokens_available", "inference_schema", "forward_schema", "outputs_schema", "keyword_names"} <= set(info)
    handler = servers[2].module_container.handler
    before = handler.rpc_info()["cache_tokens_available"]
    with seq.inference_session(max_length=128) as sess:
        sess.step(torch.randn(1, 3, config.hidden_size))
        during = handler.rpc_info()["cache_tokens_available"]
        assert during <= before - 128 * config.num_hidden_layers  # max_length tokens x n_blocks are reserved up front
    assert handler.rpc_info()["cache_tokens_available"] == before  # and released when the session closes


def test_refresh_thread_ends_with_its_manager():
    """The background refresh only holds a weak reference: dropping the model/manager ends the thread (the reference's
    managers poll the DHT until the process exits unless shutdown() is called, src/petals/client/routing/sequence_manager.py:493-519)."""
    import gc
    import threading
    import time

    from petals_b200.client.config import ClientConfig
    from petals_b200.client.routing.sequence_manager import RemoteSequenceManager
    from petals_b200.parallel.swarm import Swarm

    swarm = Swarm("empty-swarm-for-thread-test")
    config = ClientConfig(initial_peers=[swarm.address], dht_prefix="nobody", update_period=0.05, max_retries=0)
    manager = RemoteSequenceManager(config, ["nobody.0", "nobody.1"], dht=swarm)
    manager._ensure_thread()  # no server holds these blocks: the loop keeps failing quietly and retrying
    thread = manager._thread
    time.sleep(0.3)
    assert thread.is_alive() and not manager.ready.is_set()
    del manager
    gc.collect()
    thread.join(timeout=5)
    assert not thread.is_alive()
    assert not any(t.name == "sequence-manager" and t is thread for t in threading.enumerate())

"""A live stage leaves a doubly-served span for an unserved one (the reference CI's server1: ``--num_blocks 5
--mean_balance_check_period 10`` "must move off 0:5", .github/workflows/run-tests.yaml:67-69; policy in
src/petals/server/block_selection.py:40-95, loop in src/petals/server/server.py:328-384)."""
import time

import torch

from petals_b200.parallel.swarm import Swarm
from petals_b200.server.server import Server
from petals_b200.utils.auto_config import AutoDistributedConfig, AutoDistributedModelForCausalLM
from tests.utils import checkpoint, local_blocks


def _span(server):
    c = server.module_container
    return None if c is None or not c.ready.is_set() else (c.stage.start_block, c.stage.end_block)


def test_free_server_rebalances_onto_the_unserved_span():
    path = checkpoint("llama")
    n = AutoDistributedConfig.from_pretrained(path).num_hidden_layers
    assert n == 4
    swarm = Swarm("rebalance-test")
    common = dict(initial_peers=swarm, converted_model_name_or_path=path, torch_dtype="float32", device="cpu", throughput=1.0, update_period=0.3)
    free = Server(num_blocks=2, mean_balance_check_period=0.3, mean_block_selection_delay=0.0, peer_id="free", **common)
    pinned = None
    try:
        free.run_in_background(timeout=120)
        assert _span(free) == (0, 2)  # an empty swarm: every window is equally bad, the first one wins
        pinned = Server(block_indices="0:2", mean_balance_check_period=1000, peer_id="pinned", **common)
        pinned.run_in_background(timeout=120)
        deadline = time.monotonic() + 90
        while _span(free) != (2, 4):  # blocks 2..3 have no server: the bottleneck throughput is 0 until `free` moves
            assert time.monotonic() < deadline, f"still serving {_span(free)}"
            time.sleep(0.1)
        assert _span(pinned) == (0, 2)  # --block_indices pins a stage, it never moves
        # the swarm is whole now and stays that way (no ping-pong back to 0:2)
        model = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=swarm, max_retries=20, min_backoff=0.2, max_backoff=0.5)
        ids = torch.randint(0, model.config.vocab_size, (1, 5), generator=torch.Generator().manual_seed(0))
        with torch.inference_mode():
            logits = model(ids).logits
            h = model.model.embed(ids)
            for b in local_blocks(path, n):
                h = b(h)[0]
            assert torch.allclose(logits, model.lm_head(model.model.final_norm(h)), atol=1e-3)
        time.sleep(1.5)
        assert _span(free) == (2, 4)
    finally:
        free.shutdown()
        if pinned is not None:
            pinned.shutdown()

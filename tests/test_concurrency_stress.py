"""Many clients hammering one two-stage swarm at once (the reference relies on locks by construction and has no
stress test). Sessions of random shapes open, step, roll back and close concurrently with forward/backward traffic while the KV
budget is tight enough that allocations have to wait for each other; every result is checked against the local blocks and the
budget must be fully returned at the end."""
import random
import threading

import torch

from petals_b200.client.remote_sequential import RemoteSequential
from petals_b200.server.handler import CACHE_TOKENS_AVAILABLE
from petals_b200.utils.auto_config import AutoDistributedConfig
from tests.utils import checkpoint, local_blocks, swarm_of

N_THREADS, ROUNDS = 10, 6


def test_concurrent_sessions_forward_backward_and_rollbacks():
    path = checkpoint("llama")
    # 512 tokens of KV per block: at most a handful of the sessions below fit at the same time -> allocations queue (alloc_timeout)
    with swarm_of(path, ["0:2", "2:4"], attn_cache_tokens=512, inference_max_length=256, max_alloc_timeout=60) as (swarm, servers):
        config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm)
        blocks = local_blocks(path, config.num_hidden_layers)
        lock, errors = threading.Lock(), []
        idle_budget = [s.module_container.handler.rpc_info()[CACHE_TOKENS_AVAILABLE] for s in servers]

        def local(x):
            with lock, torch.no_grad():  # the oracle blocks keep scratch state (rope cache): serialise the reference computation
                h = x
                for b in blocks:
                    h = b(h)[0]
                return h

        def worker(seed: int):
            rng = random.Random(seed)
            gen = torch.Generator().manual_seed(seed)
            seq = RemoteSequential(config, dht=swarm)
            try:
                for _ in range(ROUNDS):
                    kind = rng.choice(["session", "session", "forward", "backward"])
                    B, T = rng.choice([1, 2]), rng.randint(2, 9)
                    x = torch.randn(B, T, config.hidden_size, generator=gen)
                    if kind == "session":
                        cut = rng.randint(1, T - 1)
                        with torch.inference_mode(), seq.inference_session(max_length=rng.choice([16, 64, 128]), alloc_timeout=30) as sess:
                            out = [sess.step(x[:, :cut])]
                            if rng.random() < 0.5:  # speculative-style rollback: redo the last positions
                                back = rng.randint(0, cut - 1)
                                sess.position = back
                                out = [out[0][:, :back], sess.step(x[:, back:cut])]
                            out.append(sess.step(x[:, cut:]))
                            got = torch.cat(out, dim=1)
                        assert torch.allclose(got, local(x), atol=1e-4), "session output differs"
                    elif kind == "forward":
                        with torch.no_grad():
                            assert torch.allclose(seq(x), local(x), atol=1e-4), "forward differs"
                    else:
                        xr = x.clone().requires_grad_(True)
                        seq(xr).pow(2).sum().backward()
                        assert torch.isfinite(xr.grad).all() and xr.grad.abs().sum() > 0
            except Exception as e:  # noqa: BLE001 - collected and reported by the main thread
                errors.append(f"worker {seed}: {type(e).__name__}: {e}")

        threads = [threading.Thread(target=worker, args=(i,)) for i in range(N_THREADS)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=240)
        assert not any(t.is_alive() for t in threads), "a worker is stuck"
        assert not errors, errors[:3]
        after = [s.module_container.handler.rpc_info()[CACHE_TOKENS_AVAILABLE] for s in servers]
        assert after == idle_budget, "KV pages leaked"
        for s in servers:
            snap = s.module_container.handler.metrics.snapshot()
            assert snap["sessions_active"] == 0 and snap["errors"] == {"inference": 0, "forward": 0, "backward": 0}

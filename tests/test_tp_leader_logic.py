"""Host-side logic of the tensor-parallel leader (command stream to the followers, prompt chunking, scratch sessions for
rpc_forward) against a fake engine and a fake ring on CPU. The NVLink engine itself is covered by tests/test_multi_gpu.py."""
import torch

from petals_b200.models.spec import BlockSpec
from petals_b200.parallel.tensor_parallel import MAX_ROWS
from petals_b200.parallel.tp_worker import TPLeaderEngine
from petals_b200.server.memory_cache import MemoryCache
from petals_b200.utils.auto_config import AutoDistributedConfig
from tests.utils import checkpoint


class _FakeRing:
    def __init__(self):
        self.sent = []

    def send(self, cmd):
        self.sent.append(dict(cmd))


class _FakeEngine:
    """Records what the leader asks for and returns x + 1 so the data path can be checked."""

    def __init__(self, cache, hidden, max_prefill_rows):
        self.cache, self.n_blocks, self.max_prefill_rows, self.hidden = cache, 2, max_prefill_rows, hidden
        self.calls, self._staged = [], None

    def push_inputs(self, rows):
        self._staged = rows.clone()

    def push_prefill_inputs(self, rows):
        self._staged = rows.clone()

    def _run(self, kind, session, B, T):
        assert self._staged.shape == (B * T, self.hidden)
        self.calls.append((kind, B, T, session.position))
        session.prepare_write(T)
        session.set_position(session.position + T)
        return self._staged + 1

    def run_step(self, session, B, T):
        assert B * T <= MAX_ROWS
        return self._run("step", session, B, T)

    def run_prefill(self, session, B, T):
        assert B * T <= self.max_prefill_rows
        return self._run("prefill", session, B, T)

    def check_errors(self):
        pass


def _setup(max_prefill_rows=16):
    config = AutoDistributedConfig.from_pretrained(checkpoint("llama"))
    spec: BlockSpec = config.block_spec()
    cache = MemoryCache(4096, None, n_blocks=2, spec=spec, dtype=torch.float32, device=torch.device("cpu"), paged=False, max_length=512)
    engine, ring = _FakeEngine(cache, spec.hidden_size, max_prefill_rows), _FakeRing()
    return TPLeaderEngine(engine, ring), engine, ring, cache, spec.hidden_size


def test_prompt_is_chunked_into_prefill_commands_and_decode_steps():
    leader, engine, ring, cache, H = _setup(max_prefill_rows=16)
    sess = cache.open_session(1, 128, timeout=0)
    x = torch.randn(1, 40, H)
    y = leader.inference_step(sess, x)
    assert torch.allclose(y.float(), (x + 1).to(torch.bfloat16).float(), atol=5e-2)  # bf16 wire format
    assert [c[:3] for c in engine.calls] == [("prefill", 1, 16), ("prefill", 1, 16), ("step", 1, 8)]
    assert [c[3] for c in engine.calls] == [0, 16, 32] and sess.position == 40
    ops = [(c["op"], c.get("T"), c.get("pos")) for c in ring.sent]
    assert ops == [("open", None, None), ("prefill", 16, 0), ("prefill", 16, 16), ("step", 8, 32)]
    leader.inference_step(sess, torch.randn(1, 1, H))
    assert engine.calls[-1] == ("step", 1, 1, 40) and ring.sent[-1]["op"] == "step" and ring.sent[-1]["pos"] == 40
    # rollback is propagated through the position of the next command
    sess.set_position(30)
    leader.inference_step(sess, torch.randn(1, 2, H))
    assert ring.sent[-1]["pos"] == 30 and sess.position == 32
    sess.close()
    assert ring.sent[-1] == {"op": "close", "sid": ring.sent[0]["sid"]}


def test_forward_uses_a_scratch_session_and_hypo_ids_reach_the_followers():
    leader, engine, ring, cache, H = _setup(max_prefill_rows=64)
    free_before = cache.tokens_left
    out = leader.forward(torch.randn(2, 20, H))
    assert out.shape == (2, 20, H) and cache.tokens_left == free_before  # scratch KV released
    assert [c["op"] for c in ring.sent] == ["open", "prefill", "close"] and ring.sent[1]["B"] == 2 and ring.sent[1]["T"] == 20
    sess = cache.open_session(2, 64, timeout=0)
    leader.inference_step(sess, torch.randn(2, 1, H))
    leader.inference_step(sess, torch.randn(2, 1, H), hypo_ids=torch.tensor([1, 0]))
    assert ring.sent[-1]["hypo"] == [1, 0]
    sess.close()


def _backward_worker(rank, world, port, results):
    import os
    import types

    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from petals_b200.parallel.tensor_parallel import shard_block
        from petals_b200.parallel.tp_generic import span_backward
        from petals_b200.parallel.tp_worker import tp_collective_backward
        from petals_b200.server.from_pretrained import load_pretrained_block
        from petals_b200.utils.auto_config import AutoDistributedConfig
        from tests.utils import checkpoint

        path = checkpoint("llama", hidden_size=256, intermediate_size=512, num_attention_heads=8, num_key_value_heads=4)
        spec = AutoDistributedConfig.from_pretrained(path).block_spec()
        dense = [load_pretrained_block(path, i, torch_dtype=torch.bfloat16) for i in range(3)]
        # what a TP worker holds: the engine's own Megatron shards (parallel/tensor_parallel.py), nothing dense
        engine = types.SimpleNamespace(device=torch.device("cpu"), spec=spec, rank=rank, world=world, heap=types.SimpleNamespace(group=None),
                                       shards=[shard_block(b, spec, rank, world, "cpu") for b in dense])
        torch.manual_seed(0)
        B, T, H = 2, 6, spec.hidden_size
        shapes = [[1, 2], None]
        if rank == 0:  # the leader owns the request tensors, followers learn the shapes from the command and receive by broadcast
            x, g = torch.randn(B, T, H).bfloat16(), torch.randn(B, T, H).bfloat16()
            prompts = [torch.randn(1, 2, H).bfloat16(), None]
            grad, gp = tp_collective_backward(engine, x, g, prompts, 1, 3, shapes)
            ref, ref_p = span_backward(dense[1:3], x, g, prompts)
            err = (grad.float() - ref.float()).abs().max().item() / ref.float().abs().max().item()
            errp = (gp[0].float() - ref_p[0].float()).abs().max().item() / ref_p[0].float().abs().max().item()
            results[rank] = (err < 3e-2 and errp < 3e-2 and gp[1] is None, err, errp)
        else:
            grad, gp = tp_collective_backward(engine, None, None, None, 1, 3, {"BT": (B, T), "prompts": shapes})
            results[rank] = (bool(torch.isfinite(grad.float()).all()) and grad.shape == (B, T, H), 0.0, 0.0)
    finally:
        dist.destroy_process_group()


def test_collective_backward_of_a_tp_worker_group():
    """rpc_backward through a tensor-parallel stage: leader + follower run the autograd recompute on their shards (the tensors
    the NVLink engine serves from), inputs travel by broadcast, partial gradients by all-reduce; equals the dense span."""
    import torch.multiprocessing as mp

    with mp.Manager() as manager:
        results = manager.dict()
        mp.spawn(_backward_worker, args=(2, 29761, results), nprocs=2, join=True)
        out = dict(results)
    assert out[0][0] and out[1][0], out


def test_leader_backward_publishes_the_command_and_validates_shapes(monkeypatch):
    import pytest

    from petals_b200.parallel import tp_worker

    leader, engine, ring, cache, H = _setup()
    seen = {}

    def fake_collective(eng, hidden, grad, prompts, lo, hi, shapes):
        seen.update(lo=lo, hi=hi, shapes=shapes, n_prompts=len(prompts))
        return grad * 2, [None if p is None else torch.ones_like(p) for p in prompts]

    monkeypatch.setattr(tp_worker, "tp_collective_backward", fake_collective)
    x, g = torch.randn(2, 5, H), torch.randn(2, 5, H)
    prompts = [torch.randn(1, 3, H), torch.empty(0)]  # the second block has no deep prompt (DUMMY)
    grad, gp = leader.backward(x, g, prompts, (0, 2))
    assert torch.equal(grad, g * 2) and gp[0].shape == (1, 3, H) and gp[1] is None
    assert ring.sent[-1] == {"op": "backward", "B": 2, "T": 5, "lo": 0, "hi": 2, "prompts": [[1, 3], None]}
    assert seen == {"lo": 0, "hi": 2, "shapes": [[1, 3], None], "n_prompts": 2}
    n_sent = len(ring.sent)
    with pytest.raises(ValueError):  # nothing may reach the followers when the request is malformed: they would wait in a collective
        leader.backward(x, g[:, :4], None, (0, 2))
    with pytest.raises(ValueError):
        leader.backward(x, g, [None], (0, 2))
    assert len(ring.sent) == n_sent

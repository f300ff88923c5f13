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

"""KV rollback and speculative decoding (reference tests/test_speculative_generation.py:18-85)."""
import pytest
import torch

from petals_b200.client.remote_sequential import RemoteSequential
from petals_b200.utils.auto_config import (AutoDistributedConfig, AutoDistributedModelForCausalLM, AutoDistributedSpeculativeModel)
from tests.utils import checkpoint, local_blocks, swarm_of


@pytest.fixture(scope="module")
def served():
    path = checkpoint("llama")
    with swarm_of(path, ["0:2", "2:4"]) as (swarm, servers):
        yield path, swarm


def test_remote_block_with_cache_invalidation_exact_match(served, atol=1e-4):
    path, swarm = served
    config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm)
    seq = RemoteSequential(config, dht=swarm)[2:3]
    block = local_blocks(path, 3)[2]
    x = torch.randn(1, 8, config.hidden_size)
    short = x.clone()
    short[:, 2:5] = torch.randn(1, 3, config.hidden_size)  # garbage that must be forgotten after the rollback
    with torch.inference_mode():
        with seq.inference_session(max_length=8) as sess:
            sess.step(short[:, :5])
            sess.position = 2  # roll the server-side KV back to 2 tokens
            assert sess.position == 2
            out = sess.step(x[:, 2:])
            with pytest.raises(ValueError):
                sess.position = 100  # only backwards
        ref = block(x)[0][:, 2:]
    assert torch.allclose(out, ref, atol=atol)


class _NoisyDraft(torch.nn.Module):
    """Greedy draft = the true model's greedy tokens with some positions deliberately corrupted."""

    def __init__(self, truth: torch.Tensor, noise_every: int = 3):
        super().__init__()
        self.truth, self.noise_every = truth, noise_every
        self.p = torch.nn.Parameter(torch.zeros(1))
        self.calls = 0

    def generate(self, ids, max_new_tokens, do_sample=False):
        L = ids.shape[1]
        nxt = self.truth[:, L: L + max_new_tokens].clone()
        if nxt.shape[1] < max_new_tokens:
            nxt = torch.cat([nxt, torch.zeros(1, max_new_tokens - nxt.shape[1], dtype=torch.int64)], 1)
        self.calls += 1
        if self.calls % 2 == 0 and nxt.shape[1] > 1:
            nxt[:, self.noise_every % nxt.shape[1]] = (nxt[:, self.noise_every % nxt.shape[1]] + 1) % 500
        return torch.cat([ids, nxt], dim=1)


def test_speculative_generation_equals_greedy(served):
    path, swarm = served
    model = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=swarm)
    ids = torch.randint(0, 500, (1, 5), generator=torch.Generator().manual_seed(0))
    greedy = model.generate(ids, max_new_tokens=40)
    draft = _NoisyDraft(greedy)
    spec = AutoDistributedSpeculativeModel.from_pretrained(path, initial_peers=swarm, small_model=draft)
    out = spec.generate(ids, max_new_tokens=40, speculative_chunk=6)
    assert torch.equal(out, greedy)
    assert draft.calls < 40  # several tokens were accepted per remote step


def test_speculative_generation_stops_right_after_eos():
    """An end-of-sequence token accepted in the middle of a verified chunk ends the output there (what greedy decoding would return)."""
    import torch

    from petals_b200.utils.auto_config import AutoDistributedModelForCausalLM, AutoDistributedSpeculativeModel
    from tests.utils import checkpoint, swarm_of

    path = checkpoint("llama")
    with swarm_of(path, ["0:4"]) as (swarm, _):
        plain = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=swarm)
        ids = torch.tensor([[7, 8, 9, 10]])
        with torch.inference_mode():
            greedy = plain.generate(ids, max_new_tokens=12)[0].tolist()
        eos = greedy[len(ids[0]) + 4]  # the 5th generated token plays end-of-sequence
        cut = greedy.index(eos, len(ids[0])) + 1

        class PerfectDraft:  # proposes exactly what the big model will say: whole chunks are accepted, eos lands mid-chunk
            def generate(self, x, max_new_tokens, do_sample=False):
                n = x.shape[1]
                return torch.tensor([greedy[: n + max_new_tokens] + [0] * max(0, n + max_new_tokens - len(greedy))])[:, : n + max_new_tokens]

        spec = AutoDistributedSpeculativeModel.from_pretrained(path, initial_peers=swarm, small_model=PerfectDraft())
        with torch.inference_mode():
            out = spec.generate(ids, max_new_tokens=12, speculative_chunk=8, eos_token_id=eos)[0].tolist()
        assert out == greedy[:cut]

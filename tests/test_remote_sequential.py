"""RemoteSequential forward/backward vs local blocks, slicing, deep prompts (reference tests/test_remote_sequential.py,
tests/test_chained_calls.py, tests/test_block_exact_match.py)."""
import pytest
import torch

from petals_b200.client.remote_sequential import RemoteSequential
from petals_b200.utils.auto_config import AutoDistributedConfig
from petals_b200.utils.misc import DUMMY
from tests.utils import checkpoint, local_blocks, swarm_of


@pytest.fixture(scope="module")
def served():
    path = checkpoint("llama")
    with swarm_of(path, ["0:3", "1:4", "3:4"]) as (swarm, servers):  # overlapping spans: several possible routes
        config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm)
        yield path, config, swarm


def test_forward_backward_matches_local(served):
    path, config, swarm = served
    seq = RemoteSequential(config, dht=swarm)
    assert len(seq) == config.num_hidden_layers
    blocks = local_blocks(path, config.num_hidden_layers)
    torch.manual_seed(0)
    x = torch.randn(3, 5, config.hidden_size, requires_grad=True)
    out = seq(x)
    out.pow(2).sum().backward()
    g_remote = x.grad.clone()
    x.grad = None
    h = x
    for b in blocks:
        h = b(h)[0]
    h.pow(2).sum().backward()
    assert torch.allclose(out, h, atol=1e-4)
    assert torch.allclose(g_remote, x.grad, atol=1e-3)
    # slicing: the two halves chained equal the whole
    first, second = seq[: len(seq) // 2], seq[len(seq) // 2:]
    assert len(first) + len(second) == len(seq)
    with torch.no_grad():
        assert torch.allclose(second(first(x)), out, atol=1e-4)
    assert len(seq[1]) == 1 and len(list(iter(seq))) == len(seq)


def test_micro_batching_is_invisible(served, monkeypatch):
    path, config, swarm = served
    from petals_b200.client import sequential_autograd

    seq = RemoteSequential(config, dht=swarm)
    x = torch.randn(6, 4, config.hidden_size)
    with torch.no_grad():
        whole = seq(x)
        monkeypatch.setattr(sequential_autograd, "MAX_TOKENS_IN_BATCH", 8)  # => micro-batches of 2 sequences, pipelined
        split = seq(x)
    assert torch.allclose(whole, split, atol=1e-5)


def test_deep_prompts_forward_backward(served):
    path, config, swarm = served
    seq = RemoteSequential(config, dht=swarm)
    blocks = local_blocks(path, config.num_hidden_layers)
    torch.manual_seed(1)
    B, T, P = 2, 6, 3
    x = torch.randn(B, T, config.hidden_size, requires_grad=True)
    prompts = torch.randn(len(seq), B, P, config.hidden_size, requires_grad=True)
    out = seq(x, prompts=prompts)
    out.sum().backward()
    gx, gp = x.grad.clone(), prompts.grad.clone()
    x.grad = prompts.grad = None
    h = x
    for i, b in enumerate(blocks):
        h = torch.cat([h[:, :P] + prompts[i], h[:, P:]], dim=1)
        h = b(h)[0]
    h.sum().backward()
    assert torch.allclose(out, h, atol=1e-4)
    assert torch.allclose(gx, x.grad, atol=1e-3)
    assert torch.allclose(gp, prompts.grad, atol=1e-3)


def test_block_exact_match_and_chained_inference(served):
    path, config, swarm = served
    seq = RemoteSequential(config, dht=swarm)
    blocks = local_blocks(path, config.num_hidden_layers)
    x = torch.randn(1, 8, config.hidden_size)
    with torch.inference_mode():
        # one remote block: forward vs a long-then-short inference session vs the local block
        one = seq[2]
        fwd = one(x)
        with one.inference_session(max_length=8) as sess:
            inf = torch.cat([sess.step(x[:, :7]), sess.step(x[:, 7:])], dim=1)
            with pytest.raises(ValueError, match="Maximum length exceeded"):
                sess.step(x[:, :1])
        ref = blocks[2](x)[0]
        assert torch.allclose(fwd, ref, atol=1e-4) and torch.allclose(inf, ref, atol=1e-4)
        # chain of blocks 1..3 with explicit local KV
        chain = seq[1:4]
        with chain.inference_session(max_length=8) as sess:
            got = torch.cat([sess.step(x[:, t: t + 1]) for t in range(8)], dim=1)
        h, caches = x, [None, None, None]
        outs = []
        for t in range(8):
            cur = x[:, t: t + 1]
            for j, b in enumerate(blocks[1:4]):
                cur, caches[j] = b(cur, layer_past=caches[j], use_cache=True)
            outs.append(cur)
        assert torch.allclose(got, torch.cat(outs, dim=1), atol=1e-4)

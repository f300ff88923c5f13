"""Sub-chains of a served model against the same blocks run locally (reference: tests/test_chained_calls.py:17-77, a 3-block
chain forward/backward and a 2-block chained inference vs local blocks with explicit KV).

One swarm is started per module with overlapping spans (block 3 is served twice), so a chain 3..5 may cross servers; every
check is parametrised over chain position and length."""
import pytest
import torch

from petals_b200.client.remote_sequential import RemoteSequential
from petals_b200.server.from_pretrained import load_pretrained_block
from petals_b200.utils.auto_config import AutoDistributedConfig
from tests.utils import checkpoint, swarm_of


@pytest.fixture(scope="module")
def world():
    path = checkpoint("llama", num_hidden_layers=6)
    with swarm_of(path, ["0:4", "3:6"]) as (swarm, _servers):
        config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm)
        local = {i: load_pretrained_block(path, i, torch_dtype=torch.float32) for i in range(config.num_hidden_layers)}
        yield config, swarm, local


def _local_chain(local, first, last, x, caches=None):
    """Run blocks [first, last) locally; with ``caches`` (a dict block -> (k, v)) as a cached decoding step."""
    for i in range(first, last):
        if caches is None:
            (x,) = local[i](x)
        else:
            x, caches[i] = local[i](x, use_cache=True, layer_past=caches.get(i))
    return x


@pytest.mark.parametrize("first,last,seq_len", [(3, 6, 1), (0, 2, 5), (2, 5, 3)])
def test_chain_forward_and_backward_match_local_blocks(world, first, last, seq_len):
    config, swarm, local = world
    chain = RemoteSequential(config, dht=swarm, start_block=first, end_block=last)
    assert len(chain) == last - first
    torch.manual_seed(first * 10 + last)
    x_remote = torch.randn(2, seq_len, config.hidden_size, requires_grad=True)
    x_local = x_remote.detach().clone().requires_grad_(True)
    weights = torch.randn(2, seq_len, config.hidden_size)  # a non-trivial upstream gradient

    y_remote = chain(x_remote)
    (y_remote * weights).sum().backward()
    y_local = _local_chain(local, first, last, x_local)
    (y_local * weights).sum().backward()

    assert torch.allclose(y_remote, y_local, rtol=0, atol=1e-4)
    assert torch.allclose(x_remote.grad, x_local.grad, rtol=0, atol=1e-4)


@pytest.mark.parametrize("first,last", [(3, 5), (1, 5)])
def test_chain_inference_matches_local_blocks_with_explicit_kv(world, first, last):
    config, swarm, local = world
    chain = RemoteSequential(config, dht=swarm, start_block=first, end_block=last)
    tokens = torch.randn(1, 8, config.hidden_size)
    caches, got, expected = {}, [], []
    with torch.inference_mode(), chain.inference_session(max_length=tokens.shape[1]) as session:
        for t in range(tokens.shape[1]):
            step = tokens[:, t: t + 1]
            got.append(session.step(step))
            expected.append(_local_chain(local, first, last, step, caches))
        assert session.position == tokens.shape[1]
    assert torch.allclose(torch.cat(got, 1), torch.cat(expected, 1), rtol=0, atol=1e-4)
    # the cached token-by-token run equals one cache-less pass over the whole sequence
    with torch.no_grad():
        assert torch.allclose(torch.cat(got, 1), _local_chain(local, first, last, tokens), rtol=0, atol=1e-4)


def test_long_prompt_is_ingested_as_a_wavefront_of_chunks():
    """A long step over several stages is cut into chunks that travel through the stages as a wavefront (client/pipeline.py); the
    stages' KV sessions make it exactly the unchunked computation, and the session continues normally afterwards."""
    import torch

    from petals_b200.utils.auto_config import AutoDistributedModelForCausalLM
    from tests.utils import checkpoint, swarm_of

    path = checkpoint("llama")
    with swarm_of(path, ["0:1", "1:3", "3:4"]) as (swarm, _servers):
        plain = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=swarm, pipeline_chunk_tokens=0)
        piped = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=swarm, pipeline_chunk_tokens=8)
        ids = torch.randint(0, plain.config.vocab_size, (2, 45))
        outs = {}
        for name, m in (("plain", plain), ("piped", piped)):
            with torch.inference_mode(), m.inference_session(max_length=64) as sess:
                a = m(ids[:, :37]).logits          # 37 tokens -> chunks of 8, 8, 8, 8, 5 through 3 stages
                b = m(ids[:, 37:38]).logits        # ordinary single-token step on the same caches
                c = m(ids[:, 38:]).logits          # 7 tokens: below 2 x chunk, not pipelined
                outs[name] = torch.cat([a, b, c], 1)
                assert [s.cursor for s in sess._server_sessions] == [45, 45, 45]
        assert (outs["plain"] - outs["piped"]).abs().max().item() < 1e-4 * outs["plain"].abs().max().item() + 1e-5

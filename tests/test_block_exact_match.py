"""A single served block computes what the checkpoint says, by every route a request can take
(reference: tests/test_block_exact_match.py:12-43).

Routes: the cache-less parallel forward; an inference session whose first step is *long* (more rows than
``MAX_SHORT_INFERENCE_TOKENS``: the handler walks the per-block pools) followed by *short* steps (one merged whole-span task);
and the block loaded locally from the same files.  The session must also refuse to grow past its ``max_length``."""
import pytest
import torch

from petals_b200.client.remote_sequential import RemoteSequential
from petals_b200.server.block_functions import MAX_SHORT_INFERENCE_TOKENS as SHORT
from petals_b200.server.from_pretrained import load_pretrained_block
from petals_b200.utils.auto_config import AutoDistributedConfig
from tests.utils import checkpoint, swarm_of


@pytest.fixture(scope="module", params=["llama", "bloom"])
def served(request):
    path = checkpoint(request.param)
    with swarm_of(path, ["0:4"], inference_max_length=512, attn_cache_tokens=2048) as (swarm, _servers):
        yield path, AutoDistributedConfig.from_pretrained(path, initial_peers=swarm), swarm


@pytest.mark.parametrize("index", [0, 2, 3])
def test_every_route_agrees_with_the_local_block(served, index):
    path, config, swarm = served
    remote = RemoteSequential(config, dht=swarm)[index]
    assert len(remote) == 1
    x = torch.randn(1, SHORT + 8, config.hidden_size, generator=torch.Generator().manual_seed(index))
    with torch.no_grad():
        (expected,) = load_pretrained_block(path, index, torch_dtype=torch.float32)(x)

    assert torch.allclose(remote(x), expected, rtol=0, atol=1e-4)  # rpc_forward

    with torch.inference_mode(), remote.inference_session(max_length=x.shape[1]) as session:
        pieces = [session.step(x[:, : SHORT + 1])]  # one long step ...
        pieces += [session.step(x[:, t: t + 1]) for t in range(SHORT + 1, x.shape[1])]  # ... then token by token
        assert session.position == x.shape[1]
        with pytest.raises(ValueError, match="Maximum length exceeded"):
            session.step(x[:, -1:])
    assert torch.allclose(torch.cat(pieces, dim=1), expected, rtol=0, atol=1e-3)

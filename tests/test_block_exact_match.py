"""One remote block: parallel forward == long + short inference steps == the block loaded locally from the checkpoint
(reference: tests/test_block_exact_match.py). "Long" steps (more than MAX_SHORT_INFERENCE_TOKENS rows) go through the
per-block pools, "short" ones through the merged whole-span task."""
import random

import pytest
import torch

from petals_b200.client.remote_sequential import RemoteSequential
from petals_b200.server.block_functions import MAX_SHORT_INFERENCE_TOKENS
from petals_b200.server.from_pretrained import load_pretrained_block
from petals_b200.utils.auto_config import AutoDistributedConfig
from tests.utils import checkpoint, swarm_of


@pytest.mark.parametrize("family", ["llama", "bloom"])
def test_remote_block_exact_match(family, atol_forward=1e-4, atol_inference=1e-3):
    path = checkpoint(family)
    with swarm_of(path, ["0:4"], inference_max_length=512, attn_cache_tokens=2048) as (swarm, _):
        config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm)
        remote_sequential = RemoteSequential(config, dht=swarm)
        block_index = random.randint(0, config.num_hidden_layers - 1)
        remote_block = remote_sequential[block_index]

        torch.manual_seed(block_index)
        inputs = torch.randn(1, MAX_SHORT_INFERENCE_TOKENS + 8, config.hidden_size)
        outputs_forward = remote_block(inputs)

        outputs_inference = []
        with torch.inference_mode():
            with remote_block.inference_session(max_length=inputs.shape[1]) as sess:
                outputs_inference.append(sess.step(inputs[:, : MAX_SHORT_INFERENCE_TOKENS + 1]))  # long step
                for i in range(MAX_SHORT_INFERENCE_TOKENS + 1, inputs.shape[1]):  # short steps
                    outputs_inference.append(sess.step(inputs[:, i: i + 1]))
                with pytest.raises(ValueError, match=r"Maximum length exceeded") as exc_info:
                    sess.step(inputs[:, -1:])
                assert "Maximum length exceeded" in repr(exc_info.value)
        outputs_inference = torch.cat(outputs_inference, dim=1)

        ref_block = load_pretrained_block(path, block_index, torch_dtype=torch.float32)
        with torch.no_grad():
            (outputs_local,) = ref_block(inputs)
        assert torch.allclose(outputs_local, outputs_forward, rtol=0, atol=atol_forward)
        assert torch.allclose(outputs_local, outputs_inference, rtol=0, atol=atol_inference)

"""Chains of remote blocks: forward/backward and token-by-token inference against the same blocks run locally with explicit
KV caches (reference: tests/test_chained_calls.py)."""
import torch

from petals_b200.client.remote_sequential import RemoteSequential
from petals_b200.server.from_pretrained import load_pretrained_block
from petals_b200.utils.auto_config import AutoDistributedConfig
from tests.utils import checkpoint, swarm_of


def _served():
    path = checkpoint("llama", num_hidden_layers=6)
    return path, swarm_of(path, ["0:4", "3:6"])  # block 3 is served twice; the chain 3..5 may cross servers


def test_forward_backward_exact_match(atol_forward=1e-4, atol_backward=1e-4, seq_length=1):
    path, ctx = _served()
    with ctx as (swarm, _):
        config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm)
        remote_blocks = RemoteSequential(config, dht=swarm, start_block=3, end_block=6)
        assert isinstance(remote_blocks, RemoteSequential) and len(remote_blocks) == 3
        ref_blocks = [load_pretrained_block(path, i, torch_dtype=torch.float32) for i in (3, 4, 5)]

        inputs = torch.randn(1, seq_length, config.hidden_size, requires_grad=True)
        outputs_rpc = remote_blocks.forward(inputs)
        outputs_rpc.sum().backward()
        grads_rpc = inputs.grad
        inputs.grad = None

        hidden = inputs
        for block in ref_blocks:
            hidden = block.forward(hidden)[0]
        hidden.sum().backward()
        assert torch.allclose(hidden, outputs_rpc, rtol=0, atol=atol_forward)
        assert torch.allclose(inputs.grad, grads_rpc, rtol=0, atol=atol_backward)


def test_chained_inference_exact_match(atol_inference=1e-4):
    path, ctx = _served()
    with ctx as (swarm, _):
        config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm)
        remote_blocks = RemoteSequential(config, dht=swarm, start_block=3, end_block=5)
        inputs = torch.randn(1, 8, config.hidden_size)
        outputs_inference = []
        with torch.inference_mode(), remote_blocks.inference_session(max_length=inputs.shape[1]) as sess:
            for i in range(inputs.shape[1]):
                outputs_inference.append(sess.step(inputs[:, i: i + 1]))
        outputs_inference = torch.cat(outputs_inference, dim=1)

        ref_blocks = [load_pretrained_block(path, i, torch_dtype=torch.float32) for i in (3, 4)]
        caches = [None, None]
        outputs_ref = []
        with torch.no_grad():
            for i in range(inputs.shape[1]):
                hidden = inputs[:, i: i + 1]
                for j, block in enumerate(ref_blocks):
                    hidden, caches[j] = block.forward(hidden, use_cache=True, layer_past=caches[j])
                outputs_ref.append(hidden)
        assert torch.allclose(torch.cat(outputs_ref, dim=1), outputs_inference, rtol=0, atol=atol_inference)

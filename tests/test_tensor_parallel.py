"""Sharding math of the tensor-parallel engine (reference: tests/test_tensor_parallel.py — block vs TensorParallel(block)).

The NVLink engine itself is exercised on >= 2 GPUs by tests/test_multi_gpu.py (tools/tp_selftest.py); here the Megatron split
(`shard_block`, `local_spec`) is checked on CPU: summing the per-rank partial outputs where the engine all-reduces must
reproduce the dense block — outputs, KV-cache continuation and gradients."""
import pytest
import torch

from petals_b200.parallel.tensor_parallel import local_spec, shard_oracle_blocks, tp_oracle_forward, tp_supported
from petals_b200.server.from_pretrained import load_pretrained_block
from petals_b200.utils.auto_config import AutoDistributedConfig
from tests.utils import checkpoint


@pytest.mark.parametrize("world", [2, 4])
def test_tp_block_matches_dense(world):
    path = checkpoint("llama", hidden_size=512, intermediate_size=1024, num_attention_heads=8, num_key_value_heads=4)
    config = AutoDistributedConfig.from_pretrained(path)
    spec = config.block_spec()
    assert tp_supported(spec, world)
    block = load_pretrained_block(path, 1, torch_dtype=torch.float32)
    shards = shard_oracle_blocks(block, spec, world)
    ls = local_spec(spec, world)
    assert ls.num_heads * world == spec.num_heads and ls.num_kv_heads * world == spec.num_kv_heads
    assert ls.intermediate_size * world == spec.intermediate_size and ls.hidden_size == spec.hidden_size
    # every weight element lives on exactly one rank (norms are replicated)
    assert sum(s.wqkv.numel() + s.wo.numel() + s.w_gate.numel() + s.w_up.numel() + s.w_down.numel() for s in shards) == \
        block.wqkv.numel() + block.wo.numel() + block.w_gate.numel() + block.w_up.numel() + block.w_down.numel()

    torch.manual_seed(0)
    B, prefix, T, H = 2, 5, 3, spec.hidden_size
    xp1 = torch.randn(B, prefix, H, requires_grad=True)
    x1 = torch.randn(B, T, H, requires_grad=True)
    xp2, x2 = xp1.detach().clone().requires_grad_(True), x1.detach().clone().requires_grad_(True)
    grad_proj = torch.rand(B, T, H)

    L = prefix + T
    kc = torch.zeros(B, L, spec.num_kv_heads, spec.head_dim)
    vc = torch.zeros_like(kc)
    y_prefix_ref = block.forward_cached(xp1, kc, vc, 0)
    y_ref = block.forward_cached(x1, kc.clone(), vc.clone(), prefix)  # clones: in-place cache writes must not break autograd
    y_ref.backward(grad_proj)

    caches = [(torch.zeros(B, L, ls.num_kv_heads, ls.head_dim), torch.zeros(B, L, ls.num_kv_heads, ls.head_dim)) for _ in range(world)]
    y_prefix = tp_oracle_forward(shards, xp2, caches, 0)
    y = tp_oracle_forward(shards, x2, [(k.clone(), v.clone()) for k, v in caches], prefix)
    y.backward(grad_proj)

    assert torch.allclose(y_prefix, y_prefix_ref, atol=1e-5)
    assert torch.allclose(y, y_ref, atol=1e-5)
    assert torch.allclose(x1.grad, x2.grad, atol=1e-4)
    # per-rank KV shards are disjoint slices of the dense cache
    dense_k = torch.cat([c[0] for c in caches], dim=2)
    assert torch.allclose(dense_k[:, :prefix], kc[:, :prefix], atol=1e-5)


@pytest.mark.parametrize("world", [2, 4])
def test_tp_moe_block_matches_dense(world):
    """Sparse-MoE blocks in the engine's split (parallel/tensor_parallel.py:shard_block): every rank holds 1/world of EVERY expert's FFN
    columns and the whole router; the ranks' partial MoE outputs sum to the dense block's output."""
    path = checkpoint("mixtral", hidden_size=256, intermediate_size=256, num_attention_heads=4, num_key_value_heads=4)  # head_dim 64
    config = AutoDistributedConfig.from_pretrained(path)
    spec = config.block_spec()
    assert tp_supported(spec, world)
    block = load_pretrained_block(path, 0, torch_dtype=torch.float32)
    shards = shard_oracle_blocks(block, spec, world)
    assert all(torch.equal(s.router, block.router) for s in shards)
    assert sum(s.we_gate.numel() + s.we_up.numel() + s.we_down.numel() for s in shards) == block.we_gate.numel() + block.we_up.numel() + block.we_down.numel()
    torch.manual_seed(0)
    x = torch.randn(2, 9, spec.hidden_size)
    assert torch.allclose(tp_oracle_forward(shards, x), block.forward_cached(x, None, None, 0), atol=2e-5)


def test_tp_support_matrix():
    llama = AutoDistributedConfig.from_pretrained(checkpoint("llama", hidden_size=512, intermediate_size=1024, num_attention_heads=8, num_key_value_heads=4)).block_spec()
    assert tp_supported(llama, 2) and tp_supported(llama, 4) and not tp_supported(llama, 8) and not tp_supported(llama, 3)
    falcon = AutoDistributedConfig.from_pretrained(checkpoint("falcon")).block_spec()
    assert not tp_supported(falcon, 2)  # fused interleaved QKV + parallel attention: served by pipeline stages instead


def test_engine_refuses_layouts_whose_tensors_it_does_not_shard():
    """`shard_block` carries no projection biases and no ALiBi slopes: such blocks must go to the generic path instead of
    silently losing them (e.g. Llama-style checkpoints with `attention_bias=true`)."""
    import dataclasses

    from petals_b200.parallel.tp_generic import tp_shardable

    llama = AutoDistributedConfig.from_pretrained(checkpoint("llama", hidden_size=512, intermediate_size=1024, num_attention_heads=8, num_key_value_heads=4)).block_spec()
    assert tp_supported(llama, 2)
    for field in ("qkv_bias", "out_bias", "mlp_bias", "alibi"):
        odd = dataclasses.replace(llama, **{field: True})
        assert not tp_supported(odd, 2) and tp_shardable(odd, 2), field
    biased = AutoDistributedConfig.from_pretrained(checkpoint("llama", hidden_size=512, intermediate_size=1024, num_attention_heads=8, num_key_value_heads=4,
                                                              attention_bias=True)).block_spec()
    assert biased.qkv_bias and not tp_supported(biased, 2)

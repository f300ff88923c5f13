"""Tensor parallelism of every block family on plain devices (reference tests/test_tensor_parallel.py:13-49: a block vs
``TensorParallel(block)`` over 2/3/4 fake CPU devices, forward + backward + KV cache; the reference only covers BLOOM).

``TensorParallelBlock`` must reproduce the dense block for: BLOOM (per-head interleaved QKV, ALiBi, biases, LayerNorm),
Falcon-40B style (grouped interleaved QKV, parallel attention, two LayerNorms), Falcon-7B style (multi-query: one kv head
replicated on every rank), Falcon-RW style (ALiBi, sequential, biases), Llama (GQA, RoPE, SwiGLU) and Mixtral (experts
split along the FFN dimension, replicated router) — outputs, KV-cache continuation and input gradients."""
import pytest
import torch

from petals_b200.parallel.tp_generic import TensorParallelBlock, make_shards, make_tensor_parallel, shard_spec, tp_shardable
from petals_b200.server.from_pretrained import load_pretrained_block
from petals_b200.utils.auto_config import AutoDistributedConfig
from tests.utils import checkpoint

FAMILIES = {
    "llama": ("llama", dict(num_attention_heads=8, num_key_value_heads=4)),
    "llama-attention-bias": ("llama", dict(num_attention_heads=8, num_key_value_heads=4, attention_bias=True)),
    "mixtral": ("mixtral", dict(num_attention_heads=8, num_key_value_heads=4)),
    "bloom": ("bloom", dict(n_head=8)),
    "falcon-40b-style": ("falcon", dict(num_attention_heads=8, num_kv_heads=4)),
    "falcon-7b-style": ("falcon", dict(num_attention_heads=8, new_decoder_architecture=False, multi_query=True, parallel_attn=True)),
    "falcon-rw-style": ("falcon", dict(num_attention_heads=8, new_decoder_architecture=False, multi_query=False, parallel_attn=False,
                                       alibi=True, bias=True)),
}


def _load(name):
    family, overrides = FAMILIES[name]
    path = checkpoint(family, **overrides)
    config = AutoDistributedConfig.from_pretrained(path)
    return config.block_spec(), load_pretrained_block(path, 1, torch_dtype=torch.float32)


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("name", sorted(FAMILIES))
def test_tp_block_matches_dense_block(name, world):
    spec, block = _load(name)
    assert tp_shardable(spec, world)
    tp = TensorParallelBlock(block, spec, ["cpu"] * world)
    # every sharded weight element lives on exactly one rank; replicated ones (norms, router, mqa k/v) on all of them
    dense_sharded = sum(getattr(block, n).numel() for n in spec.param_shapes()
                        if n in ("wo", "w_gate", "w_up", "w_down", "we_gate", "we_up", "we_down", "b_up"))
    assert sum(getattr(s, n).numel() for s in tp.shards for n in s.spec.param_shapes()
               if n in ("wo", "w_gate", "w_up", "w_down", "we_gate", "we_up", "we_down", "b_up")) == dense_sharded

    torch.manual_seed(0)
    B, prefix, T, H = 2, 5, 3, spec.hidden_size
    xp = torch.randn(B, prefix, H)
    x1 = torch.randn(B, T, H, requires_grad=True)
    x2 = x1.detach().clone().requires_grad_(True)
    grad_proj = torch.rand(B, T, H)
    L = prefix + T + 1

    def run(module, x):
        kc = torch.zeros(B, L, spec.num_kv_heads, spec.head_dim)
        vc = torch.zeros_like(kc)
        with torch.no_grad():
            y_prefix = module.forward_cached(xp, kc, vc, 0)
        # clones: the in-place cache writes of the no-grad prefix must stay out of the autograd graph of the step
        y = module.forward_cached(x, kc.clone(), vc.clone(), prefix)
        y.backward(grad_proj)
        with torch.no_grad():
            y_next = module.forward_cached(x.detach()[:, :1], *(c := (kc, vc)), prefix)  # writes position `prefix` of the shared cache
        return y_prefix, y, y_next, kc, vc

    ref = run(block, x1)
    got = run(tp, x2)
    for a, b in zip(ref[:3], got[:3]):
        assert torch.allclose(a, b, atol=2e-5), (a - b).abs().max()
    assert torch.allclose(x1.grad, x2.grad, atol=1e-4)
    # the dense KV cache of the primary device holds what the unsharded block would have written (all heads, all positions)
    assert torch.allclose(ref[3], got[3], atol=1e-5) and torch.allclose(ref[4], got[4], atol=1e-5)
    # the HF-style entry point agrees too
    out_ref, (k_ref, _) = block(xp, use_cache=True)
    out_tp, (k_tp, _) = tp(xp, use_cache=True)
    assert torch.allclose(out_ref, out_tp, atol=2e-5) and torch.allclose(k_ref, k_tp, atol=1e-5)


def test_shard_specs_and_support_matrix():
    spec, block = _load("bloom")
    s1 = shard_spec(spec, 1, 4)
    assert s1.num_heads == 2 and s1.num_kv_heads == 2 and s1.alibi_total_heads == 8 and s1.alibi_head_offset == 2
    assert tp_shardable(spec, 8) and not tp_shardable(spec, 3)
    mqa, _ = _load("falcon-7b-style")
    assert mqa.num_kv_heads == 1 and tp_shardable(mqa, 4) and shard_spec(mqa, 3, 4).num_kv_heads == 1
    # row-parallel biases live on rank 0 only
    rw_spec, rw_block = _load("falcon-rw-style")
    shards = make_shards(rw_block, rw_spec, ["cpu", "cpu"])
    assert torch.equal(shards[0].bo, rw_block.bo) and shards[1].bo.abs().sum() == 0
    assert torch.equal(shards[0].b_down, rw_block.b_down) and shards[1].b_down.abs().sum() == 0
    with pytest.raises(ValueError):
        TensorParallelBlock(block, spec, ["cpu"] * 3)
    assert make_tensor_parallel(block, spec, ["cpu"]) is block


def test_three_way_split():
    """The reference also checks 3 devices; that needs head counts divisible by 3."""
    path = checkpoint("bloom", n_head=6, hidden_size=192)
    config = AutoDistributedConfig.from_pretrained(path)
    spec, block = config.block_spec(), load_pretrained_block(path, 0, torch_dtype=torch.float32)
    # FFN 4*192 = 768 is divisible by 3
    tp = TensorParallelBlock(block, spec, ["cpu"] * 3)
    x = torch.randn(2, 4, spec.hidden_size)
    assert torch.allclose(block.forward_cached(x, None, None, 0), tp.forward_cached(x, None, None, 0), atol=2e-5)


@pytest.mark.parametrize("family", ["bloom", "mixtral"])
def test_server_with_tensor_parallel_cpu_devices(family):
    """The reference CI's server4: ``--tensor_parallel_devices cpu cpu`` must serve the same numbers as a plain stage
    (parallel forward, token-by-token inference with a prefix, backward) — .github/workflows/run-tests.yaml:81-83."""
    from petals_b200.client.remote_sequential import RemoteSequential
    from petals_b200.parallel.tp_generic import TensorParallelBlock
    from tests.utils import local_blocks, swarm_of

    path = checkpoint(family)
    with swarm_of(path, ["0:4"], tensor_parallel_devices=["cpu", "cpu"]) as (swarm, servers):
        stage = servers[0].module_container.stage
        assert all(isinstance(b, TensorParallelBlock) and b.world == 2 for b in stage.blocks)
        config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm)
        seq = RemoteSequential(config, dht=swarm)
        blocks = local_blocks(path, config.num_hidden_layers)
        torch.manual_seed(0)
        x = torch.randn(2, 6, config.hidden_size, requires_grad=True)
        out = seq(x)
        out.pow(2).sum().backward()
        g = x.grad.clone()
        x.grad = None
        h = x
        for b in blocks:
            h = b(h)[0]
        h.pow(2).sum().backward()
        assert torch.allclose(out, h, atol=1e-4) and torch.allclose(g, x.grad, atol=1e-3)
        with torch.no_grad(), seq.inference_session(max_length=8) as sess:
            steps = torch.cat([sess.step(x[:, :4]), sess.step(x[:, 4:5]), sess.step(x[:, 5:6])], dim=1)
        assert torch.allclose(steps, h, atol=1e-4)
        info = servers[0].module_container.handler.rpc_info()
        assert info["cache_tokens_available"] > 0


def test_measure_compute_rps_with_tensor_parallel_devices():
    """reference tests/test_aux_functions.py:38-58 measures the self-benchmark with and without TP on (cpu, cpu)."""
    from petals_b200.server.throughput import measure_compute_rps

    config = AutoDistributedConfig.from_pretrained(checkpoint("bloom"))
    for inference in (True, False):
        rps = measure_compute_rps(config, torch.device("cpu"), torch.float32, tensor_parallel_devices=(torch.device("cpu"), torch.device("cpu")),
                                  n_tokens=2, n_steps=2, inference=inference)
        assert rps > 0


def test_tensor_parallel_refuses_quantisation_and_adapters():
    from petals_b200.utils.convert_block import QuantType, convert_block

    spec, block = _load("llama")
    path = checkpoint(*FAMILIES["llama"][:1], **FAMILIES["llama"][1])
    config = AutoDistributedConfig.from_pretrained(path)
    with pytest.raises(ValueError, match="tensor-parallel"):
        convert_block(block, 0, config, ["cpu", "cpu"], torch.device("cpu"), QuantType.FP8)
    tp = convert_block(block, 0, config, ["cpu", "cpu"], torch.device("cpu"), QuantType.NONE)
    assert isinstance(tp, TensorParallelBlock) and tp.tensor_parallel_devices == (torch.device("cpu"), torch.device("cpu"))


def _dist_worker(rank, world, port, name, results):
    import os

    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from petals_b200.parallel.tp_generic import ShardedBlock, span_backward, span_forward

        spec, block = _load(name)
        sharded = ShardedBlock.from_block(block, spec)
        assert sharded.shard.spec.num_heads * world == spec.num_heads
        torch.manual_seed(0)
        x, g = torch.randn(2, 5, spec.hidden_size), torch.randn(2, 5, spec.hidden_size)
        prompts = [torch.randn(1, 2, spec.hidden_size), None]  # a broadcast deep prompt on the first block only
        y = span_forward([sharded, sharded], x, prompts)
        gi, gp = span_backward([sharded, sharded], x, g, prompts)
        y_ref = span_forward([block, block], x, prompts)
        gi_ref, gp_ref = span_backward([block, block], x, g, prompts)
        ok = (torch.allclose(y, y_ref, atol=1e-4, rtol=1e-4) and torch.allclose(gi, gi_ref, atol=1e-4, rtol=1e-4)
              and torch.allclose(gp[0], gp_ref[0], atol=1e-4, rtol=1e-4) and gp[1] is None and gp[0].shape == prompts[0].shape)
        # incremental decoding against this rank's own kv heads
        ls = sharded.shard.spec
        kc = torch.zeros(2, 8, ls.num_kv_heads, ls.head_dim)
        vc = torch.zeros_like(kc)
        with torch.no_grad():
            steps = torch.cat([sharded.forward_cached(x[:, :3], kc, vc, 0), sharded.forward_cached(x[:, 3:], kc, vc, 3)], 1)
            ok = ok and torch.allclose(steps, block.forward_cached(x, None, None, 0), atol=1e-4, rtol=1e-4)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["llama", "bloom", "falcon-7b-style", "mixtral"])
def test_one_shard_per_process_forward_backward(name):
    """The collective form (Megatron f/g autograd functions over a gloo group of 2 processes): the same numbers as the dense
    block on every rank — forward, input gradients, deep-prompt gradients, cached decoding."""
    import torch.multiprocessing as mp

    port = 29700 + sorted(FAMILIES).index(name)
    with mp.Manager() as manager:
        results = manager.dict()
        mp.spawn(_dist_worker, args=(2, port, name, results), nprocs=2, join=True)
        assert dict(results) == {0: True, 1: True}

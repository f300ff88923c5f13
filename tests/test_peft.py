"""LoRA adapters (reference tests/test_peft.py:24-66 + the LoRA branch of test_full_model): safetensors-only policy,
per-block loading, per-request activation, numerics vs merged weights."""
import json
import os

import pytest
import torch

from petals_b200.utils.auto_config import AutoDistributedConfig, AutoDistributedModelForCausalLM
from petals_b200.utils.peft import check_peft_repository, load_peft
from petals_b200.utils.safetensors_io import save_file
from tests.utils import checkpoint, local_blocks, swarm_of


def make_adapter(path: str, config, r: int = 4, alpha: int = 8, seed: int = 0) -> str:
    os.makedirs(path, exist_ok=True)
    spec = config.block_spec()
    g = torch.Generator().manual_seed(seed)
    state = {}
    shapes = {"self_attn.q_proj": (spec.num_heads * spec.head_dim, spec.hidden_size), "self_attn.v_proj": (spec.num_kv_heads * spec.head_dim, spec.hidden_size),
              "mlp.down_proj": (spec.hidden_size, spec.intermediate_size)}
    for layer in range(config.num_hidden_layers):
        for mod, (out_f, in_f) in shapes.items():
            state[f"base_model.model.model.layers.{layer}.{mod}.lora_A.weight"] = torch.randn(r, in_f, generator=g) * 0.1
            state[f"base_model.model.model.layers.{layer}.{mod}.lora_B.weight"] = torch.randn(out_f, r, generator=g) * 0.1
    save_file(state, os.path.join(path, "adapter_model.safetensors"))
    with open(os.path.join(path, "adapter_config.json"), "w") as f:
        json.dump({"peft_type": "LORA", "r": r, "lora_alpha": alpha, "bias": "none", "target_modules": ["q_proj", "v_proj", "down_proj"]}, f)
    return path


def test_safetensors_only_policy(tmp_path):
    config = AutoDistributedConfig.from_pretrained(checkpoint("llama"))
    good = make_adapter(str(tmp_path / "good"), config)
    assert check_peft_repository(good)
    bad = tmp_path / "bad"
    bad.mkdir()
    (bad / "adapter_config.json").write_text("{}")
    (bad / "adapter_model.bin").write_bytes(b"pickle")
    assert not check_peft_repository(str(bad))
    with pytest.raises(ValueError, match="safetensors"):
        load_peft(str(bad), block_idx=0)
    cfg, state = load_peft(good, block_idx=2)
    assert state and all(".2." in k for k in state)  # only the requested block is read


def test_lora_served_per_request(tmp_path):
    path = checkpoint("llama")
    config = AutoDistributedConfig.from_pretrained(path)
    adapter = make_adapter(str(tmp_path / "adapter"), config)
    with swarm_of(path, ["0:4"], adapters=[adapter]) as (swarm, servers):
        plain = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=swarm)
        tuned = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=swarm, active_adapter=adapter)
        ids = torch.randint(0, 500, (1, 6))
        with torch.inference_mode():
            base, lora = plain(ids).logits, tuned(ids).logits
            # oracle: merge the LoRA deltas into local copies of the blocks
            cfg, state = load_peft(adapter)
            scale = cfg["lora_alpha"] / cfg["r"]
            spec = config.block_spec()
            qd, kd = spec.num_heads * spec.head_dim, spec.num_kv_heads * spec.head_dim
            h = plain.model.embed(ids)
            for i, block in enumerate(local_blocks(path, config.num_hidden_layers)):
                pre = f"base_model.model.model.layers.{i}."
                dq = state[pre + "self_attn.q_proj.lora_B.weight"] @ state[pre + "self_attn.q_proj.lora_A.weight"] * scale
                dv = state[pre + "self_attn.v_proj.lora_B.weight"] @ state[pre + "self_attn.v_proj.lora_A.weight"] * scale
                dd = state[pre + "mlp.down_proj.lora_B.weight"] @ state[pre + "mlp.down_proj.lora_A.weight"] * scale
                block.wqkv.data[:qd] += dq
                block.wqkv.data[qd + kd:] += dv
                block.w_down.data += dd
                h = block(h)[0]
            ref = plain.lm_head(plain.model.final_norm(h))
            again = plain(ids).logits  # the adapter of one request must not leak into the next
        assert torch.allclose(lora, ref, atol=1e-3)
        assert not torch.allclose(lora, base, atol=1e-3)
        assert torch.allclose(again, base, atol=1e-5)
        # sessions honour the adapter too
        with torch.inference_mode(), tuned.inference_session(max_length=8):
            sess = torch.cat([tuned(ids[:, :4]).logits, tuned(ids[:, 4:]).logits], 1)
        assert torch.allclose(sess, ref, atol=1e-3)
        with pytest.raises(Exception):
            missing = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=swarm, active_adapter="nope", max_retries=1)
            missing(ids)


def test_merged_adapter_views_equal_the_lora_forward(tmp_path):
    """utils/peft.py:MergedAdapterBlock folds an adapter into copies of the targeted projections (what the kernels serve with
    PETALS_B200_LORA_ENGINE=1): a plain block holding those weights computes exactly what the LoRA branch computes."""
    from petals_b200.models.block_oracle import GenericBlock
    from petals_b200.server.from_pretrained import load_pretrained_block
    from petals_b200.utils.peft import MergedAdapterBlock, add_adapter_to_block, merge_lora, set_active_adapter

    path = checkpoint("llama")
    config = AutoDistributedConfig.from_pretrained(path)
    adapter = make_adapter(str(tmp_path / "adapter"), config)
    block = load_pretrained_block(path, 1, torch_dtype=torch.float32)
    cfg, state = load_peft(adapter, block_idx=1)
    add_adapter_to_block(block, 1, "a", cfg, state)
    view = MergedAdapterBlock(block, "a")
    spec = block.spec
    # q and v pairs land on their rows of the fused projection, the k rows and untargeted matrices are the base tensors themselves
    D, nq, nkv = spec.head_dim, spec.num_heads, spec.num_kv_heads
    assert not torch.equal(view.wqkv[: nq * D], block.wqkv[: nq * D]) and torch.equal(view.wqkv[nq * D: (nq + nkv) * D], block.wqkv[nq * D: (nq + nkv) * D])
    assert not torch.equal(view.wqkv[(nq + nkv) * D:], block.wqkv[(nq + nkv) * D:]) and not torch.equal(view.w_down, block.w_down)
    assert view.w_up is block.w_up and view.ln1_w is block.ln1_w and view.spec is spec and view._p("bqkv") is None
    assert view.merged_bytes == (block.wqkv.numel() + block.w_down.numel()) * 4
    with pytest.raises(AttributeError):
        view.wqkv = None
    with pytest.raises(KeyError):
        MergedAdapterBlock(block, "unknown")

    merged = GenericBlock(spec, dtype=torch.float32)
    with torch.no_grad():
        for name in spec.param_shapes():
            getattr(merged, name).copy_(getattr(view, name))
    x = torch.randn(2, 5, spec.hidden_size)
    set_active_adapter(block, "a")
    with torch.no_grad():
        with_lora = block.forward_cached(x, None, None, 0)
        set_active_adapter(block, None)
        without = block.forward_cached(x, None, None, 0)
        folded = merged.forward_cached(x, None, None, 0)
    assert torch.allclose(folded, with_lora, atol=1e-5) and not torch.allclose(with_lora, without, atol=1e-3)
    # several pairs on one matrix accumulate; fp32 accumulation rounds once
    w = torch.randn(6, 4).to(torch.bfloat16)
    pairs = [(torch.randn(2, 4), torch.randn(6, 2), 0.5, None), (torch.randn(2, 4), torch.randn(3, 2), 2.0, slice(3, 6))]
    ref = w.float() + 0.5 * pairs[0][1] @ pairs[0][0]
    ref[3:6] += 2.0 * pairs[1][1] @ pairs[1][0]
    assert torch.equal(merge_lora(w, pairs), ref.to(torch.bfloat16))


def test_engine_adapter_switching_keeps_separate_graph_caches():
    """server/stage_engine.py:StageEngine.use_adapter swaps (weight views, captured graphs) as a pair."""
    import types

    from petals_b200.server.stage_engine import StageEngine

    base_blocks, base_graphs = [object(), object()], {("g", 0): "base-graph"}
    eng = types.SimpleNamespace(lora_on_engine=True, lora_mode="merged", _adapter=None, _adapter_state={}, blocks=base_blocks, _graphs=base_graphs)
    made = []

    class FakeView:
        merged_bytes = 0

        def __init__(self, base, name):
            made.append((base, name))

    import petals_b200.utils.peft as peft

    real = peft.MergedAdapterBlock
    peft.MergedAdapterBlock = FakeView
    try:
        StageEngine.use_adapter(eng, None)  # no-op
        assert eng.blocks is base_blocks and not made
        StageEngine.use_adapter(eng, "a")
        assert [m[1] for m in made] == ["a", "a"] and [m[0] for m in made] == base_blocks and eng._graphs == {} and eng._adapter == "a"
        eng._graphs[("g", 0)] = "a-graph"
        StageEngine.use_adapter(eng, "b")
        assert [m[0] for m in made[2:]] == base_blocks  # built from the base blocks, not from adapter a's views
        StageEngine.use_adapter(eng, None)
        assert eng.blocks is base_blocks and eng._graphs is base_graphs
        StageEngine.use_adapter(eng, "a")
        assert eng._graphs == {("g", 0): "a-graph"} and len(made) == 4  # cached views and graphs come back
        eng.lora_on_engine = False
        with pytest.raises(RuntimeError):
            StageEngine.use_adapter(eng, "b")
    finally:
        peft.MergedAdapterBlock = real
    # low-rank mode: the adapter shares the base weight views, owns its graph cache, and an unknown name is refused
    blocks = [types.SimpleNamespace(lora_adapters={"a": {}}), types.SimpleNamespace(lora_adapters={"a": {}})]
    eng = types.SimpleNamespace(lora_on_engine=True, lora_mode="lowrank", _adapter=None, _adapter_state={}, blocks=blocks, _graphs={"k": "base"})
    StageEngine.use_adapter(eng, "a")
    assert eng.blocks is blocks and eng._graphs == {} and eng._adapter == "a"
    with pytest.raises(KeyError):
        StageEngine.use_adapter(eng, "missing")
    StageEngine.use_adapter(eng, None)
    assert eng._graphs == {"k": "base"}


def test_adapter_for_another_model_is_refused_at_load_time(tmp_path):
    from petals_b200.server.from_pretrained import load_pretrained_block
    from petals_b200.utils.peft import add_adapter_to_block

    path = checkpoint("llama")
    config = AutoDistributedConfig.from_pretrained(path)
    other = AutoDistributedConfig.from_pretrained(checkpoint("llama", hidden_size=256, intermediate_size=512))
    adapter = make_adapter(str(tmp_path / "wrong"), other)
    block = load_pretrained_block(path, 0, torch_dtype=torch.float32)
    cfg, state = load_peft(adapter, block_idx=0)
    with pytest.raises(ValueError, match="another model"):
        add_adapter_to_block(block, 0, "wrong", cfg, state)
    assert "wrong" not in getattr(block, "lora_adapters", {})
    good_cfg, good_state = load_peft(make_adapter(str(tmp_path / "right"), config), block_idx=0)
    add_adapter_to_block(block, 0, "right", good_cfg, good_state)
    assert set(block.lora_adapters["right"]) == {"wqkv", "w_down"}

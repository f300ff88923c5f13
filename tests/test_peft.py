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

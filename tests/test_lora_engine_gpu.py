"""LoRA requests on the kernels (PETALS_B200_LORA_ENGINE=1: per-adapter merged weight views + graph caches). Written after the
round's GPU budget was spent: runs only with PETALS_B200_RUN_UNVALIDATED=1 until its first hardware run has been looked at."""
import os

import pytest
import torch

from petals_b200.utils.auto_config import AutoDistributedConfig, AutoDistributedModelForCausalLM
from tests.test_peft import make_adapter
from tests.utils import checkpoint, swarm_of

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("PETALS_B200_RUN_UNVALIDATED", "0") != "1",
                                                  reason="first hardware run pending: opt in with PETALS_B200_RUN_UNVALIDATED=1")]


def test_adapter_requests_on_the_engine_match_the_pytorch_executor(tmp_path, monkeypatch):
    path = checkpoint("llama", hidden_size=1024, intermediate_size=2816, num_attention_heads=8, num_key_value_heads=2)
    config = AutoDistributedConfig.from_pretrained(path)
    adapter = make_adapter(str(tmp_path / "adapter"), config)
    ids = torch.randint(0, config.vocab_size, (1, 12), device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))

    def logits(lora_engine: bool):
        monkeypatch.setenv("PETALS_B200_LORA_ENGINE", "1" if lora_engine else "0")
        with swarm_of(path, ["0:4"], device="cuda:0", torch_dtype="bfloat16", adapters=[adapter]) as (swarm, servers):
            engine = servers[0].module_container.stage.engine
            assert engine is not None and engine.lora_on_engine == lora_engine
            plain = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=swarm, torch_dtype=torch.bfloat16).to("cuda:0")
            tuned = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=swarm, torch_dtype=torch.bfloat16, active_adapter=adapter).to("cuda:0")
            with torch.inference_mode():
                out = {}
                for name, model in (("plain", plain), ("tuned", tuned), ("plain_again", plain)):  # adapter on, off, on: switching both ways
                    with model.inference_session(max_length=16):
                        out[name] = torch.cat([model(ids[:, :8]).logits, model(ids[:, 8:9]).logits, model(ids[:, 9:]).logits], 1).float()
                out["tuned_forward"] = tuned(ids).logits.float()
            return out

    ref, got = logits(False), logits(True)
    scale = ref["tuned"].abs().mean()
    assert (ref["tuned"] - ref["plain"]).abs().mean() > 0.01 * scale  # the adapter changes the output at all
    for key in ("plain", "tuned", "tuned_forward"):
        assert (got[key] - ref[key]).abs().mean() < 0.03 * scale, key
    assert torch.equal(got["plain"], got["plain_again"])  # switching back restores the base weights and graphs

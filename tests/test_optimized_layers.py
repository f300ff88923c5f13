"""Every family's block against Hugging Face's own (unoptimised) decoder layer of that architecture, layer by layer, in one
shot and incrementally over a KV cache (reference: tests/test_optimized_layers.py compares its hand-optimised Llama/Falcon
blocks with the stock HF layers the same way)."""
import pytest
import torch

from petals_b200.utils.auto_config import AutoDistributedConfig
from tests.utils import checkpoint, local_blocks

FAMILIES = ["llama", "falcon", "bloom", "mixtral"]


def _hf_hidden_states(path, ids):
    transformers = pytest.importorskip("transformers")
    try:
        hf = transformers.AutoModelForCausalLM.from_pretrained(path, torch_dtype=torch.float32).eval()
        with torch.no_grad():
            out = hf(ids, output_hidden_states=True, use_cache=False)
    except Exception as e:  # noqa: BLE001 - HF version drift must not fail our suite
        pytest.skip(f"transformers cannot run the synthetic checkpoint: {e}")
    return out.hidden_states


@pytest.mark.parametrize("family", FAMILIES)
def test_blocks_match_hf_layers(family, atol=2e-4):
    path = checkpoint(family)
    config = AutoDistributedConfig.from_pretrained(path)
    spec = config.block_spec()
    torch.manual_seed(0)
    B, T = 2, 9
    ids = torch.randint(0, config.vocab_size, (B, T))
    hs = _hf_hidden_states(path, ids)
    n = config.num_hidden_layers
    assert len(hs) == n + 1
    blocks = local_blocks(path, n)
    with torch.no_grad():
        for i in range(n - 1):  # HF applies the final norm to the last entry, so the last layer is covered by test_full_model
            x, want = hs[i], hs[i + 1]
            got = blocks[i](x)[0]
            assert torch.allclose(got, want, atol=atol), f"{family} layer {i}: {(got - want).abs().max().item():.3g}"
            # incremental: a 4-token prefix, then token by token over the block's own cache layout [B, L, Hkv, D]
            kc = torch.zeros(B, T, spec.num_kv_heads, spec.head_dim)
            vc = torch.zeros_like(kc)
            steps = [blocks[i].forward_cached(x[:, :4], kc, vc, 0)]
            for t in range(4, T):
                steps.append(blocks[i].forward_cached(x[:, t: t + 1], kc, vc, t))
            inc = torch.cat(steps, dim=1)
            assert torch.allclose(inc, want, atol=atol), f"{family} layer {i} (cached): {(inc - want).abs().max().item():.3g}"

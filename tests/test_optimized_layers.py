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


def test_sdpa_recompute_path_matches_eager_attention(monkeypatch):
    """PETALS_B200_SDPA_BACKWARD=1 swaps the eager attention of the training recompute pass for the fused SDPA: same
    outputs and input gradients (GQA, causal), and it is only taken when autograd is recording and no KV cache is involved."""
    import torch

    from petals_b200.server.from_pretrained import load_pretrained_block
    from tests.utils import checkpoint

    path = checkpoint("llama", num_attention_heads=8, num_key_value_heads=2)
    eager = load_pretrained_block(path, 0, torch_dtype=torch.float32)
    fused = load_pretrained_block(path, 0, torch_dtype=torch.float32)
    fused.sdpa_backward = True
    torch.manual_seed(0)
    x1 = torch.randn(2, 9, eager.spec.hidden_size, requires_grad=True)
    x2 = x1.detach().clone().requires_grad_(True)
    g = torch.randn_like(x1)
    y1, y2 = eager.forward_cached(x1, None, None, 0), fused.forward_cached(x2, None, None, 0)
    y1.backward(g), y2.backward(g)
    assert torch.allclose(y1, y2, atol=1e-5) and torch.allclose(x1.grad, x2.grad, atol=1e-4)
    calls = []
    real = torch.nn.functional.scaled_dot_product_attention
    monkeypatch.setattr(torch.nn.functional, "scaled_dot_product_attention", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    with torch.no_grad():
        fused.forward_cached(x2.detach(), None, None, 0)  # inference / pass 1: the oracle's own attention
    kc = torch.zeros(2, 16, 2, eager.spec.head_dim)
    fused.forward_cached(x2, kc, torch.zeros_like(kc), 0)  # cached decoding keeps the explicit path too
    assert not calls
    fused.forward_cached(x2, None, None, 0)
    assert calls

"""LoRA requests on the kernels (reference semantics: src/petals/utils/peft.py:173-209, y = W x + scale * B (A x)).

Default mode "lowrank": the base projection runs untouched and every adapted projection gets two extra skinny launches
(x A^T, then (.) B^T accumulated into the output) — no per-adapter weight copies, the adapter request never reaches the PyTorch
executor. Mode "merged": per-adapter merged weight views. Both are compared with the PyTorch executor (oracle blocks with the adapter
active) on the same stage, for prefill, decode-graph capture + replay, a multi-token decode step and the cache-less forward."""
import pytest
import torch

from petals_b200.parallel.swarm import Swarm
from petals_b200.utils.peft import add_adapter_to_block
from petals_b200.utils.random_model import launch_random_stage, write_config_only

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _random_adapter(spec, n_layers, r=8, targets=("q_proj", "v_proj", "o_proj", "gate_proj", "down_proj"), seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = {"q_proj": (spec.num_heads * spec.head_dim, spec.hidden_size), "k_proj": (spec.num_kv_heads * spec.head_dim, spec.hidden_size),
              "v_proj": (spec.num_kv_heads * spec.head_dim, spec.hidden_size), "o_proj": (spec.hidden_size, spec.num_heads * spec.head_dim),
              "gate_proj": (spec.intermediate_size, spec.hidden_size), "up_proj": (spec.intermediate_size, spec.hidden_size),
              "down_proj": (spec.hidden_size, spec.intermediate_size)}
    state = {}
    for layer in range(n_layers):
        for mod in targets:
            out_f, in_f = shapes[mod]
            grp = "self_attn" if mod in ("q_proj", "k_proj", "v_proj", "o_proj") else "mlp"
            state[f"base_model.model.model.layers.{layer}.{grp}.{mod}.lora_A.weight"] = torch.randn(r, in_f, generator=g) * 0.05
            state[f"base_model.model.model.layers.{layer}.{grp}.{mod}.lora_B.weight"] = torch.randn(out_f, r, generator=g) * 0.05
    return {"peft_type": "LORA", "r": r, "lora_alpha": 2 * r, "bias": "none"}, state


def _run(stage, adapter, hidden):
    """Prefill 20 tokens, one-token steps (graph capture, then replay), a 3-token step, and the cache-less forward."""
    st = stage.stage
    st.use_adapter(adapter)
    outs = []
    with torch.inference_mode():
        sess = st.memory_cache.open_session(1, 64, 0.0)
        try:
            for lo, hi in ((0, 20), (20, 21), (21, 22), (22, 25)):
                outs.append(st.inference_step(sess, hidden[:, lo:hi].clone()))
        finally:
            sess.close()
        outs.append(st.forward(hidden[:, :25].clone()))
    torch.cuda.synchronize()
    st.use_adapter(None)
    return torch.cat(outs, 1).float()


@pytest.mark.parametrize("mode", ["lowrank", "merged"])
def test_adapter_requests_on_the_engine_match_the_pytorch_executor(mode, tmp_path, monkeypatch):
    monkeypatch.setenv("PETALS_B200_LORA_ENGINE", mode)
    path = write_config_only("llama-tiny", {}, str(tmp_path / "m"))
    stage = launch_random_stage(path, range(4), Swarm(f"t-lora-{mode}"), DEV)
    try:
        st = stage.stage
        engine = st.engine
        assert engine is not None and engine.lora_on_engine and engine.lora_mode == mode
        cfg, state = _random_adapter(st.spec, 4)
        for i, b in enumerate(st.blocks):
            add_adapter_to_block(b, i, "tuned", cfg, state)
        torch.manual_seed(2)
        hidden = (torch.randn(1, 25, st.spec.hidden_size, device=DEV) * 0.7).to(torch.bfloat16)
        got_plain, got_tuned, got_plain_again = _run(stage, None, hidden), _run(stage, "tuned", hidden), _run(stage, None, hidden)
        oracle_calls = []
        orig = st._oracle_inference
        st._oracle_inference = lambda *a, **k: (oracle_calls.append(1), orig(*a, **k))[1]
        _run(stage, "tuned", hidden)
        assert not oracle_calls, "an adapter request fell back to the PyTorch executor"
        st._oracle_inference = orig
        # reference: the oracle blocks with the adapter's low-rank terms active, over the whole 25-token sequence at once
        from petals_b200.utils.peft import using_adapter
        import contextlib

        with torch.inference_mode(), contextlib.ExitStack() as stack:
            for b in st.blocks:
                stack.enter_context(using_adapter(b, "tuned"))
            h = hidden[:, :25].clone()
            for b in st.blocks:
                h = b.forward_cached(h, None, None, 0)
        ref_tuned = torch.cat([h, h], 1).float()  # _run returns the session outputs followed by the cache-less forward
        scale = ref_tuned.abs().mean().item()
        assert (ref_tuned - got_plain).abs().mean().item() > 0.02 * scale, "the adapter changes nothing: the test would prove nothing"
        err = (got_tuned - ref_tuned).abs()
        assert err.mean().item() < 0.02 * scale, (err.mean().item(), scale)
        assert torch.equal(got_plain, got_plain_again)  # switching back restores the base path (weights and graphs)
    finally:
        stage.shutdown()

"""The persistent one-token span kernel (csrc/decode_span.cu) against the separate-kernel decode path and the fp32 oracle.

Single stream (B = 1, T = 1) is the only shape it takes; the session below crosses a KV page boundary (64 tokens), starts from an
empty cache in one case, and is compared token by token."""
import os

import pytest
import torch

from petals_b200.parallel.swarm import Swarm
from petals_b200.utils.random_model import MODEL_PRESETS, launch_random_stage, random_client_model, write_config_only

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _hidden_states(model, stage, ids, prompt_len, span_kernel: bool):
    """Final hidden states of a session: prompt in one step, then one token per step."""
    engine = stage.stage.engine
    engine.use_span_kernel = span_kernel and engine._span_kernel_ok()
    engine._graphs.clear()
    outs = []
    with torch.inference_mode(), model.inference_session(max_length=ids.shape[1] + 8):
        if prompt_len:
            outs.append(model.model(input_ids=ids[:, :prompt_len]).last_hidden_state)
        for t in range(prompt_len, ids.shape[1]):
            outs.append(model.model(input_ids=ids[:, t:t + 1]).last_hidden_state)
    torch.cuda.synchronize()
    engine.check_errors()
    return torch.cat(outs, 1).float()


@pytest.mark.parametrize("overrides,prompt_len,total", [
    ({}, 60, 70),                                                              # GQA 8/2, D = 128; crosses the page boundary at 64
    ({}, 0, 6),                                                                # empty cache: the first token attends to itself only
    (dict(num_attention_heads=16, num_key_value_heads=16), 30, 36),            # MHA, D = 64
    (dict(num_attention_heads=8, num_key_value_heads=1, intermediate_size=3584), 130, 136),  # one kv head (a tp8 shard's shape), 3 pages
])
def test_span_kernel_matches_separate_kernels_and_oracle(overrides, prompt_len, total, tmp_path):
    path = write_config_only("llama-tiny", overrides, str(tmp_path / "m"))
    swarm = Swarm(f"t-span-{prompt_len}-{len(overrides)}")
    stage = launch_random_stage(path, range(4), swarm, DEV)
    try:
        engine = stage.stage.engine
        assert engine is not None and engine._span_kernel_ok()
        model = random_client_model(path, swarm, DEV)
        torch.manual_seed(1)
        ids = torch.randint(0, 4000, (1, total), device=DEV)
        ref = _hidden_states(model, stage, ids, prompt_len, span_kernel=False)
        got = _hidden_states(model, stage, ids, prompt_len, span_kernel=True)
        assert engine._span_plans, "the span kernel did not run"
        # same rounding points as the separate kernels: only the attention summation order differs
        scale = ref.abs().mean().item()
        err = (got - ref).abs()
        assert err.mean().item() < 2e-2 * scale + 1e-4, (err.mean().item(), scale)
        assert err.max().item() < 0.08 * ref.abs().max().item() + 1e-2, (err.max().item(), ref.abs().max().item())
        # and the fp32 oracle of the whole model
        from tests.test_engine_gpu import _oracle_logits

        with torch.inference_mode():
            oracle = _oracle_logits(model, stage, ids).float()
            logits = model.lm_head(got.to(torch.bfloat16)).float()  # `got` is already final-normed
        assert logits.shape == oracle.shape
        assert (logits.argmax(-1) == oracle.argmax(-1)).float().mean().item() > 0.9
    finally:
        stage.shutdown()


def test_span_kernel_through_generate(tmp_path):
    """Greedy generation is identical with and without the span kernel for a short continuation (same rounding points)."""
    path = write_config_only("llama-tiny", {}, str(tmp_path / "m"))
    swarm = Swarm("t-span-gen")
    stage = launch_random_stage(path, range(4), swarm, DEV)
    try:
        model = random_client_model(path, swarm, DEV)
        engine = stage.stage.engine
        ids = torch.randint(0, 4000, (1, 20), device=DEV)
        outs = []
        for on in (False, True):
            engine.use_span_kernel = on
            engine._graphs.clear()
            outs.append(model.generate(ids, max_new_tokens=12))
        agree = (outs[0] == outs[1]).float().mean().item()
        assert agree > 0.9, (outs[0].tolist(), outs[1].tolist())
    finally:
        stage.shutdown()

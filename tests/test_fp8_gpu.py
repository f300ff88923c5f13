"""Block-scaled FP8 serving (--quant_type fp8) through the engine: close to the bf16 model, exact vs its own dequantised oracle."""
import pytest
import torch

from petals_b200.parallel.swarm import Swarm
from petals_b200.utils.convert_block import QuantType
from petals_b200.utils.random_model import launch_random_stage, random_client_model, write_config_only

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_fp8_stage_matches_dequantised_oracle(tmp_path):
    path = write_config_only("llama-tiny", {}, str(tmp_path / "m"))
    swarm = Swarm("t-fp8")
    stage = launch_random_stage(path, range(4), swarm, DEV, quant_type=QuantType.FP8)
    try:
        eng = stage.stage.engine
        assert eng is not None and eng.fp8 is not None, "fp8 stages must run on the sm_100a engine"
        assert all(getattr(stage.stage.blocks[0], n).numel() == 0 for n in eng.fp8[0]), "bf16 copies must be released"
        model = random_client_model(path, swarm, DEV)
        ids = torch.randint(0, 4000, (1, 24), device=DEV)
        with torch.inference_mode():
            with model.inference_session(max_length=64):
                a = model(ids[:, :20]).logits  # prefill: dequant scratch + tcgen05 GEMM
                b = torch.cat([model(ids[:, t: t + 1]).logits for t in range(20, 24)], 1)  # decode: fp8 weight streamer
            sess = torch.cat([a, b], 1).float()
            # oracle on the dequantised weights
            h = model.model.embed(ids)
            for i, blk in enumerate(stage.stage.blocks):
                stage.stage._materialize(i)
                h = blk.forward_cached(h, None, None, 0)
                stage.stage._dematerialize(i)
            ref = model.lm_head(model.model.final_norm(h)).float()
        scale = ref.abs().mean().item()
        assert (sess - ref).abs().mean().item() < 0.05 * scale + 1e-3
        # gradients still flow (weights are dequantised block by block for the backward pass)
        emb = model.model.embed(ids).float().requires_grad_(True)
        model.model.layers(emb.to(torch.bfloat16)).float().pow(2).mean().backward()
        assert torch.isfinite(emb.grad).all() and emb.grad.abs().sum() > 0
    finally:
        stage.shutdown()

"""Error statistics of the sm_100a engine against the oracle blocks for every preset of tests/test_engine_gpu.py (used to set the
test's per-element bounds from measurements instead of guesses). One JSON line per preset."""
import json
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from petals_b200.parallel.swarm import Swarm
    from petals_b200.utils.random_model import MODEL_PRESETS, launch_random_stage, random_client_model, write_config_only
    from tests.test_engine_gpu import PRESETS, _oracle_logits

    dev = "cuda:0"
    for i, (preset, overrides) in enumerate(PRESETS):
        path = write_config_only(preset, overrides, tempfile.mkdtemp(prefix="pb200-stats-"))
        swarm = Swarm(f"stats-{i}")
        n_layers = overrides.get("n_layer", MODEL_PRESETS[preset].get("num_hidden_layers", MODEL_PRESETS[preset].get("n_layer")))
        stage = launch_random_stage(path, range(n_layers), swarm, dev)
        try:
            model = random_client_model(path, swarm, dev)
            torch.manual_seed(0)
            ids = torch.randint(0, 4000, (2, 40), device=dev)
            with torch.inference_mode():
                ref = _oracle_logits(model, stage, ids).float()
                full = model(ids).logits.float()
                with model.inference_session(max_length=64):
                    parts = [model(ids[:, a:b]).logits for a, b in ((0, 33), (33, 34), (34, 35), (35, 38), (38, 40))]
                sess = torch.cat(parts, 1).float()
            scale = ref.abs().mean().item()
            out = {"preset": preset, "overrides": overrides, "scale": round(scale, 4)}
            for name, got in (("forward", full), ("session", sess)):
                err = (got - ref).abs() / scale
                out[name] = {"mean": round(err.mean().item(), 5), "p999": round(err.flatten().kthvalue(int(err.numel() * 0.999)).values.item(), 4),
                             "max": round(err.max().item(), 4), "argmax_agree": round((got.argmax(-1) == ref.argmax(-1)).float().mean().item(), 4)}
            print(json.dumps(out), flush=True)
        finally:
            stage.shutdown()


if __name__ == "__main__":
    main()

"""Smoke-run the three benchmark scripts with tiny sizes, as the reference CI does after its test suite
(reference .github/workflows/run-tests.yaml:109-116): they must start a private swarm, run, and print a speed."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = ["--model", "llama-tiny", "--device", "cpu", "--torch_dtype", "float32", "--warmup_steps", "1"]


def _run(script, *args):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "benchmarks", script), *COMMON, *args], capture_output=True, text=True,
                         env=env, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    return out.stdout


def _speed(stdout: str, pattern: str) -> float:
    m = re.search(pattern, stdout)
    assert m, stdout[-2000:]
    return float(m.group(1))


def test_benchmark_inference_smoke():
    out = _run("benchmark_inference.py", "--seq_len", "12", "--prompt_len", "4", "--n_stages", "2")
    assert _speed(out, r"Final result: speed=([0-9.]+) tokens/sec") > 0


def test_benchmark_forward_smoke():
    out = _run("benchmark_forward.py", "--seq_len", "16", "--batch_size", "2", "--n_steps", "3")
    assert _speed(out, r"([0-9.]+) tokens/sec") > 0


@pytest.mark.parametrize("task", ["cls", "causal_lm"])
def test_benchmark_training_smoke(task):
    out = _run("benchmark_training.py", "--task", task, "--seq_len", "8", "--batch_size", "2", "--pre_seq_len", "2", "--n_steps", "3")
    assert _speed(out, r"([0-9.]+) tokens/sec") > 0


def test_launch_box_example_serves_a_model_over_a_fabric(tmp_path):
    """examples/launch_box.py: rendezvous + one run_server per device with an even block split, all members of one landing-ring fabric."""
    import signal
    import subprocess
    import sys
    import time

    import torch

    from petals_b200.utils.auto_config import AutoDistributedConfig, AutoDistributedModelForCausalLM
    from tests.utils import checkpoint, local_blocks

    sys.path.insert(0, os.path.join(ROOT, "examples"))
    from launch_box import split_blocks

    assert split_blocks(80, 8) == [(10 * i, 10 * i + 10) for i in range(8)] and split_blocks(4, 3) == [(0, 1), (1, 3), (3, 4)]
    path = checkpoint("llama")
    rendezvous = str(tmp_path / "swarm")
    proc = subprocess.Popen([sys.executable, os.path.join(ROOT, "examples", "launch_box.py"), path, "--gpus", "2", "--device", "cpu", "--rendezvous", rendezvous,
                             "--torch_dtype", "float32", "--fabric_max_tokens", "128", "--", "--throughput", "1", "--update_period", "1"],
                            cwd=ROOT, stdout=open(tmp_path / "box.log", "w"), stderr=subprocess.STDOUT, env=dict(os.environ, PYTHONPATH=ROOT))  # a file, not a
    try:                                                                                                                                  # pipe nobody drains
        model = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=[rendezvous], max_retries=150, min_backoff=0.5, max_backoff=1.0)
        config = AutoDistributedConfig.from_pretrained(path)
        ids = torch.randint(0, config.vocab_size, (1, 6), generator=torch.Generator().manual_seed(0))
        with torch.inference_mode():
            h = model.model.embed(ids)
            for b in local_blocks(path, config.num_hidden_layers):
                h = b(h)[0]
            ref = model.lm_head(model.model.final_norm(h))
            with model.inference_session(max_length=8) as sess:
                got = torch.cat([model(ids[:, :5]).logits, model(ids[:, 5:]).logits], 1)
                hops = [s.no_history for s in sess._server_sessions]
        assert torch.allclose(got, ref, atol=1e-3) and hops == [False, True]
    finally:
        proc.send_signal(signal.SIGTERM)
        try:
            proc.wait(timeout=20)
        except subprocess.TimeoutExpired:
            proc.kill()

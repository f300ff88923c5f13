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

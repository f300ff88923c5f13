"""The stage-hop fabric protocol on CPU: 2 gloo processes, shared-memory landing zones, tensor-less control RPCs."""
import json
import os
import subprocess
import sys

from tests.utils import checkpoint

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fabric_protocol_two_processes():
    path = checkpoint("llama")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29741",
           os.path.join(ROOT, "tools", "pp_selftest_cpu.py"), path]
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
    assert proc.returncode == 0 and lines, proc.stdout[-3000:] + proc.stderr[-3000:]
    report = json.loads(lines[-1])
    assert report["pp_selftest_cpu"] == "ok", report

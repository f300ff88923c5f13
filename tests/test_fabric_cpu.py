"""The stage-hop fabric protocol on CPU: 2 gloo processes, shared-memory landing zones, tensor-less control RPCs."""
import json
import os
import subprocess
import sys

from tests.utils import checkpoint

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize("stages", [2, 4])
def test_fabric_protocol_between_processes(stages):
    """Inference steps (landing zones, rollback) and a training pass (micro-batches hop forward through the x_in rings, gradients hop
    back through the g_in rings, every stage stashes its input) over the shared-memory twin of the NVLink fabric; with 4 stages the
    middle ones both take from and push into rings."""
    path = checkpoint("llama")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={stages}", "--master-addr", "127.0.0.1", "--master-port", str(29741 + stages),
           os.path.join(ROOT, "tools", "pp_selftest_cpu.py"), path]
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
    assert proc.returncode == 0 and lines, proc.stdout[-3000:] + proc.stderr[-3000:]
    report = json.loads(lines[-1])
    assert report["pp_selftest_cpu"] == "ok", report
    assert report["training_fabric_hops"] == {"forward": 3 * stages, "backward": 3 * stages}, report


def test_fabric_rings_stay_in_step_after_a_refused_transfer():
    """A stage that fails a request AFTER its predecessor pushed into its landing slot drains that slot (server/handler.py:drain_landing);
    the client finishes the micro-batch with tensors, pauses fabric passes (cool-down), and later fabric passes work again."""
    path = checkpoint("llama")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29749",
           os.path.join(ROOT, "tools", "pp_selftest_cpu.py"), path]
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="", PP_SELFTEST_FAULT="1", PETALS_B200_FAULTS="rpc=rpc_forward,peer=stage1,after=3,times=1")
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
    assert proc.returncode == 0 and lines, proc.stdout[-3000:] + proc.stderr[-3000:]
    report = json.loads(lines[-1])
    assert report["pp_selftest_cpu"] == "ok", report
    assert report["fault_passes"]["healed"] == {"forward": 6, "backward": 6} and report["fault_passes"]["cooling"] == {"forward": 0, "backward": 0}, report

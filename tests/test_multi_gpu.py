"""Multi-GPU numerics: tensor-parallel engine and the fused pipeline stage hop (spawn torch.distributed.run; skipped on
single-GPU boxes)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_tp2_matches_oracle():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tools", "tp_selftest.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
    assert proc.returncode == 0 and lines, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert json.loads(lines[-1])["tp_selftest"] == "ok"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_tp2_sparse_moe_matches_oracle():
    """Mixtral blocks in the tensor-parallel engine: every expert's FFN columns split over the pair, replicated router, stand-alone LL
    all-reduce halves around the device-routed expert GEMVs (decode) and grouped-GEMM prefill with owner scatter (validated on 2 and 8
    B200s: profiles/r2_bench_mixtral_8x7b_tp8.json `selftests`)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29733",
           os.path.join(ROOT, "tools", "tp_selftest.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, TP_SELFTEST_MODEL="mixtral-tiny"))
    lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
    assert proc.returncode == 0 and lines, proc.stdout[-2000:] + proc.stderr[-2000:]
    report = json.loads(lines[-1])
    assert report["tp_selftest"] == "ok" and report["model"] == "mixtral-tiny"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_pp2_fused_stage_hop_matches_oracle():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29732",
           os.path.join(ROOT, "tools", "pp_selftest.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
    assert proc.returncode == 0 and lines, proc.stdout[-2000:] + proc.stderr[-2000:]
    report = json.loads(lines[-1])
    assert report["pp_selftest"] == "ok"
    # training over the fabric: every forward micro-batch and every gradient hopped through the landing rings, with and without deep prompts
    assert all(h == {"forward": 6, "backward": 6} for h in report["training_fabric_hops"].values()), report


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_server_tensor_parallel_devices():
    """`run_server --tensor_parallel_devices cuda:0 cuda:1`: one server process leads a spawned worker group."""
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "tp_server_selftest.py")], capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
    assert proc.returncode == 0 and lines, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert json.loads(lines[-1])["tp_server_selftest"] == "ok"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.skipif(os.environ.get("PETALS_B200_RUN_UNVALIDATED", "0") != "1",
                    reason="rpc_backward through a TP worker group is validated with gloo on CPU (tests/test_tp_leader_logic.py); its first NCCL run on "
                           "hardware is pending - opt in with PETALS_B200_RUN_UNVALIDATED=1")
def test_server_tensor_parallel_backward():
    """Prompt-tuning gradients through `--tensor_parallel_devices`: every rank recomputes on its shard, partials are all-reduced."""
    env = dict(os.environ, TP_SELFTEST_BACKWARD="1")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "tp_server_selftest.py")], capture_output=True, text=True, timeout=420, cwd=ROOT,
                          env=env)
    lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
    assert proc.returncode == 0 and lines, proc.stdout[-2000:] + proc.stderr[-2000:]
    report = json.loads(lines[-1])
    assert report["tp_server_selftest"] == "ok" and report["backward_rel_err"] < 0.08

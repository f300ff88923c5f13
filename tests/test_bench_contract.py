"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm answers with one JSON line and
exit code 0, and the argument defaults respect the timing rules (>= 3 warm-up steps, N = 1 by default)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_reports_unavailable_or_a_result():
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "3"],
                          capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, proc.stdout[-2000:]
    out = json.loads(lines[0])
    assert out["impl"] == "reference"
    assert "unavailable" in out or {"metric", "value", "unit", "n_gpus"} <= set(out)


def test_defaults_follow_the_timing_rules():
    sys.path.insert(0, ROOT)
    import importlib

    bench = importlib.import_module("bench")
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'add_argument("--gpus", type=int, default=1)' in src
    assert 'add_argument("--warmup", type=int, default=4)' in src and "args.warmup < 3" in src
    assert callable(bench.main) and callable(bench.reference_arm)


def test_a_failed_selftest_invalidates_a_multi_gpu_line():
    """parallel/multi_gpu_bench.py: every multi-GPU bench job runs the TP / pipeline numerics self-tests first; a report that is not
    "ok" (or a self-test that raised) marks the JSON line invalid and fails the run; --skip-selftests is the only way around."""
    from petals_b200.parallel.multi_gpu_bench import _selftests_ok

    assert _selftests_ok({"skipped": True})
    assert _selftests_ok({"tp": {"tp_selftest": "ok"}, "pp": {"pp_selftest": "ok", "world": 8}})
    assert not _selftests_ok({"tp": {"tp_selftest": "ok"}, "pp": {"pp_selftest": "FAILED"}})
    assert not _selftests_ok({"tp": {"tp_selftest": "FAILED", "error": "RuntimeError(...)"}})
    assert not _selftests_ok({"tp": {}})  # a rank-0 report must exist


def test_bench_accepts_the_pipeline_of_tp_groups_layout():
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, cwd=ROOT).stdout
    assert "ppSxtpT" in out and "--skip-selftests" in out and "--skip-pipeline" in out

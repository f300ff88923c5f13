"""Race detection for the native host runtime: the KV page allocator and the prioritised task queue are
hammered from many threads in a ThreadSanitizer build. Any data race or failed invariant fails the test."""
import os
import shutil
import subprocess

import pytest

RT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "petals_b200", "csrc", "runtime")


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_runtime_is_race_free_under_tsan(tmp_path):
    exe = str(tmp_path / "stress_tsan")
    build = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "kv_allocator.cpp", "task_queue.cpp", "tests/stress_tsan.cpp",
                            "-lpthread", "-o", exe], cwd=RT, capture_output=True, text=True)
    if build.returncode != 0 and "tsan" in (build.stderr + build.stdout).lower():
        pytest.skip("this toolchain has no ThreadSanitizer runtime")
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66"))
    out = run.stdout + run.stderr
    assert run.returncode == 0 and "ThreadSanitizer" not in out and "runtime stress ok" in out, out[-3000:]

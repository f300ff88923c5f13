"""Pieces shared by ``bench.py`` (1 GPU) and ``parallel/multi_gpu_bench.py`` (N GPUs): the metric string, the clock sampler
and the two timed loops, so that every N reports the same metric measured the same way.

The timed loops follow the reference's benchmark (benchmarks/benchmark_inference.py:44-68): one inference session, then
``model.generate(max_new_tokens=1, session=sess)`` per step.

* ``device_timed_decode``: K calls bracketed by CUDA events; tokens stay on the device.
* ``e2e_decode``: the same K calls, but every step the token to feed arrives from **pinned host memory** (H2D, through the
  session's public ``last_token_id`` setter) and the sampled token is read back to the host (D2H) before the next step."""
from __future__ import annotations

import subprocess
import threading
import time
from typing import Callable, Optional, Tuple

BASELINE_TOKENS_PER_S = 6.0  # README.md:86 of the reference ("up to 6 tokens/s" single-batch, Llama 2 70B, public swarm)
_LABELS = {"llama-3-70b": "Llama-3-70B", "llama-3-8b": "Llama-3-8B", "mixtral-8x7b": "Mixtral-8x7B"}


def model_label(name: str) -> str:
    return _LABELS.get(name, name)


def metric_name(model: str) -> str:
    """ONE string for every N (the driver checks that the metric does not change along the scaling run)."""
    return f"{model_label(model)} single-stream decode tokens/s (device-timed, max over ranks); prefill tokens/s in `prefill`"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int = 0, period: float = 0.2):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self._halt = index, period, [], threading.Event()

    def run(self) -> None:
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                self.samples.append([x.strip() for x in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._halt.wait(self.period)

    def stop(self) -> dict:
        self._halt.set()
        self.join(timeout=3)
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples if len(s) >= 6 for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def prime_session(model, sess, prompt, warmup: int):
    """Prompt ingestion + W untimed single-token steps (not part of the single-stream metric, like the reference benchmark)."""
    model.generate(prompt, max_new_tokens=1, session=sess)
    for _ in range(warmup):
        model.generate(max_new_tokens=1, session=sess)


def device_timed_decode(model, sess, steps: int, *, on_start: Optional[Callable[[], None]] = None,
                        on_end: Optional[Callable[[], None]] = None) -> Tuple[float, int]:
    """Returns (milliseconds for exactly ``steps`` single-token generate() calls, kernels launched)."""
    import torch

    from petals_b200.ops import native

    torch.cuda.synchronize()
    if on_start is not None:
        on_start()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    before = native.launch_count
    start.record()
    for _ in range(steps):
        model.generate(max_new_tokens=1, session=sess)
    end.record()
    torch.cuda.synchronize()
    if on_end is not None:
        on_end()
    return start.elapsed_time(end), native.launch_count - before


def e2e_decode(model, sess, steps: int, device) -> Tuple[float, int, int]:
    """Returns (seconds, H2D bytes per step, D2H bytes per step) for ``steps`` end-to-end single-token steps."""
    import torch

    pinned_in = torch.zeros(1, 1, dtype=torch.int64).pin_memory()
    pinned_out = torch.zeros(1, dtype=torch.int64).pin_memory()
    pinned_in.copy_(sess.output_ids[:, -1:].cpu())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        sess.last_token_id = pinned_in.to(device, non_blocking=True)  # H2D: the token this step feeds
        out = model.generate(max_new_tokens=1, session=sess)
        pinned_out.copy_(out[0, -1:], non_blocking=True)  # D2H: the token this step produced
        torch.cuda.synchronize()
        pinned_in[0, 0] = pinned_out[0]
    return time.perf_counter() - t0, pinned_in.numel() * pinned_in.element_size(), pinned_out.numel() * pinned_out.element_size()

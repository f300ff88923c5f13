"""Serving metrics of one stage worker.

The reference publishes state instead of exposing it: ``ServerInfo`` records, ``rpc_info`` and log lines.
Those are all kept; this adds what an operator of a fixed 8-GPU deployment scrapes: monotonic counters and latency
histograms per RPC, live session / KV-cache gauges, in the Prometheus text format on ``--metrics_port`` (no dependency:
the exposition format is a few lines) and as a dict under ``rpc_info()["metrics"]``.

Cost on the token path: one lock and a handful of integer adds per step.
"""
from __future__ import annotations

import bisect
import http.server
import threading
import time
from typing import Callable, Dict, List, Optional, Tuple

# seconds; spans a fused decode step (~1e-4 s of host time) to a long chunked prefill
LATENCY_BUCKETS: Tuple[float, ...] = (0.0005, 0.001, 0.002, 0.005, 0.01, 0.02, 0.05, 0.1, 0.25, 0.5, 1.0, 2.5, 5.0, 10.0, 30.0)
RPCS = ("inference", "forward", "backward")


class _Histogram:
    __slots__ = ("counts", "total", "n")

    def __init__(self):
        self.counts = [0] * (len(LATENCY_BUCKETS) + 1)
        self.total, self.n = 0.0, 0

    def observe(self, v: float) -> None:
        self.counts[bisect.bisect_left(LATENCY_BUCKETS, v)] += 1
        self.total += v
        self.n += 1


class ServerMetrics:
    def __init__(self, peer_id: str = ""):
        self.peer_id = peer_id
        self.started_at = time.time()
        self._lock = threading.Lock()
        self.requests: Dict[str, int] = {r: 0 for r in RPCS}
        self.tokens: Dict[str, int] = {r: 0 for r in RPCS}
        self.errors: Dict[str, int] = {r: 0 for r in RPCS}
        self.latency: Dict[str, _Histogram] = {r: _Histogram() for r in RPCS}
        self.sessions_opened = 0
        self.sessions_closed = 0
        self._gauges: Dict[str, Callable[[], Optional[float]]] = {}

    # ---- recording -----------------------------------------------------------------------------------------------------
    def observe(self, rpc: str, tokens: int, seconds: float) -> None:
        with self._lock:
            self.requests[rpc] += 1
            self.tokens[rpc] += int(tokens)
            self.latency[rpc].observe(seconds)

    def error(self, rpc: str) -> None:
        with self._lock:
            self.errors[rpc] += 1

    def session_opened(self) -> None:
        with self._lock:
            self.sessions_opened += 1

    def session_closed(self) -> None:
        with self._lock:
            self.sessions_closed += 1

    def gauge(self, name: str, fn: Callable[[], Optional[float]]) -> None:
        """A value read at scrape time (KV tokens left, queue depth, ...)."""
        self._gauges[name] = fn

    # ---- reading ---------------------------------------------------------------------------------------------------------
    def _read_gauges(self) -> Dict[str, float]:
        out = {}
        for name, fn in self._gauges.items():
            try:
                v = fn()
            except Exception:  # noqa: BLE001 - a gauge must never break a scrape
                v = None
            if v is not None:
                out[name] = float(v)
        return out

    def snapshot(self) -> dict:
        with self._lock:
            out = {
                "uptime_s": round(time.time() - self.started_at, 1),
                "sessions_active": self.sessions_opened - self.sessions_closed, "sessions_opened": self.sessions_opened,
                "requests": dict(self.requests), "tokens": dict(self.tokens), "errors": dict(self.errors),
                "mean_latency_ms": {r: round(1e3 * h.total / h.n, 3) if h.n else None for r, h in self.latency.items()},
            }
        out.update(self._read_gauges())
        return out

    def render_prometheus(self) -> str:
        lab = f'peer="{self.peer_id}"'
        lines: List[str] = []

        def metric(name: str, kind: str, help_: str, samples):
            lines.append(f"# HELP petals_{name} {help_}")
            lines.append(f"# TYPE petals_{name} {kind}")
            for labels, value in samples:
                all_labels = ",".join(x for x in (lab, labels) if x)
                lines.append(f"petals_{name}{{{all_labels}}} {value}")

        with self._lock:
            metric("requests_total", "counter", "RPCs served", [(f'rpc="{r}"', self.requests[r]) for r in RPCS])
            metric("tokens_total", "counter", "tokens processed (batch x new positions)", [(f'rpc="{r}"', self.tokens[r]) for r in RPCS])
            metric("errors_total", "counter", "RPCs that raised", [(f'rpc="{r}"', self.errors[r]) for r in RPCS])
            metric("sessions_opened_total", "counter", "inference sessions opened", [("", self.sessions_opened)])
            metric("sessions_active", "gauge", "inference sessions currently open", [("", self.sessions_opened - self.sessions_closed)])
            lines.append("# HELP petals_request_seconds server-side latency of one RPC (one decode step for rpc=inference)")
            lines.append("# TYPE petals_request_seconds histogram")
            for r in RPCS:
                h, acc = self.latency[r], 0
                for edge, c in zip(LATENCY_BUCKETS, h.counts):
                    acc += c
                    lines.append(f'petals_request_seconds_bucket{{{lab},rpc="{r}",le="{edge}"}} {acc}')
                lines.append(f'petals_request_seconds_bucket{{{lab},rpc="{r}",le="+Inf"}} {h.n}')
                lines.append(f'petals_request_seconds_sum{{{lab},rpc="{r}"}} {h.total:.6f}')
                lines.append(f'petals_request_seconds_count{{{lab},rpc="{r}"}} {h.n}')
        for name, v in self._read_gauges().items():
            metric(name, "gauge", name.replace("_", " "), [("", v)])
        metric("uptime_seconds", "gauge", "seconds since the stage started", [("", round(time.time() - self.started_at, 1))])
        return "\n".join(lines) + "\n"


class MetricsServer:
    """``GET /metrics`` (Prometheus text format) and ``GET /metrics.json`` on ``host:port`` (port 0 = pick one)."""

    def __init__(self, metrics_of: Callable[[], Optional[ServerMetrics]], port: int, host: str = "0.0.0.0"):
        outer = self

        class Handler(http.server.BaseHTTPRequestHandler):
            def do_GET(self):  # noqa: N802 - http.server API
                m = outer.metrics_of()
                if self.path.split("?")[0] == "/metrics.json":
                    import json

                    body, ctype = json.dumps(m.snapshot() if m is not None else {}).encode(), "application/json"
                elif self.path.split("?")[0] in ("/metrics", "/"):
                    body, ctype = (m.render_prometheus() if m is not None else "").encode(), "text/plain; version=0.0.4"
                else:
                    self.send_error(404)
                    return
                self.send_response(200)
                self.send_header("Content-Type", ctype)
                self.send_header("Content-Length", str(len(body)))
                self.end_headers()
                self.wfile.write(body)

            def log_message(self, *args):  # scrapes every few seconds must not flood the log
                pass

        self.metrics_of = metrics_of
        self._httpd = http.server.ThreadingHTTPServer((host, port), Handler)
        self._httpd.daemon_threads = True
        self.port = self._httpd.server_address[1]
        self._thread = threading.Thread(target=self._httpd.serve_forever, kwargs=dict(poll_interval=0.2), daemon=True, name="metrics-http")

    def start(self) -> "MetricsServer":
        self._thread.start()
        return self

    def shutdown(self) -> None:
        self._httpd.shutdown()
        self._httpd.server_close()

"""Serving metrics (utils/metrics.py): counters / histograms recorded by the handler, exposed through rpc_info and a
Prometheus endpoint (``--metrics_port``). The reference only logs and publishes ServerInfo."""
import json
import urllib.request

import pytest
import torch

from petals_b200.client.remote_sequential import RemoteSequential
from petals_b200.utils.auto_config import AutoDistributedConfig
from petals_b200.utils.metrics import LATENCY_BUCKETS, ServerMetrics
from tests.utils import checkpoint, swarm_of


def test_metrics_object_and_exposition_format():
    m = ServerMetrics("stageA")
    m.gauge("cache_tokens_left", lambda: 123)
    m.gauge("broken", lambda: 1 / 0)  # a failing gauge is skipped, never raised
    m.observe("inference", 1, 0.0004)
    m.observe("inference", 1, 0.003)
    m.observe("forward", 256, 40.0)  # beyond the last bucket
    m.error("backward")
    m.session_opened(), m.session_opened(), m.session_closed()
    snap = m.snapshot()
    assert snap["requests"] == {"inference": 2, "forward": 1, "backward": 0} and snap["tokens"]["forward"] == 256
    assert snap["errors"]["backward"] == 1 and snap["sessions_active"] == 1 and snap["cache_tokens_left"] == 123.0
    assert snap["mean_latency_ms"]["inference"] == pytest.approx(1.7) and snap["mean_latency_ms"]["backward"] is None and "broken" not in snap
    text = m.render_prometheus()
    assert 'petals_requests_total{peer="stageA",rpc="inference"} 2' in text
    assert f'petals_request_seconds_bucket{{peer="stageA",rpc="inference",le="{LATENCY_BUCKETS[0]}"}} 1' in text
    assert 'petals_request_seconds_bucket{peer="stageA",rpc="inference",le="+Inf"} 2' in text
    assert 'petals_request_seconds_bucket{peer="stageA",rpc="forward",le="30.0"} 0' in text  # cumulative buckets
    assert 'petals_request_seconds_count{peer="stageA",rpc="forward"} 1' in text
    assert 'petals_cache_tokens_left{peer="stageA"} 123.0' in text and "# TYPE petals_request_seconds histogram" in text
    for line in text.splitlines():  # every sample line is `name{labels} number`
        if not line.startswith("#"):
            name_labels, value = line.rsplit(" ", 1)
            float(value)
            assert name_labels.startswith("petals_") and name_labels.endswith("}")


def test_server_records_and_serves_metrics():
    path = checkpoint("llama")
    with swarm_of(path, ["0:4"], metrics_port=0) as (swarm, servers):
        server = servers[0]
        config = AutoDistributedConfig.from_pretrained(path, initial_peers=swarm)
        seq = RemoteSequential(config, dht=swarm)
        x = torch.randn(2, 5, config.hidden_size, requires_grad=True)
        seq(x).sum().backward()  # one rpc_forward + one rpc_backward of 10 tokens each
        with torch.no_grad(), seq.inference_session(max_length=16) as sess:
            sess.step(x[:, :3].detach())
            sess.step(x[:, 3:4].detach())
            info = server.module_container.handler.rpc_info()
            assert info["metrics"]["sessions_active"] == 1
            with pytest.raises(Exception):
                sess._server_sessions[0].stream.step(torch.randn(2, 20, config.hidden_size), metadata={})  # exceeds max_length
        snap = server.module_container.handler.metrics.snapshot()
        assert snap["requests"] == {"inference": 2, "forward": 1, "backward": 1}
        assert snap["tokens"] == {"inference": 8, "forward": 10, "backward": 10}
        assert snap["errors"]["inference"] == 1 and snap["sessions_active"] == 0 and snap["sessions_opened"] == 1
        assert snap["cache_tokens_left"] > 0 and snap["mean_latency_ms"]["inference"] > 0 and snap["queue_size"] == 0
        port = server.metrics_server.port
        text = urllib.request.urlopen(f"http://127.0.0.1:{port}/metrics", timeout=5).read().decode()
        assert 'rpc="inference"} 2' in text and "petals_sessions_active" in text and "petals_cache_tokens_left" in text
        js = json.loads(urllib.request.urlopen(f"http://127.0.0.1:{port}/metrics.json", timeout=5).read())
        assert js["tokens"]["forward"] == 10
        with pytest.raises(Exception):
            urllib.request.urlopen(f"http://127.0.0.1:{port}/nope", timeout=5)

"""HTTP completions front-end (client/api_server.py): plain and streamed completions, multi-turn sessions that keep their
KV caches on the stages, validation errors — against the tokens the Python API generates for the same prompts."""
import json
import urllib.error
import urllib.request

import pytest
import torch

from petals_b200.client.api_server import ApiServer, GenerationService
from petals_b200.utils.auto_config import AutoDistributedModelForCausalLM
from tests.utils import checkpoint, swarm_of


@pytest.fixture(scope="module")
def api():
    path = checkpoint("llama")
    with swarm_of(path, ["0:2", "2:4"]) as (swarm, servers):
        model = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=swarm)
        service = GenerationService(model, None, model_name="tiny-llama", max_session_length=64, session_ttl=0.5)
        server = ApiServer(service, port=0, host="127.0.0.1").start()
        try:
            yield f"http://127.0.0.1:{server.port}", model, service, servers
        finally:
            server.shutdown()


def _post(base, path, body):
    req = urllib.request.Request(base + path, data=json.dumps(body).encode(), headers={"Content-Type": "application/json"})
    with urllib.request.urlopen(req, timeout=60) as resp:
        return resp.status, resp.read().decode()


def _get(base, path):
    with urllib.request.urlopen(base + path, timeout=10) as resp:
        return json.loads(resp.read())


def test_completion_matches_the_python_api(api):
    base, model, service, _ = api
    prompt = [5, 17, 200, 3]
    with torch.inference_mode():
        expected = model.generate(torch.tensor([prompt]), max_new_tokens=6)[0, len(prompt):].tolist()
    status, body = _post(base, "/v1/completions", {"prompt": prompt, "max_tokens": 6})
    out = json.loads(body)
    assert status == 200 and out["object"] == "text_completion" and out["model"] == "tiny-llama"
    choice = out["choices"][0]
    assert choice["token_ids"] == expected and choice["finish_reason"] == "length" and choice["text"] is None  # no tokenizer offline
    assert out["usage"] == {"prompt_tokens": 4, "completion_tokens": 6, "total_tokens": 10}
    # a stop token ends the completion early
    status, body = _post(base, "/v1/completions", {"prompt": prompt, "max_tokens": 6, "stop_token_ids": [expected[2]]})
    choice = json.loads(body)["choices"][0]
    assert choice["finish_reason"] == "stop" and choice["token_ids"] == expected[: expected.index(expected[2]) + 1]
    # seeded sampling is reproducible and differs from greedy at a high temperature
    sample = {"prompt": prompt, "max_tokens": 6, "temperature": 5.0, "top_k": 50, "seed": 7}
    a, b = (json.loads(_post(base, "/v1/completions", sample)[1])["choices"][0]["token_ids"] for _ in range(2))
    assert a == b and a != expected
    assert _get(base, "/health") == {"status": "ok", "open_sessions": 0} and _get(base, "/v1/models")["data"][0]["id"] == "tiny-llama"


def test_streaming_emits_one_event_per_token(api):
    base, model, _, _ = api
    prompt = [9, 8, 7]
    with torch.inference_mode():
        expected = model.generate(torch.tensor([prompt]), max_new_tokens=5)[0, 3:].tolist()
    status, body = _post(base, "/v1/completions", {"prompt": prompt, "max_tokens": 5, "stream": True})
    events = [line[len("data: "):] for line in body.split("\n") if line.startswith("data: ")]
    assert status == 200 and events[-1] == "[DONE]"
    chunks = [json.loads(e) for e in events[:-1]]
    assert [c["choices"][0]["token_id"] for c in chunks[:-1]] == expected and all(c["object"] == "text_completion.chunk" for c in chunks[:-1])
    assert chunks[-1]["object"] == "text_completion" and chunks[-1]["choices"][0]["token_ids"] == expected


def test_conversations_keep_their_kv_caches_between_requests(api):
    base, model, service, servers = api
    first, second = [11, 12, 13], [40, 41]
    with torch.inference_mode():  # the same conversation through the Python API, in one session
        with model.inference_session(max_length=32) as sess:
            turn1 = model.generate(torch.tensor([first]), max_new_tokens=3, session=sess)[0, len(first):].tolist()
            turn2 = model.generate(torch.tensor([second]), max_new_tokens=3, session=sess)[0, -3:].tolist()
    r1 = json.loads(_post(base, "/v1/completions", {"prompt": first, "max_tokens": 3, "session_id": "chat-1"})[1])
    assert r1["choices"][0]["token_ids"] == turn1 and r1["session_id"] == "chat-1" and _get(base, "/health")["open_sessions"] == 1
    tokens_before = sum(s.module_container.handler.metrics.snapshot()["tokens"]["inference"] for s in servers)
    r2 = json.loads(_post(base, "/v1/completions", {"prompt": second, "max_tokens": 3, "session_id": "chat-1"})[1])
    assert r2["choices"][0]["token_ids"] == turn2
    # the second turn only sent its own tokens to the stages (2 prompt + 3 decode steps' worth per stage), not the history again
    tokens_after = sum(s.module_container.handler.metrics.snapshot()["tokens"]["inference"] for s in servers)
    assert tokens_after - tokens_before <= 2 * (len(second) + 3)
    # explicit close, and the idle sweep for forgotten ones
    assert json.loads(_post(base, "/v1/sessions/close", {"session_id": "chat-1"})[1]) == {"closed": True}
    _post(base, "/v1/completions", {"prompt": first, "max_tokens": 1, "session_id": "chat-2"})
    assert service.open_sessions == 1
    import time

    time.sleep(0.7)
    assert service.sweep() == 1 and service.open_sessions == 0


def test_bad_requests_are_reported_not_crashed(api):
    base, *_ = api
    for body in ({"prompt": "text needs a tokenizer"}, {"prompt": [1, 2], "max_tokens": 0}, {"prompt": [10 ** 9]}, {"prompt": []},
                 {"prompt": [1] * 60, "max_tokens": 30}, {"prompt": {"not": "valid"}}):
        with pytest.raises(urllib.error.HTTPError) as err:
            _post(base, "/v1/completions", body)
        assert err.value.code == 400 and "message" in json.loads(err.value.read())["error"]
    with pytest.raises(urllib.error.HTTPError) as err:
        _post(base, "/v1/unknown", {})
    assert err.value.code == 404
    assert json.loads(_post(base, "/v1/completions", {"prompt": [1, 2, 3], "max_tokens": 2})[1])["usage"]["completion_tokens"] == 2  # still serving


def test_concurrent_requests_do_not_interfere(api):
    """Eight clients at once, each with its own prompt: every answer equals the one computed alone (sessions are bound per call)."""
    import concurrent.futures

    base, model, _, _ = api
    prompts = [[3 + i, 50 + 2 * i, 7, 300 + i] for i in range(8)]
    with torch.inference_mode():
        alone = [model.generate(torch.tensor([p]), max_new_tokens=4)[0, len(p):].tolist() for p in prompts]
    with concurrent.futures.ThreadPoolExecutor(8) as pool:
        bodies = list(pool.map(lambda p: _post(base, "/v1/completions", {"prompt": p, "max_tokens": 4, "stream": bool(p[0] % 2)})[1], prompts))
    for body, expected in zip(bodies, alone):
        if body.startswith("data:"):
            final = json.loads([l for l in body.split("\n") if l.startswith("data: ") and l != "data: [DONE]"][-1][len("data: "):])
        else:
            final = json.loads(body)
        assert final["choices"][0]["token_ids"] == expected


def test_stream_reports_a_failure_in_band_and_releases_the_session(api, monkeypatch):
    """Once the event-stream headers are out a failure cannot become an HTTP status any more: it is sent as an error event, the stream
    ends properly and the conversation's lock / session are released."""
    base, model, service, _ = api
    real = model.generate
    calls = {"n": 0}

    def flaky(*args, **kwargs):
        calls["n"] += 1
        if calls["n"] == 3:
            raise RuntimeError("stage lost")
        return real(*args, **kwargs)

    monkeypatch.setattr(model, "generate", flaky)
    status, body = _post(base, "/v1/completions", {"prompt": [4, 5, 6], "max_tokens": 6, "stream": True})
    events = [line[len("data: "):] for line in body.split("\n") if line.startswith("data: ")]
    assert status == 200 and events[-1] == "[DONE]"
    payloads = [json.loads(e) for e in events[:-1]]
    assert len(payloads) == 3 and "error" in payloads[-1] and "stage lost" in payloads[-1]["error"]["message"]
    assert service.open_sessions == 0
    monkeypatch.setattr(model, "generate", real)
    assert json.loads(_post(base, "/v1/completions", {"prompt": [4, 5, 6], "max_tokens": 2})[1])["usage"]["completion_tokens"] == 2
    # the same failure without streaming is an HTTP error
    calls["n"] = 2
    monkeypatch.setattr(model, "generate", flaky)
    with pytest.raises(urllib.error.HTTPError) as err:
        _post(base, "/v1/completions", {"prompt": [4, 5, 6], "max_tokens": 3})
    assert err.value.code == 503

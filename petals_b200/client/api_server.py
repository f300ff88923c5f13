"""A small HTTP front-end for text generation over a swarm (completions API with streaming and multi-turn sessions).

The reference repository stops at the Python client; its public chat service is a separate project that keeps one
``InferenceSession`` per conversation and calls ``generate(..., session=...)`` turn by turn.  This module is that pattern as a
dependency-free server (``http.server``), so that a deployment can be driven by ``curl``:

    POST /v1/completions   {"prompt": "text" | [token ids], "max_tokens": 32, "temperature": 0.8, "top_p": 0.95, "top_k": 40,
                            "stop_token_ids": [2], "stream": false, "session_id": null}
    GET  /v1/models        the served model
    GET  /health           liveness + number of open sessions

* ``stream=true`` answers with server-sent events, one per generated token (each decode step is one ``generate(max_new_tokens=1)``
  inside the same server-side KV session, the pattern of the reference's benchmark_inference.py:44-68).
* ``session_id`` keeps the KV caches of a conversation on the stages between requests: the next request only sends the new tokens.
  Idle sessions are closed after ``session_ttl`` seconds; a request may also ask for ``"close_session": true``.
* Prompts are token ids unless a tokenizer could be loaded for the model directory (no network: nothing is downloaded).
"""
from __future__ import annotations

import http.server
import json
import threading
import time
import uuid
from typing import Any, Dict, Iterator, List, Optional, Sequence

import torch

from petals_b200.utils.logging import get_logger

logger = get_logger(__name__)


class BadRequest(ValueError):
    pass


class _Conversation:
    def __init__(self, session, max_length: int):
        self.session, self.max_length = session, max_length
        self.lock = threading.Lock()
        self.last_used = time.monotonic()
        self.n_tokens = 0


class GenerationService:
    """Owns the client model, the optional tokenizer and the table of live conversations."""

    def __init__(self, model, tokenizer=None, *, model_name: str = "model", max_session_length: int = 2048, session_ttl: float = 300.0):
        self.model, self.tokenizer, self.model_name = model, tokenizer, model_name
        self.max_session_length, self.session_ttl = max_session_length, session_ttl
        self._conversations: Dict[str, _Conversation] = {}
        self._lock = threading.Lock()

    # ---- sessions ----------------------------------------------------------------------------------------------------------
    def _open(self, session_id: Optional[str], needed: int) -> tuple:
        """-> (conversation, is_persistent). A persistent conversation is created on first use of its id."""
        self.sweep()
        if session_id is None:
            length = min(self.max_session_length, needed)
            return _Conversation(self._new_session(length), length), False
        with self._lock:
            conv = self._conversations.get(session_id)
            if conv is None:
                conv = _Conversation(self._new_session(self.max_session_length), self.max_session_length)
                self._conversations[session_id] = conv
        return conv, True

    def _new_session(self, max_length: int):
        """A session that is not bound to any thread's context: every ``generate(session=...)`` call binds it for its own duration."""
        from petals_b200.client.inference_session import InferenceSession

        return InferenceSession(self.model.layers.sequence_manager, max_length).__enter__()

    def close_session(self, session_id: str) -> bool:
        with self._lock:
            conv = self._conversations.pop(session_id, None)
        if conv is None:
            return False
        with conv.lock:
            conv.session.close()
        return True

    def sweep(self) -> int:
        now = time.monotonic()
        with self._lock:
            stale = [sid for sid, c in self._conversations.items() if now - c.last_used > self.session_ttl and not c.lock.locked()]
        return sum(self.close_session(sid) for sid in stale)

    @property
    def open_sessions(self) -> int:
        with self._lock:
            return len(self._conversations)

    def shutdown(self) -> None:
        for sid in list(self._conversations):
            self.close_session(sid)

    # ---- text <-> ids ------------------------------------------------------------------------------------------------------------
    def encode(self, prompt: Any) -> List[int]:
        if isinstance(prompt, str):
            if self.tokenizer is None:
                raise BadRequest("no tokenizer is available for this model: send `prompt` as a list of token ids")
            return list(self.tokenizer(prompt, add_special_tokens=True)["input_ids"])
        if isinstance(prompt, (list, tuple)) and all(isinstance(t, int) and not isinstance(t, bool) for t in prompt):
            vocab = self.model.config.vocab_size
            if any(not 0 <= t < vocab for t in prompt):
                raise BadRequest(f"token ids must be within [0, {vocab})")
            return list(prompt)
        raise BadRequest("`prompt` must be a string or a list of token ids")

    def decode(self, ids: Sequence[int]) -> Optional[str]:
        return None if self.tokenizer is None else self.tokenizer.decode(list(ids), skip_special_tokens=True)

    # ---- generation ----------------------------------------------------------------------------------------------------------------
    def stream(self, request: Dict[str, Any]) -> Iterator[Dict[str, Any]]:
        """Yields one event per generated token and a final event with ``finish_reason`` and ``usage``."""
        prompt = self.encode(request.get("prompt", []))
        max_tokens = int(request.get("max_tokens", 16))
        if max_tokens < 1:
            raise BadRequest("`max_tokens` must be >= 1")
        if not prompt and request.get("session_id") is None:
            raise BadRequest("a new conversation needs a non-empty prompt")
        temperature = float(request.get("temperature", 0.0))
        sampling = dict(do_sample=temperature > 0, temperature=max(temperature, 1e-5), top_k=request.get("top_k"), top_p=request.get("top_p"),
                        repetition_penalty=float(request.get("repetition_penalty", 1.0)))
        stop_ids = set(int(t) for t in (request.get("stop_token_ids") or []))
        seed = request.get("seed")
        generator = None if seed is None else torch.Generator(device=self.model.device).manual_seed(int(seed))
        session_id = request.get("session_id")
        conv, persistent = self._open(session_id, len(prompt) + max_tokens + 1)
        produced: List[int] = []
        finish = "length"
        try:
            with conv.lock, torch.inference_mode():
                if conv.n_tokens + len(prompt) + max_tokens + 1 > conv.max_length:
                    raise BadRequest(f"conversation would exceed the session length of {conv.max_length} tokens")
                feed = torch.tensor([prompt], dtype=torch.int64, device=self.model.device) if prompt else None
                for _ in range(max_tokens):
                    out = self.model.generate(feed, max_new_tokens=1, session=conv.session, generator=generator, **sampling)
                    feed = None  # later steps continue from the session's own history
                    token = int(out[0, -1])
                    produced.append(token)
                    yield {"token_id": token, "text": self.decode([token])}
                    if token in stop_ids:
                        finish = "stop"
                        break
                conv.n_tokens += len(prompt) + len(produced)
                conv.last_used = time.monotonic()
        finally:
            if not persistent:
                conv.session.close()
            elif request.get("close_session"):
                self.close_session(session_id)
        yield {"finish_reason": finish, "token_ids": produced, "text": self.decode(produced),
               "usage": {"prompt_tokens": len(prompt), "completion_tokens": len(produced), "total_tokens": len(prompt) + len(produced)}}

    def complete(self, request: Dict[str, Any]) -> Dict[str, Any]:
        final = None
        for event in self.stream(request):
            final = event
        return self._envelope(final, request)

    def _envelope(self, final: Dict[str, Any], request: Dict[str, Any]) -> Dict[str, Any]:
        return {"id": "cmpl-" + uuid.uuid4().hex[:24], "object": "text_completion", "created": int(time.time()), "model": self.model_name,
                "session_id": request.get("session_id"),
                "choices": [{"index": 0, "text": final["text"], "token_ids": final["token_ids"], "finish_reason": final["finish_reason"]}],
                "usage": final["usage"]}


class ApiServer:
    """``ApiServer(service, port).start()``; ``.port`` is the bound port (0 = pick one)."""

    def __init__(self, service: GenerationService, port: int = 8000, host: str = "0.0.0.0"):
        self.service = service
        outer = self

        class Handler(http.server.BaseHTTPRequestHandler):
            protocol_version = "HTTP/1.1"

            def _json(self, status: int, body: Dict[str, Any]) -> None:
                raw = json.dumps(body).encode()
                self.send_response(status)
                self.send_header("Content-Type", "application/json")
                self.send_header("Content-Length", str(len(raw)))
                self.end_headers()
                self.wfile.write(raw)

            def do_GET(self):  # noqa: N802 - http.server API
                path = self.path.split("?")[0]
                if path == "/health":
                    self._json(200, {"status": "ok", "open_sessions": outer.service.open_sessions})
                elif path == "/v1/models":
                    self._json(200, {"object": "list", "data": [{"id": outer.service.model_name, "object": "model"}]})
                else:
                    self._json(404, {"error": {"message": f"unknown path {path}"}})

            def do_POST(self):  # noqa: N802
                path = self.path.split("?")[0]
                try:
                    length = int(self.headers.get("Content-Length") or 0)
                    request = json.loads(self.rfile.read(length) or b"{}")
                    if not isinstance(request, dict):
                        raise BadRequest("the request body must be a JSON object")
                    if path == "/v1/completions":
                        if request.get("stream"):
                            return self._stream(request)
                        return self._json(200, outer.service.complete(request))
                    if path == "/v1/sessions/close":
                        return self._json(200, {"closed": outer.service.close_session(str(request.get("session_id")))})
                    self._json(404, {"error": {"message": f"unknown path {path}"}})
                except (BadRequest, json.JSONDecodeError, ValueError) as e:
                    self._json(400, {"error": {"message": str(e), "type": "invalid_request_error"}})
                except Exception as e:  # noqa: BLE001 - the swarm may be degraded; tell the caller, keep serving
                    logger.warning(f"request failed: {e!r}")
                    self._json(503, {"error": {"message": repr(e), "type": "server_error"}})

            def _stream(self, request: Dict[str, Any]) -> None:
                events = outer.service.stream(request)
                first = next(events)  # errors of validation surface before the headers are sent
                self.send_response(200)
                self.send_header("Content-Type", "text/event-stream")
                self.send_header("Cache-Control", "no-cache")
                self.send_header("Connection", "close")
                self.end_headers()
                self.close_connection = True

                def emit(payload: Dict[str, Any]) -> None:
                    self.wfile.write(b"data: " + json.dumps(payload).encode() + b"\n\n")
                    self.wfile.flush()

                event = first
                try:
                    while True:
                        if "finish_reason" in event:
                            emit(outer.service._envelope(event, request))
                            break
                        emit({"object": "text_completion.chunk", "choices": [{"index": 0, "text": event["text"], "token_id": event["token_id"]}]})
                        event = next(events)
                except (BrokenPipeError, ConnectionError):
                    events.close()  # the client went away: stop generating (the generator's finally releases the session)
                    return
                except Exception as e:  # noqa: BLE001 - the headers are out: report in-band, then end the stream
                    logger.warning(f"stream failed: {e!r}")
                    emit({"error": {"message": repr(e), "type": "server_error"}})
                self.wfile.write(b"data: [DONE]\n\n")
                self.wfile.flush()

            def log_message(self, fmt, *args):
                logger.debug("api: " + fmt % args)

        self._httpd = http.server.ThreadingHTTPServer((host, port), Handler)
        self._httpd.daemon_threads = True
        self.port = self._httpd.server_address[1]
        self._thread = threading.Thread(target=self._httpd.serve_forever, kwargs=dict(poll_interval=0.2), daemon=True, name="api-http")

    def start(self) -> "ApiServer":
        self._thread.start()
        return self

    def serve_forever(self) -> None:
        self._httpd.serve_forever(poll_interval=0.2)

    def shutdown(self) -> None:
        self._httpd.shutdown()
        self._httpd.server_close()
        self.service.shutdown()


def load_tokenizer(model_path: str):
    """The model directory's tokenizer if it ships one (nothing is fetched); None otherwise."""
    import os

    if not any(os.path.exists(os.path.join(model_path, f)) for f in ("tokenizer.json", "tokenizer.model", "tokenizer_config.json")):
        return None
    try:
        from transformers import AutoTokenizer

        return AutoTokenizer.from_pretrained(model_path, local_files_only=True)
    except Exception as e:  # noqa: BLE001
        logger.warning(f"could not load a tokenizer from {model_path}: {e}")
        return None

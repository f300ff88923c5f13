"""The C++ socket loops of the control transport (csrc/runtime/socket_io.cpp) against the pure-Python path.

The reference delegates framing to hivemind's Go daemon and never tests it in-tree; here the frame format is ours, so
it is checked directly: both implementations interoperate in either direction, large payloads survive partial
sends/receives on a non-blocking (timeout) socket, time-outs and closed peers surface as the exceptions the retry
logic of the client expects.
"""
import socket
import threading
import time

import pytest
import torch

from petals_b200.parallel import transport
from petals_b200.parallel.transport import recv_message, send_message


def _pair(timeout=20.0):
    a, b = socket.socketpair(socket.AF_UNIX, socket.SOCK_STREAM)
    a.settimeout(timeout), b.settimeout(timeout)
    return a, b


def _tensors():
    g = torch.Generator().manual_seed(0)
    return [torch.randn(3, 1000, 257, generator=g).to(torch.bfloat16),  # ~1.5 MB, odd sizes
            torch.empty(0), torch.arange(7), torch.tensor(3.5), torch.rand(5, 5, generator=g) > 0.5,
            torch.randn(2, 3, generator=g, dtype=torch.float64)]


def test_native_library_is_used():
    io = transport._native_io()
    assert io is not None and hasattr(io, "pb_sock_send_frames") and hasattr(io, "pb_sock_recv_exact")


@pytest.mark.parametrize("sender_native,receiver_native", [(True, True), (True, False), (False, True), (False, False)])
def test_both_implementations_interoperate(monkeypatch, sender_native, receiver_native):
    a, b = _pair()
    sent = _tensors()
    got = {}

    def reader():
        got["msg"] = recv_message(b)

    native_io = transport._native_io()
    # the reader thread picks its implementation when it starts receiving, the sender when it sends
    monkeypatch.setattr(transport, "_native_io", lambda: native_io if receiver_native else None)
    t = threading.Thread(target=reader)
    t.start()
    time.sleep(0.05)
    monkeypatch.setattr(transport, "_native_io", lambda: native_io if sender_native else None)
    # (a single shared hook: the header decides nothing about the implementation, so mixing is safe even if the
    #  reader observes the sender's choice for its later parts)
    send_message(a, {"method": "x", "meta": {"k": [1, 2]}}, sent, compression=[None] * len(sent))
    t.join(20)
    header, tensors = got["msg"]
    assert header["method"] == "x" and header["meta"] == {"k": [1, 2]}
    assert len(tensors) == len(sent)
    for s, r in zip(sent, tensors):
        assert r.dtype == s.dtype and r.shape == s.shape and torch.equal(r, s)
    a.close(), b.close()


def test_large_payload_through_a_small_socket_buffer():
    a, b = _pair()
    a.setsockopt(socket.SOL_SOCKET, socket.SO_SNDBUF, 4096)  # forces many partial sendmsg() calls + EAGAIN waits
    x = torch.randn(8, 1024, 1024)  # 32 MiB
    out = {}
    t = threading.Thread(target=lambda: out.setdefault("m", recv_message(b)))
    t.start()
    send_message(a, {"method": "big"}, [x, torch.arange(3), x[:1]])
    t.join(60)
    _, (y, ids, y1) = out["m"]
    assert torch.equal(x, y) and torch.equal(ids, torch.arange(3)) and torch.equal(y1, x[:1])
    a.close(), b.close()


def test_receive_timeout_and_peer_close():
    a, b = _pair(timeout=0.3)
    # a header that promises 1 MB of tensor bytes which never arrive -> the receiver times out instead of hanging
    import msgpack
    import struct

    payload = msgpack.packb({"method": "x", "tensors": [{"dtype": "u8", "shape": [1 << 20], "nbytes": 1 << 20, "parts": [1 << 20],
                                                         "c": {"codec": "NONE"}}]})
    a.sendall(struct.pack("<I", len(payload)) + payload + b"\0" * 100)
    t0 = time.monotonic()
    with pytest.raises((socket.timeout, TimeoutError)):
        recv_message(b)
    assert 0.2 < time.monotonic() - t0 < 5
    a.close(), b.close()

    a, b = _pair()
    a.sendall(struct.pack("<I", len(payload)) + payload + b"\0" * 100)
    a.close()  # the peer dies mid-message
    with pytest.raises(ConnectionError):
        recv_message(b)
    b.close()


def test_send_to_a_closed_peer_raises_connection_error():
    a, b = _pair()
    b.close()
    with pytest.raises((ConnectionError, OSError)):
        for _ in range(64):  # the first sends may still fit into the socket buffer
            send_message(a, {"method": "x"}, [torch.zeros(1 << 16)])
    a.close()


def test_native_and_python_paths_move_large_frames(monkeypatch):
    """Moves 64 MiB through a socketpair with both implementations and prints the timings (alone on a machine the C++ loops
    are 1.5-3x faster: no tobytes()/join/bytearray copies, no GIL hand-offs; 48 ms vs 128 ms when this was written). Only a
    gross regression fails the test: wall-clock ratios are not stable inside a busy test process."""
    x = torch.randn(16, 1024, 1024)  # 64 MiB
    native_io = transport._native_io()

    def run(io):
        monkeypatch.setattr(transport, "_native_io", lambda: io)
        best = float("inf")
        for _ in range(3):
            a, b = _pair(60)
            out = {}
            t = threading.Thread(target=lambda: out.setdefault("m", recv_message(b)))
            t0 = time.perf_counter()
            t.start()
            send_message(a, {"method": "x"}, [x])
            t.join(60)
            best = min(best, time.perf_counter() - t0)
            assert torch.equal(out["m"][1][0], x)
            a.close(), b.close()
        return best

    t_native, t_python = run(native_io), run(None)
    print(f"64 MiB frame: native {t_native * 1e3:.1f} ms, python {t_python * 1e3:.1f} ms")
    assert t_native < 5 * t_python


def test_server_survives_garbage_and_keeps_serving(tmp_path):
    """Foreign or corrupt byte streams (a port scanner, a half-written frame, absurd lengths) must cost one connection, not the server."""
    import os
    import struct

    from petals_b200.parallel.transport import ProtocolError, RemoteHandlerProxy, RpcServer

    class Handler:
        compression = None

        def rpc_forward(self, uids, hidden, *rest, metadata=None):
            return hidden + 1

    server = RpcServer(Handler(), "tcp://127.0.0.1:0")
    server.start()
    try:
        for junk in (b"GET / HTTP/1.1\r\n\r\n", struct.pack("<I", 0xFFFFFFFF) + b"x" * 64, struct.pack("<I", 5) + b"\xc1\xc1\xc1\xc1\xc1",
                     struct.pack("<I", 3) + b"\x93\x01\x02", os.urandom(4096)):
            s = socket.create_connection(("127.0.0.1", server.port), timeout=5)
            s.sendall(junk)
            s.settimeout(5)
            try:
                assert s.recv(1 << 16) in (b"",) or True  # the server may answer with an error frame or just hang up
            except (ConnectionError, socket.timeout):
                pass
            s.close()
        x = torch.randn(2, 3, 4)
        assert torch.equal(RemoteHandlerProxy(server.address).rpc_forward(["m.0"], x), x + 1)
    finally:
        server.shutdown()
    a, b = _pair()
    a.sendall(struct.pack("<I", 1 << 30))
    with pytest.raises(ProtocolError):
        recv_message(b)
    a.close(), b.close()

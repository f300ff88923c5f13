"""Swarms that span several boxes: a TCP registry (``run_dht --host_maddrs``) + TCP peer endpoints.

The reference CI bootstraps from ``run_dht --host_maddrs /ip4/127.0.0.1/tcp/31337`` and hands the multiaddr to four
server processes and the tests (.github/workflows/run-tests.yaml:52-91); this is the same harness on loopback TCP."""
import os
import re
import signal
import subprocess
import sys
import time

import pytest
import torch

from petals_b200.parallel.registry import RegistryServer, TcpSwarm
from petals_b200.parallel.swarm import Swarm, get_dht_time, resolve_swarm
from petals_b200.server.handler import TransformerConnectionHandler
from petals_b200.server.reachability import ReachabilityProtocol, check_direct_reachability, validate_direct_reachability
from petals_b200.parallel.transport import to_multiaddr
from petals_b200.utils.auto_config import AutoDistributedConfig, AutoDistributedModelForCausalLM
from tests.utils import checkpoint, local_blocks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Echo:
    compression = None

    def rpc_forward(self, uids, hidden, *rest, metadata=None):
        return hidden * 2

    def rpc_info(self, uids=None):
        return {"who": "echo"}


def test_registry_data_model_and_endpoints():
    registry = RegistryServer("tcp://127.0.0.1:0").start()
    try:
        a, b = TcpSwarm(registry.address, bind_host="127.0.0.1"), TcpSwarm(to_multiaddr(registry.address) + "/p2p/QmBootstrap")
        assert b.registry_address == a.registry_address  # both spellings of the address mean the same registry
        now = get_dht_time()
        assert a.store("m.0", "peerA", [1, {"x": 2.5}, None], now + 30)
        assert a.store("m.0", "peerB", [2], now - 1)  # already expired
        assert a.store("m.1", "peerA", [3], now + 30)
        assert b.get("m.0") == {"peerA": ([1, {"x": 2.5}, None], pytest.approx(now + 30))}
        many = b.get_many(["m.0", "m.1", "m.2"])
        assert set(many["m.0"]) == {"peerA"} and many["m.1"]["peerA"][0] == [3] and many["m.2"] == {}
        # endpoints: served over TCP by the process that registered them, found through the registry by the others
        a.register_endpoint("peerA", _Echo())
        assert b.peers() == ["peerA"]
        stub = b.connect("peerA")
        x = torch.randn(2, 3, 8)
        assert torch.equal(stub.rpc_forward(["m.0"], x), x * 2) and stub.rpc_info() == {"who": "echo"}
        assert a.connect("peerA").__class__ is _Echo  # same process: no socket at all
        a.unregister_endpoint("peerA")
        assert b.peers() == [] and b.get("m.0") == {} and b.get("m.1") == {}  # its records leave with it
        b.forget("peerA")
        with pytest.raises(ConnectionError):
            b.connect("peerA")
        a.close(), b.close()
    finally:
        registry.shutdown()


def test_registry_restart_is_survivable():
    registry = RegistryServer("tcp://127.0.0.1:0").start()
    port = int(registry.address.rsplit(":", 1)[1])
    swarm = TcpSwarm(registry.address, bind_host="127.0.0.1")
    swarm.register_endpoint("p", _Echo())
    swarm.store("k", "p", [1], get_dht_time() + 30)
    registry.shutdown()
    with pytest.raises(ConnectionError):
        swarm.get("k")
    registry = RegistryServer(f"tcp://127.0.0.1:{port}").start()  # comes back empty on the same address
    try:
        assert swarm.get("k") == {}
        swarm.store("k", "p", [1], get_dht_time() + 30)  # what the announcer does every update_period ...
        swarm.refresh_endpoint("p")  # ... together with its address
        assert set(swarm.get("k")) == {"p"} and swarm.peers() == ["p"]
    finally:
        swarm.close()
        registry.shutdown()


class _Helper:
    """The two RPCs of a stage handler that reachability needs."""
    compression = None

    def __init__(self, swarm):
        self.swarm = swarm

    rpc_check = TransformerConnectionHandler.rpc_check

    def rpc_ping(self):
        return None


def test_peer_assisted_reachability_check():
    registry = RegistryServer("tcp://127.0.0.1:0").start()
    a, b = TcpSwarm(registry.address, bind_host="127.0.0.1"), TcpSwarm(registry.address, bind_host="127.0.0.1")
    try:
        a.register_endpoint("A", _Helper(a))
        assert check_direct_reachability(a, "A") is None  # the first server of a swarm has nobody to ask
        b.register_endpoint("B", _Helper(b))
        assert check_direct_reachability(b, "B") is True  # A dials B's announced address
        validate_direct_reachability(b, "B")
        assert ReachabilityProtocol(b).call_check("A", check_peer="B") is True
        # B announces an address nobody can connect to (wrong --public_ip, firewall, ...): A reports it, B refuses to serve
        b._call("reg_register", peer_id="B", address="tcp://127.0.0.1:1", ttl=60)
        assert check_direct_reachability(b, "B", wait_timeout=1.0) is False
        with pytest.raises(RuntimeError, match="--public_ip"):
            validate_direct_reachability(b, "B", wait_timeout=1.0)
        # swarms inside one box have nothing to check
        assert check_direct_reachability(Swarm("inproc-reach"), "x") is True and check_direct_reachability() is True
    finally:
        a.close(), b.close()
        registry.shutdown()


def test_resolve_swarm_recognises_network_addresses():
    registry = RegistryServer("tcp://127.0.0.1:0").start()
    try:
        s1 = resolve_swarm([to_multiaddr(registry.address)])
        s2 = resolve_swarm(registry.address)
        assert isinstance(s1, TcpSwarm) and s1 is s2
        s1.close()
    finally:
        registry.shutdown()


def _spawn(args, log):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), PETALS_LOGLEVEL="INFO")
    return subprocess.Popen([sys.executable, "-m", *args], stdout=log, stderr=subprocess.STDOUT, env=env, cwd=ROOT)


def test_run_dht_registry_and_two_server_processes_over_tcp(tmp_path):
    path = checkpoint("llama")
    dht_log = open(tmp_path / "dht.log", "w+")
    dht = _spawn(["petals.cli.run_dht", "--host_maddrs", "/ip4/127.0.0.1/tcp/0", "--identity_path", "tests/bootstrap.id"], dht_log)
    procs, logs = [dht], [dht_log]
    try:
        deadline, maddr = time.monotonic() + 60, None
        while maddr is None:
            assert time.monotonic() < deadline and dht.poll() is None, open(tmp_path / "dht.log").read()
            m = re.search(r"--initial_peers (/ip4/127\.0\.0\.1/tcp/\d+)", open(tmp_path / "dht.log").read())
            maddr = m.group(1) if m else None
            time.sleep(0.1)
        common = ["--initial_peers", maddr, "--torch_dtype", "float32", "--device", "cpu", "--throughput", "1", "--update_period", "1",
                  "--host_maddrs", "/ip4/127.0.0.1/tcp/0"]
        for i, span in enumerate(["0:2", "2:4"]):
            logs.append(open(tmp_path / f"server{i}.log", "w"))
            procs.append(_spawn(["petals.cli.run_server", path, "--block_indices", span, "--peer_id", f"stage{i}", *common], logs[-1]))
        model = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=[maddr], max_retries=150, min_backoff=0.5, max_backoff=1.0)
        config = AutoDistributedConfig.from_pretrained(path)
        ids = torch.randint(0, config.vocab_size, (1, 6), generator=torch.Generator().manual_seed(0))
        with torch.inference_mode():
            logits = model(ids).logits
            assert all(p.poll() is None for p in procs), "a process died"
            h = model.model.embed(ids)
            for b in local_blocks(path, config.num_hidden_layers):
                h = b(h)[0]
            assert torch.allclose(logits, model.lm_head(model.model.final_norm(h)), atol=1e-3)
            out = model.generate(ids, max_new_tokens=3)
        assert out.shape == (1, 9)
        x = torch.randn(2, 3, config.hidden_size, requires_grad=True)
        model.model.layers(x).sum().backward()
        assert torch.isfinite(x.grad).all()
        assert sorted(model.model.layers.sequence_manager.dht.peers()) == ["stage0", "stage1"]
    finally:
        for p in procs[::-1]:
            p.send_signal(signal.SIGTERM)
        for p in procs:
            try:
                p.wait(timeout=10)
            except subprocess.TimeoutExpired:
                p.kill()
        for f in logs:
            f.close()


def test_failover_when_a_server_process_is_killed_mid_session(tmp_path):
    """Blocks 2:4 are served by two processes; the one an inference session is using is SIGKILLed between two steps. The client
    must notice the dead TCP stream, ban the peer, re-route through the replica and rebuild its KV cache by replaying the
    history — the outputs stay exact (reference fault model: client/inference_session.py:364-391; its CI never kills a server)."""
    path = checkpoint("llama")
    registry = RegistryServer("tcp://127.0.0.1:0").start()
    maddr = to_multiaddr(registry.address)
    common = ["--initial_peers", maddr, "--torch_dtype", "float32", "--device", "cpu", "--throughput", "1", "--update_period", "1",
              "--host_maddrs", "/ip4/127.0.0.1/tcp/0", "--skip_reachability_check"]
    spans = {"head": "0:2", "tail-a": "2:4", "tail-b": "2:4"}
    procs, logs = {}, []
    try:
        for name, span in spans.items():
            logs.append(open(tmp_path / f"{name}.log", "w"))
            procs[name] = _spawn(["petals.cli.run_server", path, "--block_indices", span, "--peer_id", name, *common], logs[-1])
        model = AutoDistributedModelForCausalLM.from_pretrained(path, initial_peers=[maddr], max_retries=150, min_backoff=0.3, max_backoff=1.0,
                                                                update_period=1.0)
        config = AutoDistributedConfig.from_pretrained(path)
        manager = model.model.layers.sequence_manager
        deadline = time.monotonic() + 90
        while sorted(manager.dht.peers()) != sorted(spans):  # all three have announced their endpoints
            assert time.monotonic() < deadline and all(p.poll() is None for p in procs.values())
            time.sleep(0.2)
        manager.update()
        x = torch.randn(1, 6, config.hidden_size, generator=torch.Generator().manual_seed(0))
        with torch.inference_mode():
            expected = x
            for b in local_blocks(path, config.num_hidden_layers):
                expected = b(expected)[0]
            with model.model.layers.inference_session(max_length=8) as session:
                first = session.step(x[:, :4])
                used = [s.span.peer_id for s in session._server_sessions]
                assert used[0] == "head" and used[1] in ("tail-a", "tail-b")
                procs[used[1]].kill()  # no goodbye: the stream just dies
                procs[used[1]].wait(timeout=10)
                rest = torch.cat([session.step(x[:, 4:5]), session.step(x[:, 5:6])], dim=1)
                survivor = [s.span.peer_id for s in session._server_sessions][1]
        assert survivor != used[1] and survivor in ("tail-a", "tail-b")
        assert torch.allclose(torch.cat([first, rest], dim=1), expected, atol=1e-4)
    finally:
        for p in procs.values():
            if p.poll() is None:
                p.send_signal(signal.SIGTERM)
        for p in procs.values():
            try:
                p.wait(timeout=10)
            except subprocess.TimeoutExpired:
                p.kill()
        for f in logs:
            f.close()
        registry.shutdown()


def test_client_side_wire_codecs_over_tcp():
    """`wire_compression` / `output_compression` in the client config: lossy but close results through socket peers, exact without."""
    from petals_b200.client.remote_sequential import RemoteSequential
    from petals_b200.server.server import Server

    path = checkpoint("llama")
    registry = RegistryServer("tcp://127.0.0.1:0").start()
    serving = TcpSwarm(registry.address, bind_host="127.0.0.1")
    server = Server(initial_peers=serving, converted_model_name_or_path=path, block_indices="0:4", torch_dtype="float32", device="cpu", throughput=1.0,
                    update_period=0.5, skip_reachability_check=True)
    server.run_in_background(timeout=120)
    client_swarm = TcpSwarm(registry.address, bind_host="127.0.0.1")  # a second handle = "another process": forces the socket path
    try:
        x = torch.randn(2, 5, AutoDistributedConfig.from_pretrained(path).hidden_size)
        results = {}
        for name, kw in (("exact", {}), ("lossy", dict(wire_compression="FLOAT16", output_compression="BLOCKWISE_8BIT"))):
            config = AutoDistributedConfig.from_pretrained(path, initial_peers=[registry.address], **kw)
            seq = RemoteSequential(config, dht=client_swarm)
            with torch.no_grad():
                fwd = seq(x)
                with seq.inference_session(max_length=8) as sess:
                    inf = torch.cat([sess.step(x[:, :3]), sess.step(x[:, 3:])], dim=1)
            results[name] = (fwd, inf)
            seq.sequence_manager.shutdown()
        exact_f, exact_i = results["exact"]
        lossy_f, lossy_i = results["lossy"]
        assert torch.allclose(exact_f, exact_i, atol=1e-4)
        for lossy, exact in ((lossy_f, exact_f), (lossy_i, exact_i)):
            assert not torch.equal(lossy, exact) and torch.allclose(lossy, exact, atol=0.25, rtol=0.05)
    finally:
        server.shutdown()
        client_swarm.close(), serving.close()
        registry.shutdown()

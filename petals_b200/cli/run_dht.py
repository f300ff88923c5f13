"""Create (and keep alive) a rendezvous location for a multi-process swarm
(reference: src/petals/cli/run_dht.py:37-102 starts a bootstrap DHT peer and prints its multiaddrs).

    python -m petals.cli.run_dht --rendezvous /dev/shm/petals-swarm          # one box: a shared directory
    python -m petals.cli.run_dht --host_maddrs /ip4/0.0.0.0/tcp/31337        # several boxes: a TCP registry

prints the value to pass as ``--initial_peers`` to ``run_server`` and as ``initial_peers=[...]`` to clients."""
from __future__ import annotations

import argparse
import os
import signal
import tempfile
import time

from petals_b200.parallel.swarm import FileSwarm
from petals_b200.utils.logging import get_logger

logger = get_logger(__name__)


def main(argv=None) -> None:
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--rendezvous", default=None, help="directory shared by all stage processes and clients of this swarm (one box)")
    parser.add_argument("--host_maddrs", nargs="+", default=None,
                        help="listen address of a network registry for swarms that span several boxes, e.g. /ip4/0.0.0.0/tcp/31337 or "
                             "tcp://0.0.0.0:31337 (a plain path here is taken as --rendezvous)")
    parser.add_argument("--announce_maddrs", nargs="+", default=None, help="address to print for peers when listening on 0.0.0.0")
    parser.add_argument("--initial_peers", nargs="*", default=None,
                        help="accepted for compatibility: a registry is a single process, it does not join other bootstrap peers")
    for flag in ("--no_relay", "--use_auto_relay", "--use_ipfs"):
        parser.add_argument(flag, action="store_true", help="accepted for compatibility (no NAT traversal / IPFS bootstrap on a private network)")
    parser.add_argument("--identity_path", default=None, help="accepted for compatibility (peers are named by their GPU rank)")
    parser.add_argument("--refresh_period", type=float, default=30.0, help="how often to report swarm membership")
    parser.add_argument("--once", action="store_true", help="create the rendezvous and exit (for scripts)")
    args, _unknown = parser.parse_known_args(argv)
    from petals_b200.parallel.transport import is_network_address, parse_address, to_multiaddr

    if args.host_maddrs and not is_network_address(args.host_maddrs[0]) and args.rendezvous is None:
        args.rendezvous, args.host_maddrs = args.host_maddrs[0], None
    if args.host_maddrs:
        return _run_registry(args, parse_address(args.host_maddrs[0]), to_multiaddr)
    path = args.rendezvous or os.path.join(tempfile.gettempdir(), f"petals-swarm-{os.getpid()}")
    swarm = FileSwarm(path)
    print(f"Running a swarm rendezvous at {swarm.address}", flush=True)
    print(f"To connect stages or clients, pass --initial_peers {swarm.address}", flush=True)
    if args.once:
        return
    stop = []
    signal.signal(signal.SIGTERM, lambda *_: stop.append(1))
    try:
        while not stop:
            time.sleep(args.refresh_period)
            logger.info(f"peers alive: {sorted(swarm.peers())}")
    except KeyboardInterrupt:
        pass


def _run_registry(args, listen, to_multiaddr) -> None:
    import socket

    from petals_b200.parallel.registry import RegistryServer
    from petals_b200.parallel.transport import format_address

    _, host, port = listen
    registry = RegistryServer(format_address(host, port)).start()
    bound_port = parse_port(registry.address)
    if args.announce_maddrs:
        shown = [a if a.startswith("/") else to_multiaddr(a) for a in args.announce_maddrs]
    elif host in ("0.0.0.0", "::", ""):
        hosts = ["127.0.0.1"]
        try:
            ip = socket.gethostbyname(socket.gethostname())
            if ip not in hosts:
                hosts.append(ip)
        except OSError:
            pass
        shown = [to_multiaddr(format_address(h, bound_port)) for h in hosts]
    else:
        shown = [to_multiaddr(format_address(host, bound_port))]
    print(f"Running a swarm registry, listening on {registry.address}", flush=True)
    for a in shown:
        print(f"To connect stages or clients, pass --initial_peers {a}", flush=True)
    if args.once:
        registry.shutdown()
        return
    stop = []
    signal.signal(signal.SIGTERM, lambda *_: stop.append(1))
    try:
        waited = 0.0
        while not stop:
            time.sleep(0.2)
            waited += 0.2
            if waited >= args.refresh_period:
                waited = 0.0
                logger.info(f"peers alive: {registry.peers()}")
    except KeyboardInterrupt:
        pass
    finally:
        registry.shutdown()


def parse_port(address: str) -> int:
    return int(address.rsplit(":", 1)[1])


if __name__ == "__main__":
    main()

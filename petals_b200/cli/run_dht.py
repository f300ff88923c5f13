"""Create (and keep alive) a rendezvous location for a multi-process swarm
(reference: src/petals/cli/run_dht.py:37-102 starts a bootstrap DHT peer and prints its multiaddrs).

    python -m petals.cli.run_dht --rendezvous /dev/shm/petals-swarm

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
    parser.add_argument("--rendezvous", "--host_maddrs", dest="rendezvous", default=None,
                        help="directory shared by all stage processes and clients of this swarm")
    parser.add_argument("--identity_path", default=None, help="accepted for compatibility (peers are named by their GPU rank)")
    parser.add_argument("--refresh_period", type=float, default=30.0, help="how often to report swarm membership")
    parser.add_argument("--once", action="store_true", help="create the rendezvous and exit (for scripts)")
    args, _unknown = parser.parse_known_args(argv)
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


if __name__ == "__main__":
    main()

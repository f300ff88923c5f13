"""Launcher: ``python -m petals.cli.run_dht`` (implementation in :mod:`petals_b200.cli.run_dht`)."""
from petals_b200.cli.run_dht import *  # noqa: F401,F403
from petals_b200.cli.run_dht import main

if __name__ == "__main__":
    main()

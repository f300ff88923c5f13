"""Launcher: ``python -m petals.cli.run_server`` (implementation in :mod:`petals_b200.cli.run_server`)."""
from petals_b200.cli.run_server import *  # noqa: F401,F403
from petals_b200.cli.run_server import main

if __name__ == "__main__":
    main()

"""``python -m petals.cli.run_server`` / ``run_dht`` entry points (thin launchers over petals_b200.cli)."""

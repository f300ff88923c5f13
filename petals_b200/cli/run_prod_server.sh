#!/bin/bash
# Restart loop for production stages (reference: src/petals/cli/run_prod_server.sh:5-9).
export PETALS_LOGLEVEL=${PETALS_LOGLEVEL:-INFO}
while true; do
    python -m petals_b200.cli.run_server "$@"
    echo "stage exited with code $?; restarting in 5 s" >&2
    sleep 5
done

"""Command-line entry points (reference: src/petals/cli/)."""

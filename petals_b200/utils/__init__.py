"""Utilities (reference: src/petals/utils/)."""

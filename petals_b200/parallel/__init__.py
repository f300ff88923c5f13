"""Parallelism and intra-box communication: swarm registry, control transport, NVLink symmetric memory,
tensor-parallel plans, pipeline schedules."""

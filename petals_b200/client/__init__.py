"""Client side: model shells, RemoteSequential, inference sessions, routing (reference: src/petals/client/)."""

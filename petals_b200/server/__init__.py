"""Server side: one stage (span of blocks) per GPU worker (reference: src/petals/server/)."""

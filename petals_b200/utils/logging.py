"""Logging setup (reference: src/petals/utils/logging.py:1-18). Env: PETALS_LOGGING, PETALS_LOGLEVEL."""
import logging
import os

_FORMAT = "%(asctime)s.%(msecs)03d [%(levelname)s] [%(name)s:%(lineno)d] %(message)s"
_initialized = False


def initialize_logs() -> None:
    global _initialized
    if _initialized or os.getenv("PETALS_LOGGING", "True").lower() in ("false", "0"):
        _initialized = True
        return
    handler = logging.StreamHandler()
    handler.setFormatter(logging.Formatter(_FORMAT, datefmt="%b %d %H:%M:%S"))
    root = logging.getLogger("petals_b200")
    root.addHandler(handler)
    root.setLevel(os.getenv("PETALS_LOGLEVEL", "INFO").upper())
    root.propagate = False
    _initialized = True


def get_logger(name: str) -> logging.Logger:
    initialize_logs()
    if not name.startswith("petals_b200"):
        name = "petals_b200." + name
    return logging.getLogger(name)

"""Logging setup (reference: src/petals/utils/logging.py:1-18).

Env: ``PETALS_LOGGING`` (false = leave logging to the application), ``PETALS_LOGLEVEL``, ``PETALS_ASYNCIO_LOGLEVEL`` (level of
the ``asyncio`` logger, default WARNING like the reference), ``HIVEMIND_COLORS`` / ``PETALS_COLORS`` (force level colours on
or off; default: only when stderr is a terminal)."""
import logging
import os
import sys

_COLORS = {"DEBUG": "\033[36m", "INFO": "\033[32m", "WARNING": "\033[33m", "ERROR": "\033[31m", "CRITICAL": "\033[1;31m"}
_RESET = "\033[0m"


def _use_colors() -> bool:
    flag = os.getenv("PETALS_COLORS", os.getenv("HIVEMIND_COLORS"))
    if flag is not None:
        return flag.lower() in ("1", "true", "yes")
    return hasattr(sys.stderr, "isatty") and sys.stderr.isatty()


class _Formatter(logging.Formatter):
    def __init__(self, fmt: str, datefmt: str, colors: bool):
        super().__init__(fmt, datefmt=datefmt)
        self.colors = colors

    def format(self, record: logging.LogRecord) -> str:
        if not self.colors:
            return super().format(record)
        saved = record.levelname
        record.levelname = f"{_COLORS.get(saved, '')}{saved}{_RESET}"
        try:
            return super().format(record)
        finally:
            record.levelname = saved

_FORMAT = "%(asctime)s.%(msecs)03d [%(levelname)s] [%(name)s:%(lineno)d] %(message)s"
_initialized = False


def initialize_logs() -> None:
    global _initialized
    if _initialized or os.getenv("PETALS_LOGGING", "True").lower() in ("false", "0"):
        _initialized = True
        return
    handler = logging.StreamHandler()
    handler.setFormatter(_Formatter(_FORMAT, "%b %d %H:%M:%S", _use_colors()))
    root = logging.getLogger("petals_b200")
    root.addHandler(handler)
    root.setLevel(os.getenv("PETALS_LOGLEVEL", "INFO").upper())
    root.propagate = False
    logging.getLogger("asyncio").setLevel(os.getenv("PETALS_ASYNCIO_LOGLEVEL", "WARNING").upper())
    _initialized = True


def get_logger(name: str) -> logging.Logger:
    initialize_logs()
    if not name.startswith("petals_b200"):
        name = "petals_b200." + name
    return logging.getLogger(name)

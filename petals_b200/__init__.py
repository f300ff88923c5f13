"""petals_b200 — a Blackwell (sm_100a) native pipeline-parallel LLM serving and fine-tuning engine with the
capabilities and public API of bigscience-workshop/petals, rebuilt from scratch for one NVLink/NVSwitch box.

Public API mirrors the reference package root (src/petals/__init__.py:1-35).
"""
__version__ = "0.1.0"

from petals_b200.constants import DTYPE_MAP, PUBLIC_INITIAL_PEERS  # noqa: F401

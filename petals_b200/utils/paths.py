"""Locating checkpoints on an offline box."""
from __future__ import annotations

import glob
import os


def resolve_model_path(name_or_path: str) -> str:
    """Local directory for a checkpoint: a path, or an already-cached HF hub snapshot (offline box)."""
    if os.path.isdir(name_or_path):
        return name_or_path
    roots = [os.environ.get("PETALS_CACHE"), os.environ.get("HF_HOME") and os.path.join(os.environ["HF_HOME"], "hub"),
             os.path.expanduser("~/.cache/huggingface/hub"), os.path.expanduser("~/.cache/petals")]
    folder = "models--" + name_or_path.replace("/", "--")
    for root in filter(None, roots):
        snaps = sorted(glob.glob(os.path.join(root, folder, "snapshots", "*")))
        if snaps:
            return snaps[-1]
    raise FileNotFoundError(
        f"{name_or_path!r} is neither a local checkpoint directory nor a cached hub snapshot "
        "(this build runs offline; use petals_b200.utils.checkpoints.make_random_checkpoint to synthesise one)")



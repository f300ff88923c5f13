"""Disk cache housekeeping for checkpoints (reference: src/petals/utils/disk_cache.py:1-83): shared/exclusive
``flock`` around cache use and LRU eviction honouring ``max_disk_space``. Downloads do not exist on the offline box,
but local conversions / synthetic checkpoints live under the same cache directory and obey the same limits."""
from __future__ import annotations

import fcntl
import os
import shutil
from contextlib import contextmanager
from pathlib import Path
from typing import Optional

from petals_b200.utils.logging import get_logger

logger = get_logger(__name__)
DEFAULT_CACHE_DIR = os.getenv("PETALS_CACHE", str(Path(Path.home(), ".cache", "petals_b200")))
BLOCKS_LOCK_FILE = "blocks.lock"


@contextmanager
def _blocks_lock(cache_dir: Optional[str], mode: int):
    cache_dir = cache_dir or DEFAULT_CACHE_DIR
    os.makedirs(cache_dir, exist_ok=True)
    with open(Path(cache_dir, BLOCKS_LOCK_FILE), "wb+") as f:
        fcntl.flock(f.fileno(), mode)
        try:
            yield
        finally:
            fcntl.flock(f.fileno(), fcntl.LOCK_UN)


def allow_cache_reads(cache_dir: Optional[str]):
    """Many readers may load blocks concurrently."""
    return _blocks_lock(cache_dir, fcntl.LOCK_SH)


def allow_cache_writes(cache_dir: Optional[str]):
    """Writers (conversion, eviction) are exclusive."""
    return _blocks_lock(cache_dir, fcntl.LOCK_EX)


def _dir_size(path: Path) -> int:
    return sum(f.stat().st_size for f in path.rglob("*") if f.is_file())


def free_disk_space_for(size: int, *, cache_dir: Optional[str], max_disk_space: Optional[int], os_quota: int = 1024**3) -> None:
    """Evict least-recently-used model directories until ``size`` more bytes fit under the limits."""
    cache_dir = Path(cache_dir or DEFAULT_CACHE_DIR)
    os.makedirs(cache_dir, exist_ok=True)
    entries = [p for p in cache_dir.iterdir() if p.is_dir()]
    last_used = {p: p.stat().st_atime for p in entries}  # before walking the directories: reading them refreshes atime
    sizes = {p: _dir_size(p) for p in entries}
    occupied = sum(sizes.values())
    available = shutil.disk_usage(cache_dir).free - os_quota
    if max_disk_space is not None:
        available = min(available, max_disk_space - occupied)
    if size <= available:
        return
    needed = size - available
    for p in sorted(entries, key=lambda q: last_used[q]):  # LRU first
        logger.info(f"Evicting {p} ({sizes[p] / 2**30:.2f} GiB) from the checkpoint cache")
        shutil.rmtree(p, ignore_errors=True)
        needed -= sizes[p]
        if needed <= 0:
            return
    raise RuntimeError(f"Insufficient disk space for {size / 2**30:.1f} GiB even after evicting the cache")

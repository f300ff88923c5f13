"""In-tree build of the native library (``petals_b200/_native/libpetals_b200.so``).

Every ``.cu``/``.cpp`` under ``petals_b200/csrc`` is compiled for **sm_100a only**
(``-gencode arch=compute_100a,code=sm_100a -lineinfo``) and linked into one shared object that the
Python layer loads through ``ctypes`` (no torch C++ ABI dependency, seconds to build, and the ``.so``
travels with the source tree to the GPU box).  A content hash of the sources is stored next to the
library so a fresh tree only rebuilds when a source actually changed.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
OUT_DIR = ROOT / "_native"
LIB_PATH = OUT_DIR / "libpetals_b200.so"
RT_LIB_PATH = OUT_DIR / "libpetals_b200_rt.so"
STAMP = OUT_DIR / "build.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-pthread"]


def _nvcc() -> str | None:
    cand = os.environ.get("PB_NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return cand if cand and os.path.exists(cand) else None


def cuda_sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def host_sources() -> list[Path]:
    return sorted((CSRC / "runtime").glob("*.cpp"))


def source_hash() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                    + list((CSRC / "runtime").glob("*.cpp")) + list((CSRC / "runtime").glob("*.h"))):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS + CXX_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    return LIB_PATH.exists() and RT_LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == source_hash()


def _run(cmd: list[str]) -> None:
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"build command failed: {' '.join(cmd)}\n{proc.stdout}")


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every native source. Returns the path of the CUDA library."""
    if not force and is_current():
        return LIB_PATH
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build the sm_100a kernels (set PB_NVCC)")
    OUT_DIR.mkdir(exist_ok=True)
    obj_dir = OUT_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)

    # incremental: an object is rebuilt when its source, any header, or the flags changed
    hdr = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((CSRC / "runtime").glob("*.h"))):
        hdr.update(p.name.encode() + p.read_bytes())
    hdr.update(" ".join(NVCC_FLAGS + CXX_FLAGS).encode())

    def fresh(src: Path, obj: Path) -> bool:
        key = hashlib.sha256(hdr.digest() + src.read_bytes()).hexdigest()
        tag = obj.with_suffix(".hash")
        if not force and obj.exists() and tag.exists() and tag.read_text() == key:
            return True
        tag.unlink(missing_ok=True)
        return False

    def stamp(src: Path, obj: Path) -> None:
        obj.with_suffix(".hash").write_text(hashlib.sha256(hdr.digest() + src.read_bytes()).hexdigest())

    def compile_cu(src: Path) -> Path:
        obj = obj_dir / (src.stem + ".o")
        if not fresh(src, obj):
            _run([nvcc, *NVCC_FLAGS, "-I", str(CSRC), "-c", str(src), "-o", str(obj)])
            stamp(src, obj)
        return obj

    def compile_cpp(src: Path) -> Path:
        obj = obj_dir / ("rt_" + src.stem + ".o")
        if not fresh(src, obj):
            _run(["g++", *CXX_FLAGS, "-I", str(CSRC), "-c", str(src), "-o", str(obj)])
            stamp(src, obj)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        cu_objs = list(ex.map(compile_cu, cuda_sources()))
        cpp_objs = list(ex.map(compile_cpp, host_sources()))
    _run([nvcc, "-shared", "-o", str(LIB_PATH), *map(str, cu_objs), "-cudart", "static"])
    # Host-only runtime (scheduler, KV page allocator, safetensors reader): no CUDA dependency, so it
    # also loads — and is unit-tested — on the CPU-only build box.
    _run(["g++", "-shared", "-o", str(RT_LIB_PATH), *map(str, cpp_objs), "-pthread"])
    import ctypes

    ctypes.CDLL(str(LIB_PATH))  # unresolved symbols must fail the build here, not on the GPU box
    ctypes.CDLL(str(RT_LIB_PATH))
    STAMP.write_text(source_hash())
    if verbose:
        print(f"built {LIB_PATH} and {RT_LIB_PATH}", file=sys.stderr)
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)

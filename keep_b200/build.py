"""In-tree build of libkeep_b200.so (nvcc, sm_100a only).

    python -m keep_b200.build [--force] [--verbose]

Each csrc/*.cu is compiled to keep_b200/_build/*.o (in parallel, skipped when up to date) and linked into
keep_b200/libkeep_b200.so next to this file, so the library travels with the source tree.  nvcc
cross-compiles for sm_100a without a GPU present.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "_build"
LIB_PATH = PKG_DIR / "libkeep_b200.so"
INCLUDE_DIR = PKG_DIR.parent / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


class ToolchainMissing(RuntimeError):
    """nvcc is not on this box (a prebuilt in-tree library may still be usable)."""


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise ToolchainMissing("nvcc not found: keep_b200 needs the CUDA 12.9 toolkit to build its sm_100a kernels")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list(INCLUDE_DIR.glob("*.h"))
    return max(p.stat().st_mtime for p in hdrs)


def is_stale() -> bool:
    if not LIB_PATH.exists():
        return True
    lib_m = LIB_PATH.stat().st_mtime
    return any(p.stat().st_mtime > lib_m for p in _sources()) or _deps_mtime() > lib_m


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile (if needed) and return the path of libkeep_b200.so.

    Safe to call from several processes at once (every rank of a torchrun job does): the stale check, the compile and
    the link run under an exclusive file lock, objects and the library are written to temporary names and moved into
    place with os.replace, so no process ever maps a half-written file."""
    if not force and not is_stale():
        return LIB_PATH
    import fcntl

    BUILD_DIR.mkdir(exist_ok=True)
    with open(BUILD_DIR / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():  # another process built it while this one waited
                return LIB_PATH
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> Path:
    nvcc = _nvcc()
    hdr_m = _deps_mtime()

    def compile_one(src: Path) -> Path:
        obj = BUILD_DIR / (src.stem + ".o")
        if not force and obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, hdr_m):
            return obj
        tmp = obj.with_suffix(f".o.tmp{os.getpid()}")
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(INCLUDE_DIR), "-c", str(src), "-o", str(tmp)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            tmp.unlink(missing_ok=True)
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, flush=True)
        os.replace(tmp, obj)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    # the CUDA runtime is linked statically (nvcc default): no libcudart/libcuda lookup at load time
    tmp_lib = LIB_PATH.with_suffix(f".so.tmp{os.getpid()}")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp_lib), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        tmp_lib.unlink(missing_ok=True)
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp_lib, LIB_PATH)
    return LIB_PATH


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    path = build(force=a.force, verbose=a.verbose)
    print(path)


if __name__ == "__main__":
    sys.exit(main())

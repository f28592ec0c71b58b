"""Build libp24_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

``python -m p24.build`` or ``p24.build.build()``.  nvcc cross-compiles without a GPU; the built
library lives in ``p24/_lib/`` (git-ignored, travels to the GPU box with the snapshot).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.abspath(os.path.join(HERE, "..", "csrc"))
INCLUDE = os.path.abspath(os.path.join(HERE, "..", "..", "include"))
LIBDIR = os.path.join(HERE, "_lib")
LIBNAME = "libp24_b200_timing.so" if os.environ.get("P24_TIMING") else "libp24_b200.so"

# -fmad=false: the SimOTA decisions are fp32 threshold tests evaluated in the reference's operation
# order (one rounding per op, like eager PyTorch); bounds / backward code uses explicit fmaf.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false",
              "-std=c++17", "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"] + \
             (["-DP24_TIMING"] if os.environ.get("P24_TIMING") else [])  # debug build: phase timers (tests/tools/timeline.py)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the p24 CUDA library cannot be built")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + \
            [os.path.join(INCLUDE, f) for f in sorted(os.listdir(INCLUDE))]
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode())  # names, not absolute paths: the tree moves (GPU box snapshot)
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def lib_path() -> str:
    return os.path.join(LIBDIR, LIBNAME)


def is_fresh() -> bool:
    stamp = os.path.join(LIBDIR, LIBNAME + ".stamp")
    if not (os.path.exists(lib_path()) and os.path.exists(stamp)):
        return False
    with open(stamp) as fh:
        return fh.read().strip() == _fingerprint()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile when the sources changed.  Safe when several processes (one per GPU) call it at once: an exclusive
    file lock serialises them, the library is written under a temporary name and renamed into place."""
    import fcntl
    os.makedirs(LIBDIR, exist_ok=True)
    if not force and is_fresh():
        return lib_path()
    with open(os.path.join(LIBDIR, "build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_fresh():  # another process built it while this one waited
                return lib_path()
            tmp = lib_path() + f".tmp{os.getpid()}"
            cmd = [_nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC] + (["-Xptxas", "-v"] if verbose else []) + \
                  sources() + ["-o", tmp]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                raise RuntimeError("nvcc failed building libp24_b200.so")
            if verbose:
                print(res.stdout + res.stderr)
            os.replace(tmp, lib_path())
            with open(os.path.join(LIBDIR, LIBNAME + ".stamp"), "w") as fh:
                fh.write(_fingerprint())
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return lib_path()


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

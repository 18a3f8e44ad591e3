"""In-tree build of the CUDA extension (nvcc, sm_100a only).  `python -m gym_narde_b200.build`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libnarde_b200.so")
LIB_DEBUG = os.path.join(_HERE, "libnarde_b200_debug.so")   # tools only: -DNARDE_DEBUG_HOOKS (A/B switches, phase clocks)
SOURCES = ["narde_kernels.cu", "narde_mlp.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [
    os.path.join("..", "..", "include", "narde_b200.h")]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the sm_100a CUDA extension cannot be built")
    return p


def is_stale(lib: str = LIB) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, debug_hooks: bool = False) -> str:
    lib = LIB_DEBUG if debug_hooks else LIB
    if not force and not is_stale(lib):
        return lib
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "--shared", "-Xcompiler", "-fPIC", "-o", lib] + [os.path.join(CSRC, s) for s in SOURCES]
    if debug_hooks:
        cmd.insert(1, "-DNARDE_DEBUG_HOOKS=1")
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd, cwd=CSRC)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug_hooks="--debug-hooks" in sys.argv))

"""In-tree build of libtarok_b200.so (nvcc, sm_100a only -- no other architecture is targeted).

    python -m tarok_b200.build          # build if stale
    python -m tarok_b200.build --force

The shared library lands next to this file (git-ignored, travels to the GPU box with gpurun).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libtarok_b200.so")
SOURCES = [os.path.join(CSRC, "tarok_abi.cu")]
HOST_SOURCES = [os.path.join(CSRC, "tarok_host.cpp")]          # plain C++ (g++): the host-side record serialiser
HOST_OBJ = os.path.join(CSRC, "tarok_host.o")
HEADERS = [os.path.join(CSRC, f) for f in ("tarok_kernels.cuh", "tarok_obs.cuh", "tarok_rules.cuh", "philox.cuh",
                                           "tarok_host.h")] + [os.path.join(ROOT, "include", "tarok_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; tarok_b200 cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SOURCES + HOST_SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False, variant: str = "", defines=()) -> str:
    """``variant`` / ``defines``: build libtarok_b200_<variant>.so with extra -D macros (occupancy A/B experiments; a
    variant is loaded with TAROK_B200_LIB=<path>).  The default library takes no macros."""
    if variant:
        out = os.path.join(PKG, "libtarok_b200_%s.so" % variant)
        cmd = [nvcc_path()] + NVCC_FLAGS + ["-D%s" % d for d in defines] + ["-o", out] + SOURCES + [HOST_OBJ, "-ldl", "-lpthread"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        with open(os.path.join(PKG, "build_%s.log" % variant), "w") as f:
            f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        return out
    if not force and not is_stale():
        return LIB
    gxx = shutil.which("g++") or "/usr/bin/g++"
    host = [gxx, "-O3", "-std=c++17", "-fPIC", "-pthread", "-Wall", "-c", "-o", HOST_OBJ] + HOST_SOURCES
    res = subprocess.run(host, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", LIB] + SOURCES + [HOST_OBJ, "-ldl", "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = " ".join(host) + "\n" + res.stdout + res.stderr
    with open(os.path.join(PKG, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log)
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    if "--variant" in sys.argv:                       # python -m tarok_b200.build --variant name -DMACRO=1 ...
        name = sys.argv[sys.argv.index("--variant") + 1]
        build_library(force=False)
        print(build_library(variant=name, defines=[a[2:] for a in sys.argv if a.startswith("-D")]))
    else:
        build_library(force="--force" in sys.argv, verbose=True)
        print(LIB)

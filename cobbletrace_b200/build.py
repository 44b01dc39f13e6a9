"""In-tree builds of the native libraries (no JIT cache: the .so files travel with the repo snapshot).

    libct_gpu.so   CUDA kernels + the C ABI of include/ct_gpu.h      (nvcc, sm_100a only)
    libct_host.so  host side: scene ingest, BVH build, boss loop     (g++)

``python -m cobbletrace_b200.build`` builds both.
"""
from __future__ import annotations

import contextlib
import fcntl
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
GPU_LIB = os.path.join(PKG, "libct_gpu.so")
HOST_LIB = os.path.join(PKG, "libct_host.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    # The reference's rounding points are reproduced with explicit *_rn intrinsics; -fmad=false makes
    # sure nothing else in the TU is contracted behind our back either.  Never --use_fast_math.
    "-fmad=false", "-Xcompiler", "-fPIC", "-shared",
]
GXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Wextra", "-pthread"]


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _glob(d, exts):
    out = []
    for base, _, files in os.walk(d):
        out += [os.path.join(base, f) for f in files if f.endswith(exts)]
    return sorted(out)


@contextlib.contextmanager
def _build_lock():
    """One builder at a time per checkout: the ranks of a torchrun job all call build_all() on start-up."""
    with open(os.path.join(PKG, ".build.lock"), "a+") as f:
        fcntl.flock(f, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(f, fcntl.LOCK_UN)


def _compile(cmd, target: str):
    """Compile to a temporary name and rename: a library another process has loaded (or is loading) is never rewritten."""
    tmp = f"{target}.tmp{os.getpid()}"
    try:
        subprocess.check_call([tmp if c == target else c for c in cmd], cwd=ROOT)
        os.replace(tmp, target)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)


def find_nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: cobbletrace_b200 has no CPU fallback and cannot be built without CUDA")


def build_gpu(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, "ct_gpu.cu")]
    deps = srcs + _glob(CSRC, (".cuh",)) + [os.path.join(ROOT, "include", "ct_gpu.h")]
    if not force and _newer(GPU_LIB, deps):
        return GPU_LIB
    with _build_lock():
        if not force and _newer(GPU_LIB, deps):                 # another rank built it while we waited
            return GPU_LIB
        extra = os.environ.get("CT_NVCC_EXTRA", "").split()     # experiments only, e.g. -DCT_MIN_BLOCKS=5
        cmd = [find_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", GPU_LIB] + srcs
        _compile(cmd, GPU_LIB)
    return GPU_LIB


def build_host(force: bool = False) -> str:
    hdir = os.path.join(CSRC, "host")
    srcs = _glob(hdir, (".cpp",))
    deps = srcs + _glob(hdir, (".h", ".hpp")) + [os.path.join(ROOT, "include", "ct_gpu.h"), os.path.join(ROOT, "include", "ct_host.h")]
    if not force and _newer(HOST_LIB, deps):
        return HOST_LIB
    # Deliberately the PATH g++ and not $CXX: this image exports CXX=/opt/gcc/bin/g++, whose libstdc++.so link
    # dangles, so it silently links libstdc++.a into the .so -- a second C++ runtime next to the one torch
    # loads, and exceptions thrown inside the library then crash the process.
    with _build_lock():
        if not force and _newer(HOST_LIB, deps):
            return HOST_LIB
        cmd = [shutil.which("g++") or "g++"] + GXX_FLAGS + ["-I", os.path.join(ROOT, "include"), "-o", HOST_LIB] + srcs + ["-ldl"]
        _compile(cmd, HOST_LIB)
    return HOST_LIB


def build_all(force: bool = False, verbose: bool = False):
    return build_gpu(force, verbose), build_host(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))

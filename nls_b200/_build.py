"""In-tree build of libnls_b200.so with nvcc for sm_100a (no JIT cache: the .so travels with the tree)."""

from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnls_b200.so")

SOURCES = ["api.cu", "kernels_1d.cu", "kernels_2d.cu", "fused_2d.cu", "stream_2d.cu", "resident_2d.cu", "reduce.cu", "diagnostics.cu", "pumping_gen.cu", "operators.cpp"]
HEADERS = ["internal.h", "kernels.h", "device_math.cuh", "stream_2d_core.cuh", "resident_2d_core.cuh", os.path.join("..", "..", "include", "nls_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-Xptxas", "-v",
    "-cudart", "static",
    "-shared",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libnls_b200.so cannot be built")


def is_stale():
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > built for d in deps)


def build_library(force=False, verbose=False):
    """Compile every CUDA source of the engine into ``nls_b200/libnls_b200.so``."""
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + proc.stdout)
    if verbose or proc.returncode:
        print(proc.stdout)
    if proc.returncode:
        raise RuntimeError("nvcc failed (see %s)" % log)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))

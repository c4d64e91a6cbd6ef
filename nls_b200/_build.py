"""In-tree build of libnls_b200.so with nvcc for sm_100a (no JIT cache: the .so travels with the tree)."""

from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnls_b200.so")

SOURCES = ["api.cu", "kernels_1d.cu", "kernels_2d.cu", "fused_2d.cu", "stream_2d.cu", "resident_2d.cu", "reduce.cu", "diagnostics.cu", "pumping_gen.cu", "peer.cu", "operators.cpp"]
HEADERS = ["internal.h", "kernels.h", "device_math.cuh", "diag_acc.cuh", "peer_flags.cuh", "stream_2d_core.cuh", "resident_2d_core.cuh", os.path.join("..", "..", "include", "nls_b200.h")]

COMPILE_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-Xptxas", "-v",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-shared"]


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


def _compile_one(nvcc, src, obj):
    cmd = [nvcc] + COMPILE_FLAGS + ["-c", "-o", obj, src]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return " ".join(cmd) + "\n" + proc.stdout, proc.returncode


def build_library(force=False, verbose=False):
    """Compile every CUDA source of the engine into ``nls_b200/libnls_b200.so`` (one nvcc per source, in
    parallel; objects under ``nls_b200/build/``; a source is recompiled when it or any header is newer)."""
    if not force and not is_stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = nvcc_path()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    newest_header = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    jobs, objs = [], []
    for f in SOURCES:
        src, obj = os.path.join(CSRC, f), os.path.join(objdir, f + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header):
            jobs.append((src, obj))
    logs, failed = [], False
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        for text, rc in pool.map(lambda j: _compile_one(nvcc, *j), jobs):
            logs.append(text)
            failed = failed or rc != 0
    if not failed:
        cmd = [nvcc] + LINK_FLAGS + ["-o", LIB] + objs
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        logs.append(" ".join(cmd) + "\n" + proc.stdout)
        failed = proc.returncode != 0
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as fh:
        fh.write("\n".join(logs))
    if verbose or failed:
        print("\n".join(logs))
    if failed:
        raise RuntimeError("nvcc failed (see %s)" % log)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))

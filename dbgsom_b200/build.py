"""Build libdbgsom_b200.so in-tree with nvcc for sm_100a (`python -m dbgsom_b200.build`)."""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libdbgsom_b200.so")
SOURCES = ["capi.cu", "prep.cu", "bmu_simt.cu", "bmu_tc.cu", "bmu_resolve.cu", "accumulate.cu", "smooth.cu", "post.cu", "lars.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "lars_core.cuh"), os.path.join(HERE, "..", "include", "dbgsom_b200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libdbgsom_b200.so")


OBJ_DIR = os.path.join(HERE, "build")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    return _stale(LIB_PATH, [os.path.join(CSRC, s) for s in SOURCES] + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a (one object per source, in parallel, re-using objects that are newer
    than their source and the headers) and link them into one shared library; returns its path."""
    if not force and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(OBJ_DIR, exist_ok=True)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
    if verbose:
        flags.insert(0, "-Xptxas=-v")

    def compile_one(src: str):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if not force and not _stale(obj, [os.path.join(CSRC, src)] + HEADERS):
            return obj, None
        cmd = [_nvcc(), *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            return obj, res.stdout + res.stderr
        if verbose:
            print(res.stdout + res.stderr)
        return obj, None

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    errors = [err for _, err in results if err]
    if errors:
        sys.stderr.write("\n".join(errors))
        raise RuntimeError("nvcc failed building libdbgsom_b200.so")
    link = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + [o for o, _ in results]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libdbgsom_b200.so")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

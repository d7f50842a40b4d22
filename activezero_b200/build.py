"""Ahead-of-time build of ``libaz_stereo.so`` (sm_100a only, in-tree).

    python -m activezero_b200.build [--force]

nvcc cross-compiles without a GPU; the built library is git-ignored but travels
with the ``gpurun`` snapshot.  There is no JIT at run time (the reference
JIT-compiles its one kernel with NVRTC on first call, ``utils/warp_ops.py:48-52``).
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libaz_stereo.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(
        os.path.join(os.path.dirname(PKG_DIR), "include", "*.h"))
    return any(os.path.getmtime(f) > t for f in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu to an object (in parallel, only the stale ones unless `force`) and link."""
    if not force and not _stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    nvcc = os.environ.get("NVCC", "nvcc")
    obj_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(obj_dir, exist_ok=True)
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(os.path.dirname(PKG_DIR), "include", "*.h"))
    hdr_t = max(os.path.getmtime(h) for h in hdrs)
    flags = [f for f in NVCC_FLAGS if f not in ("-shared", "-cudart", "static")]
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append([nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(subprocess.check_call, jobs))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static"]
                          + objs + ["-o", LIB_PATH])
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Build ``oracle/_ref/libwarp_ref.so``: the reference's OWN scatter-warp
kernel source, compiled for the CPU (TEST INFRASTRUCTURE ONLY).

The reference's only native code is a CUDA-C string that it JIT-compiles with
NVRTC (``/root/reference/utils/warp_ops.py:20-47``).  The two kernels are plain
scalar C apart from ``__global__`` and the ``blockIdx/blockDim/threadIdx``
builtins, so this recipe
  1. reads the string from where it lies under ``/root/reference`` (nothing is
     copied into the repository -- the extracted text goes to a temp dir that
     is deleted, only the ``.so`` lands in the git-ignored ``oracle/_ref/``),
  2. wraps it in a shim (``oracle/c/ref_kernel_shim.h``) that defines those
     builtins as thread-local variables, and
  3. compiles it with ``g++`` together with a driver that replays the
     reference launch geometry (``grid = total_l // 512 + 1``, ``block = 512``,
     ``warp_ops.py:86-93``) thread by thread.
The result executes the unmodified kernel bodies; it is what pins
``stereo_oracle.scatter_warp`` and the golden fixtures for row a10.
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_FILE = os.path.join(os.environ.get("AZ_REFERENCE_ROOT", "/root/reference"), "utils", "warp_ops.py")
OUT_DIR = os.path.join(HERE, "_ref")
OUT_SO = os.path.join(OUT_DIR, "libwarp_ref.so")

DRIVER = r"""
#include "ref_kernel_shim.h"
#include "ref_kernels.inc"

extern "C" void ref_apply_disparity(int positive, float* dst, const float* src, const int* disp,
                                    int h, int w, int c, int total_l) {
    const int block = 512;                 /* warp_ops.py:88 */
    const int grid = total_l / 512 + 1;    /* warp_ops.py:86 */
    blockDim.x = block;
    for (int b = 0; b < grid; ++b) {
        blockIdx.x = b;
        for (int t = 0; t < block; ++t) {
            threadIdx.x = t;
            if (positive) apply_disparity_pos(dst, src, disp, h, w, c, total_l);
            else          apply_disparity_neg(dst, src, disp, h, w, c, total_l);
        }
    }
}
"""


def build(force: bool = False) -> str | None:
    """Returns the path of the built library, or None if the reference tree is
    absent (GPU box) and no prebuilt library travelled with the snapshot."""
    if os.path.exists(OUT_SO) and not force:
        return OUT_SO
    if not os.path.exists(REF_FILE):
        return None
    text = open(REF_FILE).read()
    m = re.search(r'_apply_disparity_pos_kernel\s*=\s*"""(.*?)"""', text, re.S)
    if m is None:
        raise RuntimeError("kernel string not found in reference warp_ops.py")
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="az_ref_")
    try:
        with open(os.path.join(tmp, "ref_kernels.inc"), "w") as f:
            f.write(m.group(1))
        with open(os.path.join(tmp, "driver.cpp"), "w") as f:
            f.write(DRIVER)
        subprocess.check_call(
            ["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(HERE, "c"), "-I", tmp,
             os.path.join(tmp, "driver.cpp"), "-o", OUT_SO]
        )
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT_SO


def load():
    import ctypes

    path = build()
    if path is None:
        return None
    lib = ctypes.CDLL(path)
    lib.ref_apply_disparity.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.ref_apply_disparity.restype = None
    return lib


def ref_scatter_warp(img, disp):
    """Run the reference kernels on CPU tensors, with the host logic of
    ``warp_ops.py:69-95`` (asserts, sign dispatch, zero-filled output)."""
    import torch

    lib = load()
    if lib is None:
        raise RuntimeError("oracle/_ref/libwarp_ref.so unavailable")
    assert img.is_contiguous() and disp.is_contiguous() and disp.dtype == torch.int32
    assert img.dtype == torch.float32
    if bool(torch.all(disp >= 0)):
        positive = 1
    else:
        assert bool(torch.all(disp <= 0))
        positive = 0
    out = torch.zeros_like(img)
    b, c, h, w = img.shape
    lib.ref_apply_disparity(positive, out.data_ptr(), img.data_ptr(), disp.data_ptr(), h, w, c, b * c * h)
    return out


if __name__ == "__main__":
    print(build(force=True))

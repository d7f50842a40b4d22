"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly the
symbols include/az_stereo.h declares (no compute calls -- no GPU here)."""
import ctypes
import os
import shutil
import subprocess

import pytest

import activezero_b200
from activezero_b200 import _lib, build as az_build


@pytest.fixture(scope="module")
def lib_path():
    if shutil.which("nvcc") is None and not os.path.exists(_lib.LIB_PATH):
        pytest.skip("no nvcc and no prebuilt library")
    if shutil.which("nvcc") is not None:
        az_build.build()
    return _lib.LIB_PATH


def test_header_and_binding_agree():
    declared = _lib.header_symbols()
    assert declared, "no declarations parsed from include/az_stereo.h"
    assert sorted(_lib.SIGNATURES) == declared


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in _lib.header_symbols():
        assert hasattr(lib, name), f"{name} declared in include/az_stereo.h but not exported"


def test_library_loads_through_binding(lib_path):
    lib = _lib.load()
    assert lib.az_version().decode().endswith("sm_100a")
    assert b"bad argument" in lib.az_error_string(-1)
    assert _lib.query("az_reproj_workspace_bytes", 2, 10) == 2 * 10 * 16
    assert _lib.query("az_temporal_ir_workspace_bytes", 2, 4, 5) == (2 * 4 * 5 + 4) * 4  # int32 since round 2 (integer form)


def test_library_is_sm100a_only(lib_path):
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    archs = {tok for line in out.splitlines() for tok in line.replace(".", " ").split() if tok.startswith("sm_")}
    assert archs == {"sm_100a"}, archs


def test_bad_arguments_are_rejected_without_a_gpu(lib_path):
    lib = _lib.load()
    null = ctypes.c_void_p(0)
    assert lib.az_concat_volume_fwd(null, null, null, 1, 1, 1, 1, 1, null) == -1
    assert lib.az_soft_argmin_fwd(null, null, null, 1, 1, 1, 1, null) == -1
    assert lib.az_scatter_warp(null, null, null, null, 1, 1, 1, 1, null) == -1


def test_ops_refuse_cpu_tensors():
    """No CPU path and no silent fallback: the operators raise on CPU tensors."""
    import torch

    from activezero_b200 import ops

    x = torch.zeros(1, 2, 4, 8)
    with pytest.raises(ValueError):
        ops.build_concat_volume(x, x, 2)
    with pytest.raises(ValueError):
        ops.soft_argmin(torch.zeros(1, 4, 4, 8))
    with pytest.raises(ValueError):
        ops.warp(x, torch.zeros(1, 1, 4, 8))
    with pytest.raises(AssertionError):
        ops.scatter_warp(x, torch.zeros(1, 1, 4, 8, dtype=torch.int32))


def test_package_does_not_import_oracle():
    import sys

    assert activezero_b200.__version__
    pkg_dir = os.path.dirname(activezero_b200.__file__)
    for root, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
    del sys

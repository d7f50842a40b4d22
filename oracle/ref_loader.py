"""Import the real reference modules from ``/root/reference`` (authoring
container only -- that path does not exist on the GPU box, so nothing in the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this).

``utils/reprojection.py:10`` imports ``utils/warp_ops.py`` which imports
``cupy`` / ``pynvrtc`` at module scope (``warp_ops.py:9-10``); neither is
installed, so both are stubbed in ``sys.modules``.  The stubs are never
executed: the NVRTC kernel path needs a GPU and is restated in
``stereo_oracle.scatter_warp`` instead.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("AZ_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "utils"))


def _stub(name: str, **attrs):
    if name not in sys.modules:
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod
    return sys.modules[name]


def load():
    """Returns (reprojection_module, psmnet_submodule_module)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    fn = _stub("cupy.cuda.function", Module=object)
    cuda = _stub("cupy.cuda", function=fn)
    _stub("cupy", cuda=cuda)
    comp = _stub("pynvrtc.compiler", Program=object)
    _stub("pynvrtc", compiler=comp)
    # The reference's top-level package names (`utils`, `nets`) are generic;
    # import them under a private root so they cannot shadow anything else.
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    try:
        reproj = importlib.import_module("utils.reprojection")
        sub = importlib.import_module("nets.psmnet.psmnet_submodule")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    return reproj, sub

"""activezero_b200 -- B200 (sm_100a) implementation of ActiveZero's stereo
hot path: cost-volume construction, soft-argmin disparity regression and the
disparity-driven warps / reprojection losses, behind the reference's own Python
call sites:

    activezero_b200.nets.psmnet.psmnet.PSMNet          <- nets/psmnet/psmnet.py
    activezero_b200.nets.psmnet.psmnet_3.PSMNet        <- nets/psmnet/psmnet_3.py
    activezero_b200.utils.reprojection.*               <- utils/reprojection.py
    activezero_b200.utils.warp_ops.apply_disparity_cu  <- utils/warp_ops.py
    activezero_b200.tools.temporal_ir.*                <- tools/temporal_ir.py

All arithmetic runs in ``libaz_stereo.so`` (hand-written CUDA, C ABI in
``include/az_stereo.h``); importing this package never touches ``oracle/``.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"

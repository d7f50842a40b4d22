"""Drop-in for ``/root/reference/utils/warp_ops.py``: the integer scatter warp.

Same name, argument meaning, asserts and return value as the reference's
``apply_disparity_cu`` (``warp_ops.py:55-95``); the NVRTC-JIT one-thread-per-row
kernels (``:20-47``) are replaced by the ahead-of-time compiled
``az_scatter_warp`` (one CTA per disparity row, shared-memory integer atomicMax,
coalesced stores)."""
import torch

from .. import ops


def apply_disparity_cu(img: torch.Tensor, disp: torch.Tensor):
    """
    :param img: tensor needed warping. (N, C, H, W)
    :param disp: (N, H, W) or (N, 1, H, W), int32, all >= 0 or all <= 0
    :return: warped tensor like ``img``; un-hit pixels are 0
    """
    return ops.scatter_warp(img, disp, check_sign=True)

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


def disp_gt_from_right_view(img_disp_r_2x: torch.Tensor, max_disp: float):
    """Opt-in replacement of the twelve lines around ``apply_disparity_cu`` in the trainer
    (``/root/reference/train.py:255-272``; same chain in ``test.py:109-110``): the double-resolution right-view
    disparity -> nearest x0.5 -> ``.type(torch.int)`` -> scatter warp of the disparity by itself -> training mask.

    :param img_disp_r_2x: (N, 1, 2H, 2W) float32, what ``sample["img_disp_R"]`` holds
    :return: ``(disp_gt_l (N,1,H,W) float32, mask (N,1,H,W) bool)`` -- identical to the reference chain
    """
    return ops.scatter_warp_gt(img_disp_r_2x, max_disp, check_sign=True)

"""GPU counterparts of the IR-pattern helpers of ``/root/reference/datasets/dataset_utils.py``
(SURVEY.md §8f rank 3).  In the reference they run per sample on the CPU inside the DataLoader
workers (``datasets/messytable.py:221-232, 408-428``); here a batch of images already on the device is
processed by ``az_sim_ir_pattern``.  Same names and argument meaning; inputs are CUDA tensors --
uint8 grey levels as loaded from the PNGs, or float64 images already divided by 255 -- of shape
``[H,W]`` or ``[B,H,W]``; the {0,1} pattern comes back as float32 (the dataset converts it to float32
right away, ``messytable.py:227-232``)."""
from .. import ops


def get_ir_pattern(img_ir, img, threshold=0.005):
    """dataset_utils.py:12-17."""
    return ops.sim_ir_pattern(img_ir, img, ks=0, threshold=threshold)


def get_smoothed_ir_pattern2(img_ir, img, ks=11, threshold=0.005):
    """dataset_utils.py:33-46."""
    return ops.sim_ir_pattern(img_ir, img, ks=ks, threshold=threshold)

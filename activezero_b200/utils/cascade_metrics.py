"""Drop-in for ``compute_err_metric`` of ``/root/reference/utils/cascade_metrics.py:16-57``
(SURVEY.md §8f rank 4): same arguments, same returned dict of Python floats, but the seven
boolean-mask gathers + ``.item()`` syncs per training step (``train.py:348-351``) become one
fused masked reduction and a single 64-byte device->host read."""
import torch

from .. import ops


@torch.no_grad()
def compute_err_metric(disp_gt, depth_gt, disp_pred, focal_length, baseline, mask, depth_pred=None):
    """
    :param disp_gt, depth_gt, disp_pred: [bs, 1, H, W]
    :param focal_length, baseline: one value per sample (any shape with bs elements, e.g. [bs,1,1,1])
    :param mask: selected pixels, bool [bs, 1, H, W]
    :return: {"epe","bad1","bad2","depth_abs_err","depth_err2","depth_err4","depth_err8"}
    """
    s = ops.error_metric_sums(disp_gt, depth_gt, disp_pred, mask, depth_pred=depth_pred,
                              focal_length=focal_length, baseline=baseline).tolist()
    n = s[0]
    div = (lambda v: v / n) if n > 0 else (lambda v: float("nan"))  # empty mask: the reference divides by zero
    return {"epe": div(s[1]), "bad1": div(s[2]), "bad2": div(s[3]), "depth_abs_err": div(s[4]),
            "depth_err2": div(s[5]), "depth_err4": div(s[6]), "depth_err8": div(s[7])}

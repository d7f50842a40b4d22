"""Drop-in for ``/root/reference/nets/psmnet/psmnet_submodule.py`` (6-channel
feature extractor: image + adapter-transformed image)."""
import numpy as np
import torch
import torch.nn as nn

from ._backbone import BasicBlock, FeatureExtractionBase, conv, convbn, convbn_3d  # noqa: F401


class DisparityRegression(nn.Module):
    """psmnet_submodule.py:80-89, kept for API compatibility: ``sum_d prob[:,d]*d``
    of a PROBABILITY volume.  The PSMNet modules of this package do not use it --
    they call the fused ``ops.soft_argmin`` on the logits instead, which never
    materialises the probabilities.  Device-agnostic (the reference hard-codes
    ``.cuda()`` at :83-85)."""

    def __init__(self, maxdisp):
        super().__init__()
        self.register_buffer("disp", torch.tensor(np.arange(maxdisp), dtype=torch.float32).view(1, maxdisp, 1, 1),
                             persistent=False)

    def forward(self, x):
        return torch.sum(x * self.disp.to(x.device), 1, keepdim=True)


class FeatureExtraction(FeatureExtractionBase):
    def __init__(self):
        super().__init__(in_channels=6)

    def forward(self, x, x_transformed):
        """[bs,3,H,W] x2 -> [bs,32,H/4,W/4]"""
        return self._features(torch.cat((x, x_transformed), 1))

"""Drop-in for ``/root/reference/nets/psmnet/psmnet_submodule_3.py`` (3-channel
feature extractor, no adapter input)."""
from ._backbone import BasicBlock, FeatureExtractionBase, conv, convbn, convbn_3d  # noqa: F401
from .psmnet_submodule import DisparityRegression  # noqa: F401


class FeatureExtraction(FeatureExtractionBase):
    def __init__(self):
        super().__init__(in_channels=3)

    def forward(self, x):
        """[bs,3,H,W] -> [bs,32,H/4,W/4]"""
        return self._features(x)

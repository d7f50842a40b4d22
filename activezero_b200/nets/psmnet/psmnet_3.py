"""Drop-in for ``/root/reference/nets/psmnet/psmnet_3.py``: PSMNet without the
adapter input (3-channel images, two-argument ``forward``, :144-146)."""
from ._backbone import PSMNetBase, hourglass  # noqa: F401
from .psmnet_submodule_3 import *  # noqa: F401,F403
from .psmnet_submodule_3 import FeatureExtraction


class PSMNet(PSMNetBase):
    def __init__(self, maxdisp=192):
        super().__init__(FeatureExtraction(), maxdisp)

    def forward(self, img_L, img_R):
        ref = self.feature_extraction(img_L)
        tgt = self.feature_extraction(img_R)
        return self._forward_features(ref, tgt)

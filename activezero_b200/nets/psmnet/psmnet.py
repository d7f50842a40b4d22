"""Drop-in for ``/root/reference/nets/psmnet/psmnet.py``: the adapter variant of
PSMNet.  ``forward`` is the reference's (:144-222) with the inline cost-volume
loop (:151-165) replaced by ``ops.build_concat_volume`` and every
``F.softmax`` + ``DisparityRegression`` pair (:200-201, 204-205, 212-217) by
``ops.soft_argmin``; module names and ``state_dict`` keys are unchanged."""
from ._backbone import PSMNetBase, hourglass  # noqa: F401
from .psmnet_submodule import *  # noqa: F401,F403  (the reference module star-imports its submodule)
from .psmnet_submodule import FeatureExtraction


class PSMNet(PSMNetBase):
    def __init__(self, maxdisp=192):
        super().__init__(FeatureExtraction(), maxdisp)

    def forward(self, img_L, img_R, img_L_transformed, img_R_transformed):
        ref = self.feature_extraction(img_L, img_L_transformed)  # [bs,32,H/4,W/4]
        tgt = self.feature_extraction(img_R, img_R_transformed)
        return self._forward_features(ref, tgt)  # (pred3,pred2,pred1) when training, else pred3

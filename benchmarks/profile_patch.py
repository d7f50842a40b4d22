#!/usr/bin/env python
"""Launch the patch loss + Fold kernel a few times at BASELINE config-2 size on the bench's disparity field
(for `ncu --set full -k regex:patch_loss_fold`, see profiles/README.md).  AZ_PATCH_IMPL selects the kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from activezero_b200 import ops  # noqa: E402

B, H, W, D, PS = 8, 544, 960, 192, 11
gen = torch.Generator().manual_seed(5)
pL = (torch.rand(B, 1, H, W, generator=gen) > 0.5).float().cuda()
pR = (torch.rand(B, 1, H, W, generator=gen) > 0.5).float().cuda()
mask = (torch.rand(B, 1, H, W, generator=gen) > 0.2).cuda()
low = (torch.randn(B, 1, D // 4, H // 4, W // 4, generator=gen) * 4).cuda()
d = ops.upsample_soft_argmin(low, (D, H, W))
for _ in range(3):
    loss, vis = ops.reproj_loss(pL, pR, d, mask, ps=PS, want_warped=True)
torch.cuda.synchronize()
print("loss", float(loss))

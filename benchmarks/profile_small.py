#!/usr/bin/env python
"""Launch the image-sized operators twice at B=8, 544x960 for `ncu --set full` (no timing here)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from activezero_b200 import ops  # noqa: E402

DEV = "cuda:0"
B, H, W = 8, 544, 960
torch.manual_seed(1)
pL = (torch.rand(B, 1, H, W, device=DEV) > 0.5).float()
pR = (torch.rand(B, 1, H, W, device=DEV) > 0.5).float()
d = torch.rand(B, 1, H, W, device=DEV) * 64
mask = torch.rand(B, 1, H, W, device=DEV) > 0.2
di = (torch.rand(B, 1, H, W, device=DEV) * 64).int()
fr = torch.randint(0, 255, (B, 7, H, W), dtype=torch.uint8, device=DEV)
for _ in range(2):
    ops.warp(pR, d)
    ops.reproj_loss(pL, pR, d, mask, ps=1)
    ops.scatter_warp(d, di, check_sign=False)
    ops.local_contrast_norm(pL, 9)
    ops.temporal_ir_pattern(fr)
    torch.cuda.synchronize()
print("done")

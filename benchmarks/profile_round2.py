#!/usr/bin/env python
"""Launch the kernels that changed in round 2 at BASELINE config-2 sizes, for one `ncu --set full` capture
(profiles/README.md has the command): concat forward on the LSU path and on the bulk-store (TMA) path, the packed
soft-argmin forward, the fused upsample + soft-argmin forward, gwc forward / backward (both backward kernels), the
patch loss + Fold kernel and the tcgen05 first convolution.  No timing here."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from activezero_b200 import ops  # noqa: E402

B, C, Hq, Wq, D, PS = 8, 32, 136, 240, 192, 11
H, W, Dq = 4 * Hq, 4 * Wq, D // 4
torch.manual_seed(0)
L, R = torch.randn(B, C, Hq, Wq, device="cuda"), torch.randn(B, C, Hq, Wq, device="cuda")
low = torch.randn(B, 1, Dq, Hq, Wq, device="cuda") * 4
cost = torch.empty(B, D, H, W, device="cuda")
for b in range(B):
    cost[b] = torch.nn.functional.interpolate(low[b:b + 1], size=(D, H, W), mode="trilinear", align_corners=False)[0, 0]
pL = (torch.rand(B, 1, H, W, device="cuda") > 0.5).float()
pR = (torch.rand(B, 1, H, W, device="cuda") > 0.5).float()
mask = torch.rand(B, 1, H, W, device="cuda") > 0.2
w = torch.randn(32, 64, 3, 3, 3, device="cuda") * 0.05
wp = ops.pack_volume_conv_weight(w)
with torch.no_grad():
    for it in range(2):
        for impl in ("0", "1"):
            os.environ["AZ_CONCAT_FWD"] = impl
            v = ops.build_concat_volume(L, R, Dq)
            del v
        os.environ["AZ_CONCAT_FWD"] = "0"
        d = ops.soft_argmin(cost)
        d2 = ops.upsample_soft_argmin(low, (D, H, W))
        ops.reproj_loss(pL, pR, d, mask, ps=PS, want_warped=True)
        gv = ops.build_gwc_volume(L, R, Dq, 8)
        ops.volume_conv0(L[:2], R[:2], wp, Dq)
        torch.cuda.synchronize()
g = torch.randn(B, 8, Dq, Hq, Wq, device="cuda")
for impl in ("0", "1"):
    os.environ["AZ_GWC_BWD"] = impl
    Lg, Rg = L.clone().requires_grad_(True), R.clone().requires_grad_(True)
    ops.build_gwc_volume(Lg, Rg, Dq, 8).backward(g)
torch.cuda.synchronize()
print("profiled launches done")

#!/usr/bin/env python
"""Launch every kernel of libaz_stereo.so once (after one warm-up) at BASELINE config-2 per-pair
sizes, for `ncu --set full` (profiles/README.md has the command).  No timing here."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from activezero_b200 import ops  # noqa: E402

DEV = "cuda:0"
B, C, Hq, Wq, D, PS = 2, 32, 136, 240, 192, 11
H, W, Dq = 4 * Hq, 4 * Wq, D // 4


def main():
    torch.manual_seed(0)
    L = torch.randn(B, C, Hq, Wq, device=DEV, requires_grad=True)
    R = torch.randn(B, C, Hq, Wq, device=DEV, requires_grad=True)
    cost = (torch.randn(B, D, H, W, device=DEV) * 4).requires_grad_(True)
    low = (torch.randn(B, 1, Dq, Hq, Wq, device=DEV) * 4).requires_grad_(True)
    pL = (torch.rand(B, 1, H, W, device=DEV) > 0.5).float()
    pR = (torch.rand(B, 1, H, W, device=DEV) > 0.5).float()
    mask = torch.rand(B, 1, H, W, device=DEV) > 0.2
    g1 = torch.randn(B, 1, H, W, device=DEV)
    di = (torch.rand(B, 1, H, W, device=DEV) * 64).int()
    frames = torch.randint(0, 255, (B, 7, H, W), dtype=torch.uint8, device=DEV)
    for it in range(2):
        vol = ops.build_concat_volume(L, R, Dq)
        vol.backward(torch.ones_like(vol))
        volc = ops.build_concat_volume(L, R, Dq, channels_last=True)  # channels_last_3d order, SURVEY 8f rank 2
        volc.backward(torch.ones_like(volc))
        gv = ops.build_gwc_volume(L, R, Dq, 8)
        gv.backward(torch.ones_like(gv))
        disp = ops.soft_argmin(cost)
        disp.backward(g1)
        d2 = ops.upsample_soft_argmin(low, (D, H, W))
        d2.backward(g1)
        dd = disp.detach().requires_grad_(True)
        loss, _ = ops.reproj_loss(pL, pR, dd, mask, ps=PS)
        loss.backward()
        dd.grad = None
        loss, _ = ops.reproj_loss(pL, pR, dd, mask, ps=PS, want_warped=True)  # loss + Fold image in one pass
        loss.backward()
        ops.reproj_loss(pL, pR, dd.detach(), mask, ps=1, want_warped=True)
        ops.patch_fold(pR, dd.detach(), PS)
        w = ops.warp(pR.requires_grad_(False), dd)
        w.backward(torch.ones_like(w))
        ops.scatter_warp(disp.detach(), di, check_sign=False)
        ops.temporal_ir_pattern(frames)
        ops.local_contrast_norm(pL, 9)
        for t in (L, R, cost, low):
            t.grad = None
        torch.cuda.synchronize()
    print("profiled launches done")


if __name__ == "__main__":
    main()

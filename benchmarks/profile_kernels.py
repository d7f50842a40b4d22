#!/usr/bin/env python
"""Launch every kernel of libaz_stereo.so once (after one warm-up; `--iters 1` skips the warm-up pass) at BASELINE
config-2 per-pair sizes, for `ncu --set full` (profiles/README.md has the command).  No timing here."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from activezero_b200 import ops  # noqa: E402
from activezero_b200.utils import reprojection as rp  # noqa: E402

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
    d2x = torch.rand(B, 1, 2 * H, 2 * W, device=DEV) * 60  # the trainer's double-resolution right-view disparity
    depth = torch.rand(B, 1, H, W, device=DEV) + 0.5
    focal, base = torch.full((B,), 446.0, device=DEV), torch.full((B,), 0.055, device=DEV)
    ir = torch.randint(0, 255, (B, H, W), dtype=torch.uint8, device=DEV)
    no_ir = torch.randint(0, 255, (B, H, W), dtype=torch.uint8, device=DEV)
    w0 = torch.randn(32, 64, 3, 3, 3, device=DEV) * 0.05
    iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 2
    for it in range(iters):
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
        # round-2 entry points: image gradient of the warp, the GT chain, a9's rescaling + multi-scale loss,
        # error metrics, simulated-IR pattern, the first aggregation convolution on the implicit volume
        im = pR.clone().requires_grad_(True)
        ops.warp(im, dd.detach()).backward(g1)
        ops.scatter_warp_gt(d2x, 192.0, check_sign=False)
        ms = rp.get_reprojection_error_diff_ratio(pL, pR, dd, mask)
        (ms[0] if isinstance(ms, (tuple, list)) else ms).backward()
        dd.grad = None
        ops.error_metric_sums(disp.detach() + 1, depth, disp.detach(), mask, focal_length=focal, baseline=base)
        ops.sim_ir_pattern(ir, no_ir)
        ops.volume_conv0(L, R, ops.pack_volume_conv_weight(w0), Dq)
        for t in (L, R, cost, low):
            t.grad = None
        torch.cuda.synchronize()
    print("profiled launches done")


if __name__ == "__main__":
    main()

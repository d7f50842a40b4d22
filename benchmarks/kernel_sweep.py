#!/usr/bin/env python
"""Per-kernel sweep of the hot path on one B200 (BASELINE config 5): every forward
and backward kernel at D in {96,192,288} x (H/4,W/4) in {64x128, 136x240, 272x480},
achieved GB/s on the ALGORITHMIC bytes against the measured HBM peak.

    python benchmarks/kernel_sweep.py [--quick] [--out profiles/rN_kernel_sweep.json]

Timing: CUDA events on the launch stream, 3 warm-ups, >= 10 timed launches; batch
chosen so that every launch streams >= ~1.5 GB (>> 126 MB L2) where the op is
volume-sized, and a 256 MB L2 flush write between launches for the image-sized ops.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from activezero_b200 import _lib, ops  # noqa: E402

DEV = "cuda:0"
C, G, PS = 32, 8, 11


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


_flush = None


def time_ms(fn, iters=10, warm=3, flush=False, graph=False):
    """Median device time of fn() in ms.  graph=True captures fn into a CUDA graph first and times replays:
    the image-sized ops run 10-60 us, less than the Python side of an operator call (allocations, ctypes), so
    timing the eager call would measure the host."""
    global _flush
    if flush and _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    for _ in range(warm):
        fn()
    run = fn
    if graph:
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        run = g.replay
        run()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    torch.cuda.synchronize()
    for a, b in evs:
        if flush:
            _flush.fill_(1)
        a.record()
        run()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def sweep(quick):
    pk = peak()
    rows = []

    def add(kernel, cfg, ms, nbytes):
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append({"kernel": kernel, **cfg, "ms": round(ms, 4), "algo_MB": round(nbytes / 1e6, 2),
                     "GBps": round(gbs, 1), "frac_of_measured_peak": round(gbs / pk, 3)})
        print(json.dumps(rows[-1]), flush=True)

    sizes = [(64, 128), (136, 240), (272, 480)]
    disps = [96, 192, 288]
    if quick:
        sizes, disps = [(136, 240)], [192]
    for (Hq, Wq) in sizes:
        for D in disps:
            Dq, H, W = D // 4, 4 * Hq, 4 * Wq
            vol_bytes = 4 * 2 * C * Dq * Hq * Wq
            B = max(1, min(8, int(1.6e9 // vol_bytes)))
            cfg = {"Hq": Hq, "Wq": Wq, "D": D, "B": B, "int64_index": B * 2 * C * Dq * Hq * Wq >= 2 ** 31}
            torch.manual_seed(0)
            L = torch.randn(B, C, Hq, Wq, device=DEV)
            R = torch.randn(B, C, Hq, Wq, device=DEV)
            feat_b = 4 * B * C * Hq * Wq
            # concat
            vol = ops.build_concat_volume(L, R, Dq)
            add("concat_volume_fwd", cfg, time_ms(lambda: ops.build_concat_volume(L, R, Dq)), 2 * feat_b + vol.numel() * 4)
            gL, gR = torch.empty_like(L), torch.empty_like(R)
            add("concat_volume_bwd", cfg, time_ms(lambda: _lib.call(
                "az_concat_volume_bwd", ops._ptr(vol), ops._ptr(gL), ops._ptr(gR), B, C, Hq, Wq, Dq, ops._stream())),
                2 * feat_b + vol.numel() * 4)
            # the same volume / gradient in channels_last_3d memory order
            add("concat_volume_fwd_ndhwc", cfg, time_ms(lambda: ops.build_concat_volume(L, R, Dq, channels_last=True)),
                2 * feat_b + vol.numel() * 4)
            add("concat_volume_bwd_ndhwc", cfg, time_ms(lambda: _lib.call(
                "az_concat_volume_bwd_ndhwc", ops._ptr(vol), ops._ptr(gL), ops._ptr(gR), B, C, Hq, Wq, Dq, ops._stream())),
                2 * feat_b + vol.numel() * 4)
            del vol
            # gwc (the volume is 8x smaller: batch it up to stay >> L2)
            gv = ops.build_gwc_volume(L, R, Dq, G)
            add("gwc_volume_fwd", cfg, time_ms(lambda: ops.build_gwc_volume(L, R, Dq, G), flush=True), 2 * feat_b + gv.numel() * 4)
            add("gwc_volume_bwd", cfg, time_ms(lambda: _lib.call(
                "az_gwc_volume_bwd", ops._ptr(gv), ops._ptr(L), ops._ptr(R), ops._ptr(gL), ops._ptr(gR), B, C, Hq, Wq,
                Dq, G, ops._stream()), flush=True), 4 * feat_b + gv.numel() * 4)
            del gv
            # soft-argmin on the full-resolution logits
            cost_bytes = 4 * D * H * W
            Bs = max(1, min(8, int(1.6e9 // cost_bytes)))
            cfg_s = {**cfg, "B": Bs, "int64_index": Bs * D * H * W >= 2 ** 31}
            cost = torch.randn(Bs, D, H, W, device=DEV) * 4
            disp, lse = torch.empty(Bs, 1, H, W, device=DEV), torch.empty(Bs, 2, H, W, device=DEV)
            add("soft_argmin_fwd", cfg_s, time_ms(lambda: _lib.call(
                "az_soft_argmin_fwd", ops._ptr(cost), ops._ptr(disp), ops._ptr(lse), Bs, D, H, W, ops._stream())),
                4 * Bs * (D * H * W + H * W))
            g = torch.randn(Bs, 1, H, W, device=DEV)
            gcost = torch.empty_like(cost)
            add("soft_argmin_bwd", cfg_s, time_ms(lambda: _lib.call(
                "az_soft_argmin_bwd", ops._ptr(cost), ops._ptr(disp), ops._ptr(lse), ops._ptr(g), ops._ptr(gcost), Bs, D,
                H, W, ops._stream())), 4 * Bs * (2 * D * H * W + 2 * H * W))
            del cost, gcost
            # fused trilinear upsample + soft-argmin: reads the low-res logits only
            low = torch.randn(Bs, 1, Dq, Hq, Wq, device=DEV) * 4
            add("upsample_soft_argmin_fwd", cfg_s, time_ms(lambda: ops.upsample_soft_argmin(low, (D, H, W)), flush=True),
                4 * Bs * (Dq * Hq * Wq + H * W))
            lowg = low.clone().requires_grad_(True)

            def fused_fb():
                lowg.grad = None
                ops.upsample_soft_argmin(lowg, (D, H, W)).backward(g)

            add("upsample_soft_argmin_fwd+bwd", cfg_s, time_ms(fused_fb, flush=True),
                4 * Bs * (2 * Dq * Hq * Wq + 4 * H * W))
            # context: the reference dataflow for one head on the GPU (stock torch trilinear
            # interpolate materialising the logits, then this repo's soft-argmin)
            import torch.nn.functional as F

            def unfused_f():
                return ops.soft_argmin(torch.squeeze(F.interpolate(low, (D, H, W), mode="trilinear", align_corners=False), 1))

            def unfused_fb():
                lowg.grad = None
                ops.soft_argmin(torch.squeeze(F.interpolate(lowg, (D, H, W), mode="trilinear", align_corners=False), 1)).backward(g)

            add("context:torch_interpolate+soft_argmin_fwd", cfg_s, time_ms(unfused_f), 4 * Bs * (2 * D * H * W + H * W))
            add("context:torch_interpolate+soft_argmin_fwd+bwd", cfg_s, time_ms(unfused_fb), 4 * Bs * (6 * D * H * W))
            del low, lowg
    # image-sized ops at the two frame sizes (not D dependent)
    for (H, W) in ([(544, 960)] if quick else [(256, 512), (544, 960), (1088, 1920)]):
        B = 8
        cfg = {"H": H, "W": W, "B": B, "ps": PS}
        torch.manual_seed(1)
        pL = (torch.rand(B, 1, H, W, device=DEV) > 0.5).float()
        pR = (torch.rand(B, 1, H, W, device=DEV) > 0.5).float()
        d = torch.rand(B, 1, H, W, device=DEV) * 64
        mask = torch.rand(B, 1, H, W, device=DEV) > 0.2
        hw = B * H * W
        add("warp_fwd", cfg, time_ms(lambda: ops.warp(pR, d), flush=True, graph=True), 4 * 3 * hw)
        dg = d.clone().requires_grad_(True)

        def patch_fb():
            dg.grad = None
            loss, _ = ops.reproj_loss(pL, pR, dg, mask, ps=PS)
            loss.backward()

        add("reproj_ps1_loss_fwd", cfg, time_ms(lambda: ops.reproj_loss(pL, pR, d, mask, ps=1), flush=True, graph=True), 4 * 3 * hw + hw)
        add("reproj_patch_loss_fwd", cfg, time_ms(lambda: ops.reproj_loss(pL, pR, d, mask, ps=PS), flush=True, graph=True), 4 * 3 * hw + hw)
        add("reproj_patch_loss_fwd+bwd", cfg, time_ms(patch_fb, flush=True), 4 * 3 * hw + hw + 3 * 4 * hw)
        add("patch_fold", cfg, time_ms(lambda: ops.patch_fold(pR, d, PS), flush=True, graph=True), 4 * 3 * hw)
        add("reproj_patch_loss+fold_fwd", cfg, time_ms(lambda: ops.reproj_loss(pL, pR, d, mask, ps=PS, want_warped=True),
                                                       flush=True, graph=True), 4 * 4 * hw + hw)
        # the same with a SMOOTH disparity field (what a trained network predicts): the per-pixel random
        # disparities above make ~half of the shared-memory wavefronts bank conflicts
        yy, xx = torch.meshgrid(torch.arange(H, device=DEV), torch.arange(W, device=DEV), indexing="ij")
        ds = (20.0 + 30.0 * torch.sin(xx / 97.0) * torch.cos(yy / 61.0) + 0.013 * xx).view(1, 1, H, W).expand(B, 1, H, W).contiguous()
        cfg_s = {**cfg, "disp": "smooth"}
        add("reproj_patch_loss_fwd", cfg_s, time_ms(lambda: ops.reproj_loss(pL, pR, ds, mask, ps=PS), flush=True, graph=True), 4 * 3 * hw + hw)
        add("patch_fold", cfg_s, time_ms(lambda: ops.patch_fold(pR, ds, PS), flush=True, graph=True), 4 * 3 * hw)
        add("reproj_patch_loss+fold_fwd", cfg_s, time_ms(lambda: ops.reproj_loss(pL, pR, ds, mask, ps=PS, want_warped=True),
                                                         flush=True, graph=True), 4 * 4 * hw + hw)
        di = (torch.rand(B, 1, H, W, device=DEV) * 64).int()
        add("scatter_warp", cfg, time_ms(lambda: ops.scatter_warp(d, di, check_sign=False), flush=True, graph=True), 4 * 3 * hw)
        # the trainer's GT chain (train.py:255-272) fused, against the reference's six calls on this library's scatter warp
        d2x = torch.rand(B, 1, 2 * H, 2 * W, device=DEV) * 100

        def gt_unfused():
            r = torch.nn.functional.interpolate(d2x, scale_factor=0.5, mode="nearest", recompute_scale_factor=False)
            o = ops.scatter_warp(r, r.type(torch.int), check_sign=False)
            return o, (o < 192) * (o > 0)

        nb_gt = hw * (4 + 4 + 1)  # the quarter of the 2x image that is read + disparity out + mask out
        add("gt_chain_fused", cfg, time_ms(lambda: ops.scatter_warp_gt(d2x, 192.0, check_sign=False), flush=True, graph=True), nb_gt)
        add("context:gt_chain_unfused(torch interpolate+cast+az scatter+mask)", cfg, time_ms(gt_unfused, flush=True, graph=True), nb_gt)
        add("local_contrast_norm", cfg, time_ms(lambda: ops.local_contrast_norm(pL, 9), flush=True, graph=True), 4 * 3 * hw)
        T = 7
        fr = torch.randint(0, 255, (B, T, H, W), dtype=torch.uint8, device=DEV)
        add("temporal_ir", {**cfg, "T": T}, time_ms(lambda: ops.temporal_ir_pattern(fr), flush=True, graph=True), T * hw + 4 * hw)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rows = sweep(args.quick)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump({"peak_gbs": peak(), "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

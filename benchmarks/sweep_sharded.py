#!/usr/bin/env python
"""BASELINE config 5, sharded: the cost-volume / regression / warp kernels at D in {96,192,288} x (H/4,W/4) in
{64x128, 136x240, 272x480}, a job of B_TOTAL stereo pairs split evenly over the N ranks of one box (strong
scaling: every rank runs B_TOTAL / N pairs, no collective on the data path), against the HBM roofline.

    python benchmarks/sweep_sharded.py                                   (1 GPU)
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/sweep_sharded.py [--out f.json]

Timing: CUDA events on every rank, 3 warm-ups, 5 timed launches bracketed by a barrier + device sync; the job's
time is the MAX over ranks; GB/s = algorithmic bytes of the WHOLE job / that time, fraction = per-GPU share of the
measured HBM peak.  A 256 MB flush write precedes every launch whose working set could sit in L2."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from activezero_b200 import _lib, dist_util, ops  # noqa: E402
from activezero_b200.ops import _ptr, _stream  # noqa: E402

C, G, PS, B_TOTAL = 32, 8, 11, 8


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    rank, world, local = dist_util.env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_util.init_from_env("nccl", dev)
    first, nb = dist_util.shard_pairs(B_TOTAL, rank, world)
    assert nb >= 1, "more ranks than pairs"
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    pk = peak()
    rows = []

    def timed(fn, iters=5, warm=3):
        for _ in range(warm):
            fn()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        dist_util.barrier()
        torch.cuda.synchronize()
        for a, b in evs:
            flush.fill_(1)
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        (ms,) = dist_util.max_over_ranks([ts[len(ts) // 2]], dev)
        return ms

    def add(kernel, cfg, ms, bytes_per_pair):
        total = bytes_per_pair * B_TOTAL
        gbs = total / (ms * 1e-3) / 1e9
        r = {"kernel": kernel, **cfg, "n_gpus": world, "pairs_total": B_TOTAL, "pairs_per_gpu": nb, "ms": round(ms, 4),
             "job_GBps": round(gbs, 1), "per_gpu_frac_of_measured_peak": round(gbs / world / pk, 3)}
        rows.append(r)
        if rank == 0:
            print(json.dumps(r), flush=True)

    sizes = [(136, 240)] if args.quick else [(64, 128), (136, 240), (272, 480)]
    disps = [192] if args.quick else [96, 192, 288]
    torch.manual_seed(first)
    for (Hq, Wq) in sizes:
        for D in disps:
            Dq, H, W = D // 4, 4 * Hq, 4 * Wq
            cfg = {"Hq": Hq, "Wq": Wq, "D": D}
            L, R = torch.randn(nb, C, Hq, Wq, device=dev), torch.randn(nb, C, Hq, Wq, device=dev)
            feat, vol_b = 4 * 2 * C * Hq * Wq, 4 * 2 * C * Dq * Hq * Wq
            vol = ops.build_concat_volume(L, R, Dq)
            add("concat_volume_fwd", cfg, timed(lambda: ops.build_concat_volume(L, R, Dq)), feat + vol_b)
            gL, gR = torch.empty_like(L), torch.empty_like(R)
            add("concat_volume_bwd", cfg, timed(lambda: _lib.call("az_concat_volume_bwd", _ptr(vol), _ptr(gL), _ptr(gR), nb, C,
                                                                  Hq, Wq, Dq, _stream())), feat + vol_b)
            del vol
            gv = ops.build_gwc_volume(L, R, Dq, G)
            add("gwc_volume_fwd", cfg, timed(lambda: ops.build_gwc_volume(L, R, Dq, G)), feat + 4 * G * Dq * Hq * Wq)
            add("gwc_volume_bwd", cfg, timed(lambda: _lib.call("az_gwc_volume_bwd", _ptr(gv), _ptr(L), _ptr(R), _ptr(gL), _ptr(gR),
                                                               nb, C, Hq, Wq, Dq, G, _stream())), 2 * feat + 4 * G * Dq * Hq * Wq)
            del gv
            cost = torch.randn(nb, D, H, W, device=dev) * 4
            disp, lse = torch.empty(nb, 1, H, W, device=dev), torch.empty(nb, 2, H, W, device=dev)
            add("soft_argmin_fwd", cfg, timed(lambda: _lib.call("az_soft_argmin_fwd", _ptr(cost), _ptr(disp), _ptr(lse), nb, D, H, W,
                                                                _stream())), 4 * (D * H * W + 3 * H * W))
            g = torch.randn(nb, 1, H, W, device=dev)
            gcost = torch.empty_like(cost)
            add("soft_argmin_bwd", cfg, timed(lambda: _lib.call("az_soft_argmin_bwd", _ptr(cost), _ptr(disp), _ptr(lse), _ptr(g),
                                                                _ptr(gcost), nb, D, H, W, _stream())), 4 * (2 * D * H * W + 4 * H * W))
            del cost, gcost
            low = torch.randn(nb, 1, Dq, Hq, Wq, device=dev) * 4
            add("upsample_soft_argmin_fwd", cfg, timed(lambda: ops.upsample_soft_argmin(low, (D, H, W))), 4 * (Dq * Hq * Wq + H * W))
        H, W = 4 * Hq, 4 * Wq
        cfg = {"H": H, "W": W, "ps": PS}
        pL = (torch.rand(nb, 1, H, W, device=dev) > 0.5).float()
        pR = (torch.rand(nb, 1, H, W, device=dev) > 0.5).float()
        d = torch.rand(nb, 1, H, W, device=dev) * 64
        mask = torch.rand(nb, 1, H, W, device=dev) > 0.2
        hw = H * W
        add("warp_fwd", cfg, timed(lambda: ops.warp(pR, d)), 4 * 3 * hw)
        add("reproj_ps1_loss_fwd", cfg, timed(lambda: ops.reproj_loss(pL, pR, d, mask, ps=1)), 4 * 3 * hw + hw)
        add("reproj_patch_loss+fold_fwd", cfg, timed(lambda: ops.reproj_loss(pL, pR, d, mask, ps=PS, want_warped=True)), 4 * 4 * hw + hw)
        di = (torch.rand(nb, 1, H, W, device=dev) * 64).int()
        add("scatter_warp", cfg, timed(lambda: ops.scatter_warp(d, di, check_sign=False)), 4 * 3 * hw)
    if rank == 0 and args.out:
        json.dump({"n_gpus": world, "peak_GBps_per_gpu": pk, "pairs_total": B_TOTAL, "rows": rows}, open(args.out, "w"), indent=1)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Config-2 step with the GPU split into two SM partitions (CUDA green contexts) and the step software-pipelined
across batches: the compute-bound patch loss of batch i runs on k SMs WHILE the two HBM-bound kernels (soft-argmin,
concat volume) of batch i+1 stream on the other 148-k SMs.

Why: in the bench step the three kernels run one after the other (1.27 ms): the patch kernel (0.32 ms, shared-memory
bound, one CTA of 15 warps x 124 registers per SM) leaves HBM idle, the two volume-sized kernels (0.96 ms) leave the
SMs' arithmetic idle, and they cannot share an SM (no guest CTA becomes resident beside the patch CTA).
Giving each side its own SMs lets both run all the time; the patch kernel sizes its bands for its partition
(AZ_PATCH_SMS).  Batches are independent (inference), so batch i's loss may trail batch i+1's volume by one step; all
work of the K timed steps, the last loss included, lies inside the timed region.

Informational experiment (one B200); prints one JSON line per configuration.  Not the bench's `value`.

    python benchmarks/sm_partition_pipeline.py [--steps 20] [--warmup 3] [--sms 40,48,56,64] [--out FILE]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from activezero_b200 import ops  # noqa: E402

DEV = torch.device("cuda", 0)
B, C, HQ, WQ, D, PS = 8, 32, 136, 240, 192, 11
H, W, DQ = 4 * HQ, 4 * WQ, D // 4


def make_inputs():
    g = torch.Generator(device=DEV).manual_seed(1000)
    L = torch.randn(B, C, HQ, WQ, generator=g, device=DEV)
    R = torch.randn(B, C, HQ, WQ, generator=g, device=DEV)
    cost = torch.empty(B, D, H, W, device=DEV)
    for b in range(B):  # the logits PSMNet feeds to softmax: trilinear upsample of low-resolution logits (psmnet.py:186-197)
        low = torch.randn(1, 1, DQ, HQ, WQ, generator=g, device=DEV) * 4.0
        cost[b] = torch.nn.functional.interpolate(low, size=(D, H, W), mode="trilinear", align_corners=False)[0, 0]
    pat_L = (torch.rand(B, 1, H, W, generator=g, device=DEV) > 0.5).float()
    pat_R = (torch.rand(B, 1, H, W, generator=g, device=DEV) > 0.5).float()
    mask = torch.rand(B, 1, H, W, generator=g, device=DEV) > 0.2
    return L, R, cost, pat_L, pat_R, mask


def green_streams(k: int):
    """Two streams on disjoint SM partitions: (stream with >= k SMs, stream with the rest, their SM counts)."""
    from cuda.bindings import driver as cu

    def ck(res, what):
        if res[0] != cu.CUresult.CUDA_SUCCESS:
            raise RuntimeError(f"{what}: {res[0]}")
        return res[1] if len(res) == 2 else res[1:]

    dev = ck(cu.cuDeviceGet(DEV.index), "cuDeviceGet")
    res = ck(cu.cuDeviceGetDevResource(dev, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM), "cuDeviceGetDevResource")
    groups, _, rest = ck(cu.cuDevSmResourceSplitByCount(1, res, 0, k), "cuDevSmResourceSplitByCount")
    streams, counts, keep = [], [], []
    for r in (groups[0], rest):
        desc = ck(cu.cuDevResourceGenerateDesc([r], 1), "cuDevResourceGenerateDesc")
        gctx = ck(cu.cuGreenCtxCreate(desc, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM), "cuGreenCtxCreate")
        st = ck(cu.cuGreenCtxStreamCreate(gctx, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0), "cuGreenCtxStreamCreate")
        streams.append(torch.cuda.get_stream_from_external(int(st), DEV))
        counts.append(int(r.sm.smCount))
        keep.append(gctx)
    return streams[0], streams[1], counts[0], counts[1], keep


def sequential(inp, steps, warm):
    L, R, cost, pL, pR, m = inp
    out = None
    for _ in range(warm):
        disp = ops.soft_argmin(cost)
        vol = ops.build_concat_volume(L, R, DQ)
        out = ops.reproj_loss(pL, pR, disp, m, ps=PS, sign=-1.0, want_warped=True)
        del vol
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        disp = ops.soft_argmin(cost)
        vol = ops.build_concat_volume(L, R, DQ)
        out = ops.reproj_loss(pL, pR, disp, m, ps=PS, sign=-1.0, want_warped=True)
        del vol
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def pipelined(inp, s_patch, s_hbm, steps, warm):
    """HBM kernels of batch i on s_hbm; patch loss of batch i on s_patch as soon as its disparity exists, i.e. under
    the concat volume of batch i and the soft-argmin of batch i+1."""
    L, R, cost, pL, pR, m = inp
    out = None

    def run(n):
        nonlocal out
        for _ in range(n):
            with torch.cuda.stream(s_hbm):
                disp = ops.soft_argmin(cost)
                ready = torch.cuda.Event()
                ready.record(s_hbm)
                vol = ops.build_concat_volume(L, R, DQ)
                del vol  # returns to the allocator's pool of s_hbm; reused by the next step on the same stream
            # allocated on s_hbm, read on s_patch: the caching allocator recycles the block only after s_patch has
            # passed this point, so the loop allocates nothing new after its first steps (a first form that kept every
            # output alive needed two fresh cudaMalloc calls per step: 1.13 ms per step in this script's small process,
            # 4.15 ms inside bench.py, whose process already maps tens of GB)
            disp.record_stream(s_patch)
            with torch.cuda.stream(s_patch):
                s_patch.wait_event(ready)
                out = ops.reproj_loss(pL, pR, disp, m, ps=PS, sign=-1.0, want_warped=True)

    cur = torch.cuda.current_stream()
    s_hbm.wait_stream(cur)
    s_patch.wait_stream(cur)
    run(warm)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    eh, ep = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s_hbm)
    run(steps)
    eh.record(s_hbm)
    ep.record(s_patch)
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(eh), e0.elapsed_time(ep)) / steps  # every kernel of the K steps, the last loss included
    return ms, out, e0.elapsed_time(eh) / steps, e0.elapsed_time(ep) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--sms", default="40,48,56,64")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rows = []

    def emit(r):
        rows.append(r)
        print(json.dumps(r), flush=True)

    with torch.no_grad():
        inp = make_inputs()
        torch.cuda.synchronize()
        ms, ref = sequential(inp, args.steps, args.warmup)
        ref_loss, ref_vis = float(ref[0]), ref[1].clone()
        emit({"config": "sequential, one stream, all 148 SMs (the bench step, eager launches)", "ms_per_step": round(ms, 4),
              "pairs_per_s": round(B / ms * 1e3, 1), "loss": ref_loss})

        def check(out):
            return {"loss": float(out[0]), "loss_rel_diff": abs(float(out[0]) - ref_loss) / abs(ref_loss),
                    "fold_image_max_abs_diff": float((out[1] - ref_vis).abs().max())}

        # two ordinary streams, no partition: what the block scheduler does on its own
        for sms in (148, 48):
            os.environ["AZ_PATCH_SMS"] = str(sms)
            try:
                sa, sb = torch.cuda.Stream(device=DEV, priority=-1), torch.cuda.Stream(device=DEV)
                ms, out, mh, mp = pipelined(inp, sa, sb, args.steps, args.warmup)
                emit({"config": f"two plain streams (patch stream high priority), patch bands sized for {sms} SMs",
                      "ms_per_step": round(ms, 4), "pairs_per_s": round(B / ms * 1e3, 1), "hbm_stream_ms": round(mh, 4),
                      "patch_stream_ms": round(mp, 4), **check(out)})
            except Exception as e:  # pragma: no cover
                emit({"config": f"two plain streams, {sms}", "error": repr(e)})
        for k in [int(v) for v in args.sms.split(",") if v]:
            try:
                sa, sb, na, nb, keep = green_streams(k)
                os.environ["AZ_PATCH_SMS"] = str(na)
                ms, out, mh, mp = pipelined(inp, sa, sb, args.steps, args.warmup)
                emit({"config": f"green contexts: patch loss on {na} SMs, soft-argmin + concat volume on {nb} SMs",
                      "patch_sms": na, "hbm_sms": nb, "ms_per_step": round(ms, 4), "pairs_per_s": round(B / ms * 1e3, 1),
                      "hbm_stream_ms": round(mh, 4), "patch_stream_ms": round(mp, 4), **check(out)})
                # the two sides alone on their partitions (no overlap): how much each loses to its smaller SM count
                L, R, cost, pL, pR, m = inp
                disp = ops.soft_argmin(cost)
                torch.cuda.synchronize()
                for name, st, fn in (("hbm kernels alone on their partition", sb,
                                      lambda: (ops.soft_argmin(cost), ops.build_concat_volume(L, R, DQ))),
                                     ("patch loss alone on its partition", sa,
                                      lambda: ops.reproj_loss(pL, pR, disp, m, ps=PS, sign=-1.0, want_warped=True))):
                    with torch.cuda.stream(st):
                        for _ in range(2):
                            fn()
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record(st)
                        for _ in range(5):
                            fn()
                        b.record(st)
                    torch.cuda.synchronize()
                    emit({"config": f"  {name} ({nb if st is sb else na} SMs)", "ms": round(a.elapsed_time(b) / 5, 4)})
                del keep
            except Exception as e:  # pragma: no cover
                emit({"config": f"green contexts, k={k}", "error": repr(e)})
        os.environ.pop("AZ_PATCH_SMS", None)
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"device": torch.cuda.get_device_name(0), "workload": f"B={B} {H}x{W} D={D} ps={PS}", "steps": args.steps,
                       "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""A/B timing of the kernel variants behind the AZ_* tuning knobs (csrc/common.cuh `tuning()`), at BASELINE
config 2 sizes (B = 8, 544x960, D = 192) on one B200.  CUDA events, median of 10 after 3 warm-ups; every launch
streams >> 126 MB (L2) or is preceded by a 256 MB flush write.

    python benchmarks/variant_bench.py [--out profiles/r2_variants.json] [--only sa,concat,gwc,usa,patch]
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
from benchmarks.kernel_sweep import peak, time_ms  # noqa: E402

DEV = "cuda:0"
B, C, Hq, Wq, D, G, PS = 8, 32, 136, 240, 192, 8, 11
H, W, Dq = 4 * Hq, 4 * Wq, D // 4


class env:
    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default="sa,concat,gwc,usa,patch,vconv")
    args = ap.parse_args()
    only = set(args.only.split(","))
    pk = peak()
    rows = []

    def add(kernel, variant, ms, nbytes, extra=None):
        r = {"kernel": kernel, "variant": variant, "ms": round(ms, 4), "algo_MB": round(nbytes / 1e6, 1),
             "GBps": round(nbytes / ms / 1e6, 1), "frac_of_measured_peak": round(nbytes / ms / 1e6 / pk, 3)}
        if extra:
            r.update(extra)
        rows.append(r)
        print(json.dumps(r), flush=True)

    torch.manual_seed(0)
    with torch.no_grad():
        if "sa" in only:
            low = torch.randn(B, 1, Dq, Hq, Wq, device=DEV) * 4
            cost = torch.empty(B, D, H, W, device=DEV)
            for b in range(B):
                cost[b] = torch.nn.functional.interpolate(low[b:b + 1], size=(D, H, W), mode="trilinear", align_corners=False)[0, 0]
            nb = 4 * (D * H * W + H * W) * B
            ref = None
            for v in (0, 1, 2, 3, 4):
                with env(AZ_SA_FWD=v):
                    ms = time_ms(lambda: ops.soft_argmin(cost))
                    out = ops.soft_argmin(cost)
                ref = out if ref is None else ref
                add("soft_argmin_fwd", f"AZ_SA_FWD={v}", ms, nb, {"max_abs_diff_vs_v0": float((out - ref).abs().max())})
            noise = torch.randn(B, D, H, W, device=DEV) * 4  # per-pixel random logits (max moves ~4x per pixel)
            for v in (0, 2):
                with env(AZ_SA_FWD=v):
                    add("soft_argmin_fwd(random logits)", f"AZ_SA_FWD={v}", time_ms(lambda: ops.soft_argmin(noise)), nb)
            del cost, noise
        if "usa" in only:
            low = torch.randn(B, 1, Dq, Hq, Wq, device=DEV) * 4
            nb = 4 * (Dq * Hq * Wq + H * W) * B
            ref = None
            for v in (0, 1):
                with env(AZ_USA_FWD=v):
                    ms = time_ms(lambda: ops.upsample_soft_argmin(low, (D, H, W)), flush=True)
                    out = ops.upsample_soft_argmin(low, (D, H, W))
                ref = out if ref is None else ref
                ex2 = (4 if v == 0 else 2) * (Dq - 1) * H * W * B
                add("upsample_soft_argmin_fwd", f"AZ_USA_FWD={v}", ms, nb,
                    {"max_abs_diff_vs_v0": float((out - ref).abs().max()), "Tex2_per_s": round(ex2 / ms / 1e9, 3)})
        if "concat" in only:
            L, R = torch.randn(B, C, Hq, Wq, device=DEV), torch.randn(B, C, Hq, Wq, device=DEV)
            nb = 4 * (2 * C * Hq * Wq + 2 * C * Dq * Hq * Wq) * B
            ref = None
            for v in (0, 1, 2, 3, 4, 5, 0, 3):  # 3 / 4 / 5 = 256-bit loads + stores (128 / 256 / 64 threads); 0 and 3 repeated
                with env(AZ_CONCAT_FWD=v):
                    ms = time_ms(lambda: ops.build_concat_volume(L, R, Dq))
                    out = ops.build_concat_volume(L, R, Dq)
                ref = out if ref is None else ref
                add("concat_volume_fwd", f"AZ_CONCAT_FWD={v}", ms, nb, {"bit_exact_vs_v0": bool(torch.equal(out, ref))})
                del out
            del ref
        if "gwc" in only:
            L, R = torch.randn(B, C, Hq, Wq, device=DEV), torch.randn(B, C, Hq, Wq, device=DEV)
            nbf = 4 * (2 * C * Hq * Wq + G * Dq * Hq * Wq) * B
            nbb = 4 * (4 * C * Hq * Wq + G * Dq * Hq * Wq) * B
            g = torch.randn(B, G, Dq, Hq, Wq, device=DEV)
            for v in (0, 8, 16, 48):
                with env(AZ_GWC_DG=v):
                    add("gwc_volume_fwd", f"AZ_GWC_DG={v}", time_ms(lambda: ops.build_gwc_volume(L, R, Dq, G), flush=True), nbf)
            from activezero_b200 import _lib
            from activezero_b200.ops import _ptr, _stream
            gL, gR = torch.empty_like(L), torch.empty_like(R)
            for v, nt, stg in ((0, 0, 0), (1, 0, 0), (2, 128, 4), (2, 128, 5), (2, 256, 4), (2, 256, 5)):
                with env(AZ_GWC_BWD=v, AZ_GWC_BWD_NT=nt, AZ_GWC_BWD_STAGES=stg):
                    add("gwc_volume_bwd", f"AZ_GWC_BWD={v}" + (f",NT={nt},STAGES={stg}" if nt else ""),
                        time_ms(lambda: _lib.call("az_gwc_volume_bwd", _ptr(g), _ptr(L), _ptr(R), _ptr(gL), _ptr(gR), B, C, Hq, Wq,
                                                  Dq, G, _stream()), flush=True), nbb)
        if "vconv" in only:
            import torch.nn.functional as F

            Bc = 2
            L, R = torch.randn(Bc, C, Hq, Wq, device=DEV), torch.randn(Bc, C, Hq, Wq, device=DEV)
            w = torch.randn(32, 64, 3, 3, 3, device=DEV) * 0.05
            wp = ops.pack_volume_conv_weight(w)
            flops = 2.0 * Bc * Dq * Hq * Wq * 32 * 64 * 27
            nb = 4 * (2 * C * Hq * Wq + 32 * Dq * Hq * Wq) * Bc  # features in, conv output out (no volume)

            def stock():
                return F.conv3d(ops.build_concat_volume(L, R, Dq), w, padding=1)

            wcl = w.contiguous(memory_format=torch.channels_last_3d)

            def stock_cl():
                return F.conv3d(ops.build_concat_volume(L, R, Dq, channels_last=True), wcl, padding=1)

            for name, fn in (("implicit volume, tcgen05 TF32 (this repo)", lambda: ops.volume_conv0(L, R, wp, Dq)),
                             ("az volume + cuDNN conv3d NCDHW (TF32)", stock), ("az volume (channels_last_3d) + cuDNN conv3d", stock_cl)):
                ms = time_ms(fn, flush=True)
                add("dres0_first_conv", name, ms, nb, {"TFLOPs": round(flops / ms / 1e9, 1), "pairs": Bc})
        if "patch" in only:
            gen = torch.Generator().manual_seed(5)
            pL = (torch.rand(B, 1, H, W, generator=gen) > 0.5).float().to(DEV)
            pR = (torch.rand(B, 1, H, W, generator=gen) > 0.5).float().to(DEV)
            mask = (torch.rand(B, 1, H, W, generator=gen) > 0.2).to(DEV)
            fields = {
                "random U(0,64)": (torch.rand(B, 1, H, W, generator=gen) * 64).to(DEV),
                "smooth sinusoid": (32 + 20 * torch.sin(torch.arange(W).float() / 40).view(1, 1, 1, W)
                                    + 8 * torch.cos(torch.arange(H).float() / 25).view(1, 1, H, 1)).expand(B, 1, H, W).contiguous().to(DEV),
            }
            low = torch.randn(B, 1, Dq, Hq, Wq, generator=gen).to(DEV) * 4
            fields["bench field (soft-argmin of upsampled random logits)"] = ops.upsample_soft_argmin(low, (D, H, W))
            nb = (4 * 3 * H * W + H * W + 4 * H * W) * B
            for name, d in fields.items():
                for v in (0, 1):
                    with env(AZ_PATCH_IMPL=v):
                        ms = time_ms(lambda: ops.reproj_loss(pL, pR, d, mask, ps=PS, want_warped=True), flush=True, graph=True)
                        loss, vis = ops.reproj_loss(pL, pR, d, mask, ps=PS, want_warped=True)
                    add("patch_loss+fold_fwd", f"AZ_PATCH_IMPL={v} / {name}", ms, nb, {"loss": float(loss), "vis_sum": float(vis.double().sum())})
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"device": torch.cuda.get_device_name(0), "peak_GBps": pk, "config": f"B={B} {H}x{W} D={D}", "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()

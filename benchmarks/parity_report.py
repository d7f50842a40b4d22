#!/usr/bin/env python
"""Worst observed errors per SURVEY §8 row, against the oracle evaluated in float64 (and bit-exactness where the gate is
bit-exact), at one moderate shape per row.  The pass / fail gates live in tests/; this prints the MARGINS
("report the worst relative error per row").

    python benchmarks/parity_report.py > profiles/r2_parity_report.json

The oracle is test infrastructure: this script is a checker, not a product path."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from activezero_b200 import ops  # noqa: E402
from activezero_b200.utils import reprojection as az_rp  # noqa: E402
from oracle import stereo_oracle as so  # noqa: E402

DEV = "cuda:0"
rows = []


def err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    e = (a - b).abs()
    scale = float(b.abs().max()) or 1.0
    big = b.abs() > 1e-3 * scale  # relative error is only meaningful away from zero crossings
    return {"max_abs": float(e.max()), "max_abs_over_scale": float(e.max()) / scale,
            "max_rel_where_|ref|>1e-3*scale": float((e[big] / b.abs()[big]).max()) if bool(big.any()) else 0.0}


def add(row, what, **kw):
    rows.append({"row": row, "what": what, **kw})


def main():
    torch.manual_seed(0)
    B, C, Hq, Wq, Dq = 1, 32, 34, 60, 24
    L64 = torch.randn(B, C, Hq, Wq, dtype=torch.float64, requires_grad=True)
    R64 = torch.randn(B, C, Hq, Wq, dtype=torch.float64, requires_grad=True)
    Lg = L64.detach().float().to(DEV).requires_grad_(True)
    Rg = R64.detach().float().to(DEV).requires_grad_(True)
    # a1 / a2
    vol64 = so.concat_volume(L64, R64, Dq)
    g = torch.randn(vol64.shape, dtype=torch.float64)
    vol64.backward(g)
    vol = ops.build_concat_volume(Lg, Rg, Dq)
    vol.backward(g.float().to(DEV))
    add("a1", "concat volume fwd", bit_exact=bool(torch.equal(vol.detach().cpu(), so.concat_volume(L64.detach().float(), R64.detach().float(), Dq))))
    add("a2", "concat volume bwd (gL, gR) vs float64 autograd", gL=err(Lg.grad, L64.grad), gR=err(Rg.grad, R64.grad))
    # a3
    L64.grad = R64.grad = None
    Lg.grad = Rg.grad = None
    gw64 = so.gwc_volume(L64, R64, Dq, 8)
    g = torch.randn(gw64.shape, dtype=torch.float64)
    gw64.backward(g)
    gw = ops.build_gwc_volume(Lg, Rg, Dq, 8)
    gw.backward(g.float().to(DEV))
    add("a3", "gwc volume fwd / bwd vs float64 restatement (parity unpinned: no reference operator)", fwd=err(gw, gw64), gL=err(Lg.grad, L64.grad),
        gR=err(Rg.grad, R64.grad))
    # a4 / a5
    D, H, W = 96, 68, 120
    for scale in (1.0, 10.0):
        c64 = (torch.randn(1, D, H, W, dtype=torch.float64) * scale).requires_grad_(True)
        cg = c64.detach().float().to(DEV).requires_grad_(True)
        d64 = so.soft_argmin(c64)
        gd = torch.randn(d64.shape, dtype=torch.float64)
        d64.backward(gd)
        dg = ops.soft_argmin(cg)
        dg.backward(gd.float().to(DEV))
        add("a4", f"soft-argmin fwd, logits N(0,1) x {scale:g}: error in pixels vs float64", **err(dg, d64))
        add("a5", f"soft-argmin bwd, logits x {scale:g} vs float64 autograd", **err(cg.grad, c64.grad))
    # f1
    low64 = (torch.randn(1, 1, 24, 17, 30, dtype=torch.float64) * 4).requires_grad_(True)
    lowg = low64.detach().float().to(DEV).requires_grad_(True)
    up64 = F.interpolate(low64, (96, 68, 120), mode="trilinear", align_corners=False).squeeze(1)
    d64 = so.soft_argmin(up64)
    gd = torch.randn(d64.shape, dtype=torch.float64)
    d64.backward(gd)
    dfu = ops.upsample_soft_argmin(lowg, (96, 68, 120))
    dfu.backward(gd.float().to(DEV))
    add("f1", "fused trilinear upsample + soft-argmin fwd (pixels) / bwd vs float64", fwd=err(dfu, d64), bwd=err(lowg.grad, low64.grad))
    # a6
    Hh, Ww = 64, 128
    img = torch.rand(2, 3, Hh, Ww)
    disp = (torch.rand(2, 1, Hh, Ww) * 1.3 - 0.15) * 40
    w = ops.warp(img.to(DEV), disp.to(DEV))
    add("a6", "bilinear disparity warp fwd vs apply_disparity (CPU torch, fp32)", bit_exact=bool(torch.equal(w.cpu(), so.apply_disparity(img, disp))))
    i64 = img.double().requires_grad_(True)
    d64 = disp.double().requires_grad_(True)
    go = torch.randn(2, 3, Hh, Ww, dtype=torch.float64)
    so.apply_disparity(i64, d64).backward(go)
    ig, dgp = img.to(DEV).requires_grad_(True), disp.to(DEV).requires_grad_(True)
    ops.warp(ig, dgp).backward(go.float().to(DEV))
    add("a6", "warp bwd (image, disparity gradients) vs float64 autograd", gimg=err(ig.grad, i64.grad), gdisp=err(dgp.grad, d64.grad))
    # a7 / a8
    pL, pR = (torch.rand(2, 1, Hh, Ww) > 0.5).float(), (torch.rand(2, 1, Hh, Ww) > 0.5).float()
    mask = torch.rand(2, 1, Hh, Ww) > 0.2
    dsp = torch.rand(2, 1, Hh, Ww) * 30
    for ps in (11, 1):
        d64 = dsp.double().requires_grad_(True)
        if ps == 1:
            l64, wv64, _ = so.reproj_error_old(pL.double(), pR.double(), d64, mask)
        else:
            l64, wv64, _ = so.reproj_error_patch(pL.double(), pR.double(), d64, mask, ps=ps)
        l64.backward()
        dg = dsp.to(DEV).requires_grad_(True)
        if ps == 1:
            lg, wv, _ = az_rp.get_reprojection_error_old(pL.to(DEV), pR.to(DEV), dg, mask.to(DEV))
        else:
            lg, wv, _ = az_rp.get_reproj_error_patch(pL.to(DEV), pR.to(DEV), dg, mask.to(DEV), ps=ps)
        lg.backward()
        add("a7" if ps > 1 else "a8", f"reprojection loss ps={ps}: loss, image output, disparity gradient vs float64",
            loss_rel=abs(float(lg.detach()) - float(l64.detach())) / abs(float(l64.detach())), image=err(wv, wv64), gdisp=err(dg.grad, d64.grad))
    # a10
    di = ((torch.rand(2, 1, Hh, Ww) * 60).int()) * (torch.rand(2, 1, Hh, Ww) > 0.3).int()
    src = torch.rand(2, 1, Hh, Ww)
    add("a10", "integer scatter warp", bit_exact=bool(torch.equal(ops.scatter_warp(src.to(DEV), di.to(DEV)).cpu(), so.scatter_warp(src, di))))
    # a11
    fr = np.random.RandomState(1).randint(0, 256, (7, 180, 320)).astype(np.uint8)
    ref = so.temporal_ir_pattern(fr)
    pat = ops.temporal_ir_pattern(torch.from_numpy(fr).to(DEV)).cpu().numpy()
    add("a11", "temporal IR pattern: mismatching pixels", mismatches=int((pat != ref).sum()), of=int(ref.size))
    # a12
    im = torch.rand(2, 1, Hh, Ww)
    n64, s64 = so.local_contrast_norm(im.double(), 9)
    n, sd = ops.local_contrast_norm(im.to(DEV), 9)
    add("a12", "local contrast normalisation vs float64", normed=err(n, n64), std=err(sd, s64))
    # f2
    Lf, Rf = torch.randn(1, 32, 9, 140, device=DEV), torch.randn(1, 32, 9, 140, device=DEV)
    wc = torch.randn(32, 64, 3, 3, 3, device=DEV) * 0.05
    out = ops.volume_conv0(Lf, Rf, ops.pack_volume_conv_weight(wc), 12)
    vol64 = ops.build_concat_volume(Lf, Rf, 12).double().cpu()
    ref64 = F.conv3d(vol64, wc.double().cpu(), padding=1)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True
    cud = F.conv3d(ops.build_concat_volume(Lf, Rf, 12), wc, padding=1)
    torch.backends.cudnn.allow_tf32 = old
    add("f2", "first aggregation conv on the implicit volume (TF32 operands) vs float64; cuDNN TF32 beside it", this=err(out, ref64),
        cudnn_tf32=err(cud, ref64))
    json.dump(rows, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()

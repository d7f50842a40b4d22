#!/usr/bin/env python
"""Probe for SURVEY.md §8f rank 2 ("emit the volume in the layout cuDNN wants"): time the 3-D aggregation
(dres0-4 + classif heads, plain torch/cuDNN, NOT part of this library) on the concat volume in NCDHW versus
channels_last_3d, fp32 (TF32 allowed, torch's default for cuDNN) and bf16 autocast.  Informational only."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from activezero_b200 import ops  # noqa: E402
from activezero_b200.nets.psmnet.psmnet_3 import PSMNet  # noqa: E402


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    torch.manual_seed(1)
    net = PSMNet(maxdisp=192).cuda().eval()
    fl = torch.randn(B, 32, 136, 240, device="cuda")
    fr = torch.randn(B, 32, 136, 240, device="cuda")
    res = {"batch": B}
    with torch.no_grad():
        vol = ops.build_concat_volume(fl, fr, 48)
        res["ncdhw_fp32_ms"] = timed(lambda: net._aggregate(vol))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            res["ncdhw_bf16_ms"] = timed(lambda: net._aggregate(vol))
        net_cl = net
        for m in net_cl.modules():  # only the 3-D convolutions (the 2-D feature CNN has rank-4 weights)
            if isinstance(m, (torch.nn.Conv3d, torch.nn.ConvTranspose3d)):
                m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last_3d)
        vol_cl = vol.contiguous(memory_format=torch.channels_last_3d)
        res["ndhwc_fp32_ms"] = timed(lambda: net_cl._aggregate(vol_cl))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            res["ndhwc_bf16_ms"] = timed(lambda: net_cl._aggregate(vol_cl))
        res["layout_convert_ms"] = timed(lambda: vol.contiguous(memory_format=torch.channels_last_3d))
    print(json.dumps(res))


if __name__ == "__main__":
    main()

"""The drop-in PSMNet modules: checkpoint compatibility (CPU), plain-torch parts
against the reference modules (CPU, authoring container only), and the whole
network against outputs captured from the real reference (GPU)."""
import importlib
import json
import os
import sys

import numpy as np
import pytest
import torch

from activezero_b200.nets.psmnet import psmnet as az_psm6
from activezero_b200.nets.psmnet import psmnet_3 as az_psm3
from activezero_b200.nets.psmnet.psmnet_submodule import DisparityRegression
from oracle import ref_loader

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name,mod", [("psmnet", az_psm6), ("psmnet_3", az_psm3)])
def test_state_dict_keys_match_reference(name, mod):
    want = json.load(open(os.path.join(GOLDEN, "psmnet_state_dict_keys.json")))[name]
    got = [[k, list(v.shape)] for k, v in mod.PSMNet(maxdisp=192).state_dict().items()]
    assert got == want


def test_forward_signatures():
    import inspect

    assert list(inspect.signature(az_psm6.PSMNet.forward).parameters) == [
        "self", "img_L", "img_R", "img_L_transformed", "img_R_transformed"]
    assert list(inspect.signature(az_psm3.PSMNet.forward).parameters) == ["self", "img_L", "img_R"]
    assert az_psm3.PSMNet().maxdisp == 192


def test_disparity_regression_api():
    p = torch.softmax(torch.randn(2, 16, 3, 4), 1)
    out = DisparityRegression(16)(p)
    ref = (p * torch.arange(16.0).view(1, 16, 1, 1)).sum(1, keepdim=True)
    assert torch.equal(out, ref)


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference (authoring container)")
def test_plain_torch_parts_match_reference_modules():
    """Same weights -> identical feature maps and aggregated logits on the CPU."""
    ref_loader.load()
    sys.path.insert(0, ref_loader.REFERENCE_ROOT)
    try:
        ref3 = importlib.import_module("nets.psmnet.psmnet_3")
    finally:
        sys.path.remove(ref_loader.REFERENCE_ROOT)
    torch.manual_seed(1)
    ref = ref3.PSMNet(maxdisp=48).eval()
    mine = az_psm3.PSMNet(maxdisp=48).eval()
    mine.load_state_dict(ref.state_dict())
    x = torch.rand(1, 3, 256, 256)
    with torch.no_grad():
        fr, fm = ref.feature_extraction(x), mine.feature_extraction(x)
        assert torch.equal(fr, fm)
        vol = torch.randn(1, 64, 12, 16, 16)
        c0 = ref.dres0(vol)
        c0 = ref.dres1(c0) + c0
        o1, p1, q1 = ref.dres2(c0, None, None)
        o1 = o1 + c0
        o2, p2, q2 = ref.dres3(o1, p1, q1)
        o2 = o2 + c0
        o3, _, _ = ref.dres4(o2, p1, q2)
        o3 = o3 + c0
        r1 = ref.classif1(o1)
        r2 = ref.classif2(o2) + r1
        r3 = ref.classif3(o3) + r2
        m1, m2, m3 = mine._aggregate(vol)
        assert torch.equal(m1, r1) and torch.equal(m2, r2) and torch.equal(m3, r3)


@pytest.mark.gpu
def test_psmnet_end_to_end_against_reference_outputs():
    """tests/golden/psmnet_inline.npz was captured from the reference PSMNet_3 built
    under torch.manual_seed(1) and fed torch.rand inputs from the same stream; the
    drop-in consumes the RNG identically, so it is the same network on the same
    images.  cuDNN vs the CPU convolutions differ at ~1e-4 relative."""
    g = np.load(os.path.join(GOLDEN, "psmnet_inline.npz"))
    torch.manual_seed(1)
    net = az_psm3.PSMNet(maxdisp=192).eval()
    img_L, img_R = torch.rand(1, 3, 256, 256), torch.rand(1, 3, 256, 256)
    net = net.cuda()
    feats = []
    net.feature_extraction.register_forward_hook(lambda m, i, o: feats.append(o.detach()))
    vols = []
    net.dres0.register_forward_pre_hook(lambda m, i: vols.append(i[0].detach()))
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            pred = net(img_L.cuda(), img_R.cuda())
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    assert pred.shape == (1, 1, 256, 256)
    ch, rows = [3, 17], slice(20, 24)
    fl = feats[0][:, ch, rows].cpu().numpy()
    np.testing.assert_allclose(fl, g["feat_L"], rtol=2e-3, atol=2e-3 * np.abs(g["feat_L"]).max())
    # the volume is an exact rearrangement of the features that entered it
    from oracle import stereo_oracle as so

    assert torch.equal(vols[0], so.concat_volume(feats[0], feats[1], 48))
    np.testing.assert_allclose(pred[:, :, 100:104, 64:96].cpu().numpy(), g["pred"], atol=2e-2)


@pytest.mark.gpu
def test_psmnet_train_mode_backward():
    torch.manual_seed(0)
    net = az_psm6.PSMNet(maxdisp=48).cuda().train()
    a = [torch.rand(2, 3, 256, 256, device="cuda") for _ in range(4)]  # BN in the 1x1 SPP branch needs > 1 sample
    p3, p2, p1 = net(*a)
    assert p3.shape == p2.shape == p1.shape == (2, 1, 256, 256)
    (p3.mean() + 0.7 * p2.mean() + 0.5 * p1.mean()).backward()
    g = net.feature_extraction.firstconv[0][0].weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0
    g3 = net.classif3[2].weight.grad
    assert torch.isfinite(g3).all() and float(g3.abs().max()) > 0


@pytest.mark.gpu
def test_psmnet_channels_last_3d_volume_same_predictions_and_gradients():
    """use_channels_last_3d() changes strides only: same state_dict, same prediction (TF32 off so that the
    different cuDNN algorithms agree to rounding), and training gradients flow through the in-place
    channels_last_3d backward of the volume."""
    torch.manual_seed(3)
    net = az_psm3.PSMNet(maxdisp=96).cuda().eval()
    keys = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    x, y = torch.rand(1, 3, 256, 256, device="cuda"), torch.rand(1, 3, 256, 256, device="cuda")  # SPP needs >= 256
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            p0 = net(x, y)
            net.use_channels_last_3d(True)
            vols = []
            h = net.dres0.register_forward_pre_hook(lambda m, i: vols.append(i[0]))
            p1 = net(x, y)
            h.remove()
        assert vols[0].is_contiguous(memory_format=torch.channels_last_3d) and not vols[0].is_contiguous()
        assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == keys
        assert float((p0 - p1).abs().max()) <= 2e-3, float((p0 - p1).abs().max())
        # gradients: channels_last_3d vs NCDHW volume, same network
        grads = []
        for cl in (True, False):
            net.use_channels_last_3d(cl).train()
            net.zero_grad(set_to_none=True)
            torch.manual_seed(5)
            p3, p2, pa = net(torch.cat([x, x * 0.5]), torch.cat([y, y * 0.5]))
            (p3.mean() + 0.7 * p2.mean() + 0.5 * pa.mean()).backward()
            grads.append(net.feature_extraction.firstconv[0][0].weight.grad.clone())
        scale = float(grads[1].abs().max())
        assert scale > 0 and float((grads[0] - grads[1]).abs().max()) <= 2e-2 * scale
    finally:
        torch.backends.cudnn.allow_tf32 = prev

"""Generate ``tests/golden/*.npz`` by EXECUTING THE REAL REFERENCE CODE
(authoring container only; needs ``/root/reference``).

    python -m oracle.make_golden

How each fixture is produced (all inputs seeded, all on CPU):

* ``psmnet_inline.npz`` -- the reference ``PSMNet`` (``nets/psmnet/psmnet_3.py``)
  is instantiated and its ``forward`` executed unmodified; the only
  intervention is that ``torch.Tensor.cuda`` is a no-op for the duration, so the
  inline cost-volume code (``psmnet_3.py:149-163`` = ``psmnet.py:151-165``) and
  ``DisparityRegression`` (``psmnet_submodule_3.py:80-89``) run on the CPU.
  Forward hooks capture the two feature maps, the volume entering ``dres0``,
  the logits entering the last ``F.softmax`` and the returned disparity.  All
  ops on this path are independent per (channel,row) / per pixel, so a slice of
  the captured tensors is stored, not the 50 MB volume.
* ``soft_argmin_synth.npz`` -- ``F.softmax`` + the reference
  ``DisparityRegression`` class on peaky synthetic logits (scale 1 and 10),
  with the autograd gradient for a seeded upstream gradient.
* ``reprojection.npz`` -- ``utils/reprojection.py`` functions called directly
  (module imported with cupy/pynvrtc stubbed): ``apply_disparity``,
  ``get_reprojection_error_old``, ``get_reprojection_error`` (masks given),
  ``get_reproj_error_patch`` (ps 5 and 11, with/without mask),
  ``get_reprojection_error_diff_ratio``, ``local_contrast_norm``; autograd
  gradients w.r.t. the disparity where the trainer needs them.
* ``scatter_warp.npz`` -- the reference's CUDA-C kernel string compiled for
  the CPU by ``oracle/build_ref.py`` (unmodified kernel bodies).
* ``temporal_ir.npz`` -- ``tools/temporal_ir.py:main()`` executed on a temp
  directory of synthetic PNG frames, with ``matplotlib`` stubbed so that
  ``plt.imsave`` hands the pattern array back instead of writing a PNG.
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile
import types

import numpy as np
import torch

from . import build_ref, ref_loader

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def gen_psmnet_inline():
    ref_loader.load()
    sys.path.insert(0, ref_loader.REFERENCE_ROOT)
    try:
        psm = importlib.import_module("nets.psmnet.psmnet_3")
    finally:
        sys.path.remove(ref_loader.REFERENCE_ROOT)
    torch.manual_seed(1)
    saved_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    cap = {}
    real_softmax = psm.F.softmax

    def spy_softmax(x, dim=None, **kw):
        cap["logits"] = x.detach().clone()
        return real_softmax(x, dim=dim, **kw)

    try:
        net = psm.PSMNet(maxdisp=192)
        net.eval()
        feats = []
        net.feature_extraction.register_forward_hook(lambda m, i, o: feats.append(o.detach().clone()))
        net.dres0.register_forward_pre_hook(lambda m, i: cap.__setitem__("vol", i[0].detach().clone()))
        psm.F.softmax = spy_softmax
        with torch.no_grad():
            pred = net(torch.rand(1, 3, 256, 256), torch.rand(1, 3, 256, 256))
    finally:
        psm.F.softmax = real_softmax
        torch.Tensor.cuda = saved_cuda
    L, R = feats
    vol = cap["vol"]
    C = L.shape[1]
    assert vol.shape == (1, 2 * C, 48, 64, 64)
    ch, rows = [3, 17], slice(20, 24)
    np.savez_compressed(
        os.path.join(GOLDEN, "psmnet_inline.npz"),
        feat_L=_np(L[:, ch, rows]), feat_R=_np(R[:, ch, rows]),
        vol=_np(vol[:, ch + [C + c for c in ch], :, rows]),
        logits=_np(cap["logits"][:, :, 100:104, 64:96]),
        pred=_np(pred[:, :, 100:104, 64:96]),
        num_disp=48,
    )


def gen_soft_argmin():
    _, sub = ref_loader.load()
    saved_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    out = {}
    try:
        for tag, scale, D in (("s1", 1.0, 192), ("s10", 10.0, 192), ("d96", 4.0, 96)):
            torch.manual_seed(7)
            cost = (torch.randn(1, D, 4, 20) * scale).requires_grad_(True)
            g = torch.randn(1, 1, 4, 20)
            pred = sub.DisparityRegression(D)(torch.nn.functional.softmax(cost, dim=1))
            pred.backward(g)
            out.update({f"{tag}_cost": _np(cost), f"{tag}_g": _np(g), f"{tag}_pred": _np(pred),
                        f"{tag}_gcost": _np(cost.grad)})
    finally:
        torch.Tensor.cuda = saved_cuda
    np.savez_compressed(os.path.join(GOLDEN, "soft_argmin_synth.npz"), **out)


def gen_reprojection():
    rp, _ = ref_loader.load()
    out = {}
    torch.manual_seed(3)
    B, H, W = 2, 24, 40
    L1 = torch.rand(B, 1, H, W)
    R1 = torch.rand(B, 1, H, W)
    L3 = torch.rand(B, 3, H, W)
    R3 = torch.rand(B, 3, H, W)
    disp = torch.rand(B, 1, H, W) * 12.0
    disp[0, 0, :4] += 40.0  # push samples out of bounds on the left
    disp[1, 0, -3:] = -6.5  # and (negative disparity) on the right
    disp_r = torch.rand(B, 1, H, W) * 9.0
    mask = torch.rand(B, 1, H, W) > 0.4
    mask_r = torch.rand(B, 1, H, W) > 0.5
    out.update(L1=_np(L1), R1=_np(R1), L3=_np(L3), R3=_np(R3), disp=_np(disp), disp_r=_np(disp_r),
               mask=_np(mask), mask_r=_np(mask_r))

    # a6 apply_disparity, value + grads
    d = disp.clone().requires_grad_(True)
    img = R3.clone().requires_grad_(True)
    w = rp.apply_disparity(img, -d)
    gw = torch.rand_like(w)
    w.backward(gw)
    out.update(warp3=_np(w), warp3_gout=_np(gw), warp3_gdisp=_np(d.grad), warp3_gimg=_np(img.grad))
    out["warp1_pos"] = _np(rp.apply_disparity(L1, disp_r))

    # a8 single-scale
    for tag, Li, Ri, m in (("old1m", L1, R1, mask), ("old3m", L3, R3, mask), ("old1", L1, R1, None)):
        d = disp.clone().requires_grad_(True)
        loss, warped, mi = rp.get_reprojection_error_old(Li, Ri, d, m)
        loss.backward()
        out.update({f"{tag}_loss": _np(loss), f"{tag}_warped": _np(warped), f"{tag}_mask": _np(mi),
                    f"{tag}_gdisp": _np(d.grad)})
    dl = disp.clone().requires_grad_(True)
    dr = disp_r.clone().requires_grad_(True)
    ll, lr, wl, wr, ml, mr = rp.get_reprojection_error(L3, R3, dl, dr, mask, mask_r)
    (ll + 2 * lr).backward()
    out.update(bi_loss_l=_np(ll), bi_loss_r=_np(lr), bi_warp_l=_np(wl), bi_warp_r=_np(wr),
               bi_mask_l=_np(ml), bi_mask_r=_np(mr), bi_gdisp_l=_np(dl.grad), bi_gdisp_r=_np(dr.grad))

    # a7 patch loss
    for tag, Li, Ri, m, ps in (("p5m", L1, R1, mask, 5), ("p11m", L1, R1, mask, 11), ("p11", L1, R1, None, 11),
                               ("p3c3m", L3, R3, mask, 3), ("p1m", L1, R1, mask, 1)):
        d = disp.clone().requires_grad_(True)
        loss, vis, mi = rp.get_reproj_error_patch(Li, Ri, d, m, ps=ps)
        loss.backward()
        out.update({f"{tag}_loss": _np(loss), f"{tag}_vis": _np(vis), f"{tag}_mask": _np(mi),
                    f"{tag}_gdisp": _np(d.grad)})
    # empty mask -> NaN
    loss, _, _ = rp.get_reproj_error_patch(L1, R1, disp, torch.zeros_like(mask), ps=5)
    out["p5empty_loss"] = _np(loss)

    # a9 multi-scale
    d = disp.clone().requires_grad_(True)
    tot, stages, ld = rp.get_reprojection_error_diff_ratio(L3, R3, d, mask)
    tot.backward()
    out.update(ms_loss=_np(tot), ms_gdisp=_np(d.grad))
    for k in range(3):
        out[f"ms_stage{k}_loss"] = np.float32(ld[f"stage{k}"])
        out[f"ms_stage{k}_warped"] = _np(stages[f"stage{k}"]["warped"])
        out[f"ms_stage{k}_mask"] = _np(stages[f"stage{k}"]["mask"])

    # a12 LCN
    img = torch.rand(2, 2, 20, 30)
    n9, s9 = rp.local_contrast_norm(img, 9)
    n5, s5 = rp.local_contrast_norm(img, 5, eps=1e-3)
    out.update(lcn_img=_np(img), lcn9_norm=_np(n9), lcn9_std=_np(s9), lcn5_norm=_np(n5), lcn5_std=_np(s5))
    np.savez_compressed(os.path.join(GOLDEN, "reprojection.npz"), **out)


def gen_reprojection_wide():
    """get_reproj_error_patch (utils/reprojection.py:99-127) of the REAL reference on frames wide enough for the
    one-pass loss + Fold kernel to use several strips per row, several bands per image and all its warp groups:
    Bernoulli IR patterns (they compress), sloped disparities with an out-of-range region and boundary-cell samples."""
    rp, _ = ref_loader.load()
    out = {}
    torch.manual_seed(21)
    for tag, (B, C, H, W), ps, masked in (("w11", (1, 1, 44, 300), 11, True), ("w7", (1, 2, 30, 260), 7, False)):
        L = (torch.rand(B, C, H, W) > 0.5).float()
        R = (torch.rand(B, C, H, W) > 0.5).float()
        xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W).expand(B, 1, H, W)
        ys = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1).expand(B, 1, H, W)
        disp = (9.0 + 0.07 * xs + 0.11 * ys + torch.rand(B, 1, H, W)).contiguous()
        disp[:, :, :5] = xs[:, :, :5] + 0.25          # samples in the boundary cell x0 = -1 / 0
        disp[:, :, -4:, : W // 2] += 2.0 * W          # far out of range on the left
        mask = (torch.rand(B, 1, H, W) > 0.3) if masked else None
        d = disp.clone().requires_grad_(True)
        loss, vis, mi = rp.get_reproj_error_patch(L, R, d, mask, ps=ps)
        loss.backward()
        out.update({f"{tag}_L": _np(L).astype(np.uint8), f"{tag}_R": _np(R).astype(np.uint8), f"{tag}_disp": _np(disp),
                    f"{tag}_loss": _np(loss), f"{tag}_vis": _np(vis), f"{tag}_gdisp": _np(d.grad)})
        if masked:
            out[f"{tag}_maskin"] = _np(mask)
    np.savez_compressed(os.path.join(GOLDEN, "reprojection_wide.npz"), **out)


def reprojection_720_inputs():
    """Deterministic inputs of the 720x1280 fixture (the real IR frame size, datasets/messytable.py:325), rebuilt
    identically by the generator and by the tests (numpy RandomState: stable across versions and machines)."""
    rs = np.random.RandomState(720)
    H, W = 720, 1280
    L = (rs.rand(1, 1, H, W) > 0.5).astype(np.float32)
    R = (rs.rand(1, 1, H, W) > 0.5).astype(np.float32)
    xs = np.arange(W, dtype=np.float32)[None, None, None, :]
    ys = np.arange(H, dtype=np.float32)[None, None, :, None]
    disp = (40.0 + 28.0 * np.sin(xs / 37.0) * np.cos(ys / 53.0) + 6.0 * rs.rand(1, 1, H, W)).astype(np.float32)
    disp[:, :, 100:140, :200] = (xs[:, :, :, :200] + 0.3)              # boundary cell x0 = -1 / 0 on the left edge
    disp[:, :, 300:330, 900:] -= 150.0                                  # negative: samples leave on the right
    mask = rs.rand(1, 1, H, W) > 0.25
    return torch.from_numpy(L), torch.from_numpy(R), torch.from_numpy(np.ascontiguousarray(disp)), torch.from_numpy(mask)


def gen_reprojection_720():
    """get_reproj_error_patch (utils/reprojection.py:99-127) of the REAL reference on one 720x1280 frame, ps = 11.
    The inputs are rebuilt from a seed by the test; the fixture stores the loss, strided samples of the Fold image
    and of d loss / d disp (strides 7 x 5: every residue of the kernels' tilings is hit) and their full sums."""
    rp, _ = ref_loader.load()
    L, R, disp, mask = reprojection_720_inputs()
    d = disp.clone().requires_grad_(True)
    loss, vis, mi = rp.get_reproj_error_patch(L, R, d, mask, ps=11)
    loss.backward()
    g = d.grad
    out = {"loss": _np(loss), "vis_s": _np(vis)[:, :, ::7, ::5], "gdisp_s": _np(g)[:, :, ::7, ::5],
           "vis_sum": np.float64(_np(vis).astype(np.float64).sum()), "vis_abs_max": np.float64(_np(vis).max()),
           "gdisp_sum": np.float64(_np(g).astype(np.float64).sum()), "gdisp_abs_sum": np.float64(np.abs(_np(g).astype(np.float64)).sum()),
           "gdisp_abs_max": np.float64(np.abs(_np(g)).max()), "mask_sum": np.int64(_np(mi).sum())}
    np.savez_compressed(os.path.join(GOLDEN, "reprojection_720.npz"), **out)


def gen_scatter_warp():
    assert build_ref.build(force=True), "reference kernel source not found"
    rng = np.random.default_rng(11)
    out = {}
    N, C, H, W = 2, 2, 16, 48
    img = torch.from_numpy(rng.random((N, C, H, W), dtype=np.float32) * 100 + 1)
    d = rng.integers(0, 40, size=(N, 1, H, W)).astype(np.int32)
    d[rng.random(d.shape) < 0.3] = 0
    dpos = torch.from_numpy(d)
    dneg = torch.from_numpy(-d)
    out.update(img=_np(img), disp_pos=d, disp_neg=-d,
               out_pos=_np(build_ref.ref_scatter_warp(img, dpos)),
               out_neg=_np(build_ref.ref_scatter_warp(img, dneg)))
    # disparity warped by itself (the trainer's use, train.py:266-268) and W not a multiple of 4
    dm = torch.from_numpy(rng.random((1, 1, 9, 37), dtype=np.float32) * 30)
    di = dm.type(torch.int)
    out.update(self_img=_np(dm), self_disp=_np(di), self_out=_np(build_ref.ref_scatter_warp(dm, di)))
    np.savez_compressed(os.path.join(GOLDEN, "scatter_warp.npz"), **out)


def gen_temporal_ir():
    from PIL import Image

    rng = np.random.default_rng(5)
    T, H, W = 7, 48, 64
    base = rng.integers(0, 256, size=(H, W)).astype(np.float64) * 0.5
    dots = (rng.random((H, W)) < 0.1).astype(np.float64)
    seqs = {}
    for direction in ("irL", "irR"):
        noise = rng.integers(0, 6, size=(T, H, W))
        fr = np.stack([base + t * dots * 10.0 for t in range(T)]) + noise
        if direction == "irR":
            fr = np.roll(fr, 5, axis=2)
        seqs[direction] = np.clip(fr, 0, 255).astype(np.uint8)
    names = ["off", "060", "120", "180", "240", "300", "360"]  # temporal_ir.py:64-70
    captured = {}
    with tempfile.TemporaryDirectory() as tmp:
        scene = "scene0"
        os.makedirs(os.path.join(tmp, scene))
        for direction, fr in seqs.items():
            for t, nm in enumerate(names):
                Image.fromarray(fr[t], mode="L").save(os.path.join(tmp, scene, f"1024_{direction}_real_{nm}.png"))
        split = os.path.join(tmp, "split.txt")
        open(split, "w").write(scene + "\n")
        plt = types.ModuleType("matplotlib.pyplot")
        plt.imsave = lambda path, arr, **kw: captured.__setitem__(os.path.basename(path), np.array(arr))
        plt.close = lambda *a, **k: None
        mpl = types.ModuleType("matplotlib")
        mpl.pyplot = plt
        saved = {k: sys.modules.get(k) for k in ("matplotlib", "matplotlib.pyplot")}
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
        argv = sys.argv
        sys.argv = ["temporal_ir.py", "--split-file", split, "--data-folder", tmp]
        sys.path.insert(0, ref_loader.REFERENCE_ROOT)
        try:
            tir = importlib.import_module("tools.temporal_ir")
            tir.main()
        finally:
            sys.path.remove(ref_loader.REFERENCE_ROOT)
            sys.argv = argv
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
    np.savez_compressed(
        os.path.join(GOLDEN, "temporal_ir.npz"),
        frames_L=seqs["irL"], frames_R=seqs["irR"],
        pattern_L=captured["1024_irL_real_temporal.png"], pattern_R=captured["1024_irR_real_temporal.png"],
    )


def gen_err_metrics():
    """utils/cascade_metrics.py:compute_err_metric called directly (train.py:348-351 shapes:
    focal_length / baseline are [bs,1,1,1])."""
    ref_loader.load()
    sys.path.insert(0, ref_loader.REFERENCE_ROOT)
    try:
        cm = importlib.import_module("utils.cascade_metrics")
    finally:
        sys.path.remove(ref_loader.REFERENCE_ROOT)
    torch.manual_seed(9)
    B, H, W = 2, 20, 36
    disp_gt = torch.rand(B, 1, H, W) * 80 + 1
    disp_pred = disp_gt + torch.randn(B, 1, H, W) * 1.5
    focal = torch.tensor([[[[430.0]]], [[[455.5]]]])
    base = torch.tensor([[[[0.055]]], [[[0.0545]]]])
    depth_gt = focal * base / disp_gt + torch.randn(B, 1, H, W) * 1e-3
    mask = torch.rand(B, 1, H, W) > 0.3
    m1 = cm.compute_err_metric(disp_gt, depth_gt, disp_pred, focal, base, mask)
    depth_pred = focal * base / disp_pred + 2e-3
    m2 = cm.compute_err_metric(disp_gt, depth_gt, disp_pred, focal, base, mask, depth_pred=depth_pred)
    out = dict(disp_gt=_np(disp_gt), disp_pred=_np(disp_pred), depth_gt=_np(depth_gt), focal=_np(focal), base=_np(base),
               mask=_np(mask), depth_pred=_np(depth_pred))
    for k, v in m1.items():
        out["m1_" + k] = np.float64(v)
    for k, v in m2.items():
        out["m2_" + k] = np.float64(v)
    np.savez_compressed(os.path.join(GOLDEN, "err_metrics.npz"), **out)


def gen_sim_ir_pattern():
    """datasets/dataset_utils.py:get_ir_pattern / get_smoothed_ir_pattern2 called directly (the module's
    `from configs.config import cfg` needs yacs, which is not installed: a stub module stands in; cfg is
    not used by the two functions)."""
    ref_loader.load()
    stub = types.ModuleType("configs.config")
    stub.cfg = types.SimpleNamespace()
    pkg = types.ModuleType("configs")
    pkg.config = stub
    saved = {k: sys.modules.get(k) for k in ("configs", "configs.config")}
    sys.modules["configs"], sys.modules["configs.config"] = pkg, stub
    sys.path.insert(0, ref_loader.REFERENCE_ROOT)
    try:
        du = importlib.import_module("datasets.dataset_utils")
    finally:
        sys.path.remove(ref_loader.REFERENCE_ROOT)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    rng = np.random.default_rng(21)
    out = {}
    for tag, (h, w) in (("a", (54, 96)), ("b", (77, 121)), ("c", (40, 61))):
        base = rng.integers(0, 200, size=(h, w))
        dots = (rng.random((h, w)) < 0.12) * rng.integers(20, 56, size=(h, w))
        img_u8 = base.astype(np.uint8)
        ir_u8 = np.clip(base + dots + rng.integers(0, 3, size=(h, w)), 0, 255).astype(np.uint8)
        img, ir = img_u8 / 255, ir_u8 / 255  # datasets/messytable.py loads PNGs as uint8 and divides by 255
        out.update({f"{tag}_img_u8": img_u8, f"{tag}_ir_u8": ir_u8,
                    f"{tag}_p1": du.get_ir_pattern(ir, img),
                    f"{tag}_p2": du.get_smoothed_ir_pattern2(ir, img),
                    f"{tag}_p2_k5": du.get_smoothed_ir_pattern2(ir, img, ks=5, threshold=0.01)})
    np.savez_compressed(os.path.join(GOLDEN, "sim_ir_pattern.npz"), **out)


def gen_state_dict_keys():
    """Names and shapes of every state_dict entry of both reference PSMNet variants
    (checkpoint compatibility contract, test.py:341-342 / train.py:155-170)."""
    import json

    ref_loader.load()
    sys.path.insert(0, ref_loader.REFERENCE_ROOT)
    try:
        mods = {"psmnet": importlib.import_module("nets.psmnet.psmnet"),
                "psmnet_3": importlib.import_module("nets.psmnet.psmnet_3")}
    finally:
        sys.path.remove(ref_loader.REFERENCE_ROOT)
    out = {}
    for name, mod in mods.items():
        sd = mod.PSMNet(maxdisp=192).state_dict()
        out[name] = [[k, list(v.shape)] for k, v in sd.items()]
    with open(os.path.join(GOLDEN, "psmnet_state_dict_keys.json"), "w") as f:
        json.dump(out, f)


def main():
    assert ref_loader.available(), "run in the authoring container (needs /root/reference)"
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    for fn in (gen_scatter_warp, gen_temporal_ir, gen_reprojection, gen_soft_argmin, gen_psmnet_inline,
               gen_state_dict_keys, gen_err_metrics, gen_sim_ir_pattern, gen_reprojection_wide, gen_reprojection_720):
        print("generating", fn.__name__, flush=True)
        fn()
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()

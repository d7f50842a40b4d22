"""GPU: seeded random small shapes for every operator against the oracle -- ragged sizes (W not a
multiple of 4/32, H or W smaller than the patch, Dq > W, C > 1, B > 1), the edge cases the fixed-shape
tests do not enumerate."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import stereo_oracle as so  # noqa: E402

if torch.cuda.is_available():
    from activezero_b200 import ops
    from activezero_b200.utils import reprojection as az_rp
    from activezero_b200.utils import warp_ops as az_wo

DEV = "cuda:0"


def close(a, b, rtol=1e-5, floor=1e-5):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    atol = floor * float(b.abs().max()) + 1e-30 if b.numel() else 0.0
    err = (a - b).abs()
    assert bool((err <= atol + rtol * b.abs()).all()), f"max err {float(err.max()):.3e}, atol {atol:.3e}, shape {tuple(b.shape)}"


def _cases(seed, n):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        yield rng


@pytest.mark.parametrize("case", range(24))
def test_volumes_random(case):
    rng = np.random.default_rng(1000 + case)
    B, H = int(rng.integers(1, 4)), int(rng.integers(1, 13))
    W = int(rng.choice([1, 3, 4, 7, 8, 12, 20, 33, 64, 100]))
    G = int(rng.choice([1, 2, 4]))
    C = G * int(rng.choice([1, 2, 4, 8, 3]))
    Dq = int(rng.choice([1, 2, 5, 8, 12, 17, 48]))
    torch.manual_seed(case)
    L = torch.randn(B, C, H, W, dtype=torch.float64, requires_grad=True)
    R = torch.randn(B, C, H, W, dtype=torch.float64, requires_grad=True)
    vol = so.concat_volume(L, R, Dq)
    g = torch.randn(vol.shape, dtype=torch.float64)
    vol.backward(g)
    Lg = L.detach().float().to(DEV).requires_grad_(True)
    Rg = R.detach().float().to(DEV).requires_grad_(True)
    out = ops.build_concat_volume(Lg, Rg, Dq)
    assert torch.equal(out.detach().cpu(), so.concat_volume(L.detach().float(), R.detach().float(), Dq))
    out.backward(g.float().to(DEV))
    close(Lg.grad, L.grad)
    close(Rg.grad, R.grad)
    # gwc
    L.grad = R.grad = None
    gv = so.gwc_volume(L, R, Dq, G)
    gg = torch.randn(gv.shape, dtype=torch.float64)
    gv.backward(gg)
    Lg.grad = Rg.grad = None
    og = ops.build_gwc_volume(Lg, Rg, Dq, G)
    close(og, gv)
    og.backward(gg.float().to(DEV))
    close(Lg.grad, L.grad)
    close(Rg.grad, R.grad)


@pytest.mark.parametrize("case", range(16))
def test_soft_argmin_and_fused_upsample_random(case):
    rng = np.random.default_rng(2000 + case)
    B, D = int(rng.integers(1, 3)), int(rng.choice([1, 2, 7, 8, 9, 48, 100, 192]))
    H, W = int(rng.integers(1, 9)), int(rng.choice([1, 2, 3, 4, 5, 8, 13, 32, 37]))
    scale = float(rng.choice([0.1, 1.0, 5.0, 30.0]))
    torch.manual_seed(case)
    cost = torch.randn(B, D, H, W) * scale
    c64 = cost.double().requires_grad_(True)
    d = torch.arange(D, dtype=torch.float64).view(1, -1, 1, 1)
    ref = (torch.softmax(c64, 1) * d).sum(1, keepdim=True)
    g = torch.randn(ref.shape)
    ref.backward(g.double())
    cg = cost.to(DEV).requires_grad_(True)
    out = ops.soft_argmin(cg)
    out.backward(g.to(DEV))
    assert float((out.detach().cpu().double() - ref.detach()).abs().max()) <= 2e-5
    close(cg.grad, c64.grad)
    # fused upsample: low-res logits [B,1,Dq,Hq,Wq] -> (f*Dq, fh*Hq, fw*Wq)
    Dq, Hq, Wq = int(rng.integers(2, 13)), int(rng.integers(1, 6)), int(rng.integers(1, 9))
    fd, fh, fw = int(rng.choice([1, 2, 3, 4])), int(rng.choice([1, 2, 4])), int(rng.choice([1, 3, 4]))
    size = (fd * Dq + int(rng.integers(0, 2)), fh * Hq, fw * Wq + int(rng.integers(0, 3)))
    low = torch.randn(B, 1, Dq, Hq, Wq) * scale
    l64 = low.double().requires_grad_(True)
    up = torch.nn.functional.interpolate(l64, size, mode="trilinear", align_corners=False).squeeze(1)
    dd = torch.arange(size[0], dtype=torch.float64).view(1, -1, 1, 1)
    ref = (torch.softmax(up, 1) * dd).sum(1, keepdim=True)
    g = torch.randn(ref.shape)
    ref.backward(g.double())
    lg = low.to(DEV).requires_grad_(True)
    out = ops.upsample_soft_argmin(lg, size)
    out.backward(g.to(DEV))
    assert float((out.detach().cpu().double() - ref.detach()).abs().max()) <= 1e-4, (size, (Dq, Hq, Wq))
    close(lg.grad, l64.grad, rtol=3e-5, floor=3e-5)


@pytest.mark.parametrize("case", range(24))
def test_reprojection_random(case):
    rng = np.random.default_rng(3000 + case)
    B, C = int(rng.integers(1, 3)), int(rng.choice([1, 1, 2, 3]))
    H, W = int(rng.integers(2, 40)), int(rng.choice([2, 3, 5, 8, 17, 31, 54, 55, 64, 109, 130]))
    ps = int(rng.choice([1, 3, 5, 7, 9, 11, 13, 15]))
    torch.manual_seed(case)
    L, R = torch.rand(B, C, H, W), torch.rand(B, C, H, W)
    disp = (torch.rand(B, 1, H, W) - 0.2) * float(rng.choice([2.0, W / 2, 2.0 * W]))
    mask = (torch.rand(B, 1, H, W) > 0.3) if rng.random() < 0.7 else None
    if mask is not None and not bool(mask.any()):
        mask[0, 0, 0, 0] = True
    d32 = disp.clone().requires_grad_(True)
    rl, rvis, rm = so.reproj_error_patch(L, R, d32, mask, ps=ps)
    rl.backward()
    dg = disp.to(DEV).requires_grad_(True)
    loss, vis, mi = az_rp.get_reproj_error_patch(L.to(DEV), R.to(DEV), dg, None if mask is None else mask.to(DEV), ps=ps)
    loss.backward()
    np.testing.assert_allclose(loss.item(), rl.item(), rtol=1e-5)
    close(vis, rvis)
    assert torch.equal(mi.cpu(), rm)
    close(dg.grad, d32.grad)
    # the stand-alone kernels agree with the fused pass
    l2, _ = ops.reproj_loss(L.to(DEV), R.to(DEV), disp.to(DEV), mask.to(DEV) if mask is not None else None, ps=ps)
    np.testing.assert_allclose(l2.item(), rl.item(), rtol=1e-5)
    close(ops.patch_fold(R.to(DEV), disp.to(DEV), ps), rvis)
    # warp + single-scale loss
    w = ops.warp(R.to(DEV), -disp.to(DEV))
    assert torch.equal(w.cpu(), so.apply_disparity(R, -disp))
    lo, wo, _ = az_rp.get_reprojection_error_old(L.to(DEV), R.to(DEV), disp.to(DEV), None if mask is None else mask.to(DEV))
    ro, rw, _ = so.reproj_error_old(L, R, disp, mask)
    np.testing.assert_allclose(lo.item(), ro.item(), rtol=1e-5)
    close(wo, rw)


@pytest.mark.parametrize("case", range(12))
def test_scatter_warp_lcn_tir_random(case):
    rng = np.random.default_rng(4000 + case)
    N, C = int(rng.integers(1, 3)), int(rng.integers(1, 4))
    H, W = int(rng.integers(1, 30)), int(rng.choice([1, 2, 5, 16, 33, 100, 257]))
    gen = torch.Generator().manual_seed(case)
    img = torch.rand(N, C, H, W, generator=gen) + 1
    d = torch.randint(0, max(1, min(192, 2 * W)), (N, 1, H, W), generator=gen, dtype=torch.int32)
    d = d * (1 if case % 2 == 0 else -1)
    assert torch.equal(az_wo.apply_disparity_cu(img.to(DEV), d.to(DEV)).cpu(), so.scatter_warp(img, d))
    ks = int(rng.choice([1, 3, 9, 11]))
    rn, rs = so.local_contrast_norm(img, ks)
    on, os_ = az_rp.local_contrast_norm(img.to(DEV), ks)
    close(on, rn, rtol=2e-5, floor=2e-5)
    close(os_, rs, rtol=2e-5, floor=2e-5)
    T = int(rng.choice([2, 4, 7, 9]))
    Hh, Ww = int(rng.integers(12, 40)), int(rng.choice([12, 16, 30, 47]))
    fr = rng.integers(0, 256, size=(T, Hh, Ww)).astype(np.uint8)
    fr[:, ::3, ::4] = np.clip(fr[:, ::3, ::4].astype(int) + np.arange(T)[:, None, None] * 9, 0, 255).astype(np.uint8)
    ref = so.temporal_ir_pattern(fr)
    pat = ops.temporal_ir_pattern(torch.from_numpy(fr).to(DEV)).cpu().numpy()
    assert (pat != ref).mean() <= 2e-3, (pat != ref).mean()


# Shapes chosen for the one-pass loss + Fold kernel (csrc/patch_loss_fold.cu): several strips per row with 1, 2 and
# 3 warp groups, several bands per image (rows shared between bands), a short last band, W % 4 != 0 (scalar
# staging), C > 1 with gradients accumulated over channels, smooth / sloped / out-of-range disparities, and a width
# whose shared-memory plan does not fit (falls back to the stand-alone loss and Fold kernels).
@pytest.mark.parametrize("shape,ps,kind", [
    ((1, 1, 40, 300, ), 5, "smooth"), ((2, 2, 70, 260), 13, "random"), ((1, 1, 64, 700), 7, "slope"),
    ((1, 1, 33, 1000), 11, "random"), ((1, 1, 200, 130), 11, "smooth"), ((3, 1, 150, 131), 9, "random"),
    ((1, 3, 37, 243), 3, "outside"), ((1, 1, 30, 1200), 11, "slope"), ((2, 1, 23, 480), 11, "edge"),
])
def test_patch_loss_fold_strips_bands_groups(shape, ps, kind):
    B, C, H, W = shape
    torch.manual_seed(H * W + ps)
    L, R = torch.rand(B, C, H, W), torch.rand(B, C, H, W)
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W).expand(B, 1, H, W)
    ys = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1).expand(B, 1, H, W)
    if kind == "smooth":
        disp = 7.3 + 0.01 * xs + 0.02 * ys
    elif kind == "slope":
        disp = 3.0 + 0.4 * xs - 0.1 * ys
    elif kind == "outside":
        disp = (torch.rand(B, 1, H, W) - 0.5) * 4.0 * W
    elif kind == "edge":  # samples landing in the boundary cells x0 = -1 and x0 = W-1
        disp = torch.where(torch.rand(B, 1, H, W) > 0.5, xs + 0.25, xs - (W - 1) + 0.25 - 1.0).clone()
    else:
        disp = torch.rand(B, 1, H, W) * min(64.0, W / 4)
    disp = disp.contiguous()
    mask = torch.rand(B, 1, H, W) > 0.3
    d32 = disp.clone().requires_grad_(True)
    rl, rvis, rm = so.reproj_error_patch(L, R, d32, mask, ps=ps)
    rl.backward()
    dg = disp.to(DEV).requires_grad_(True)
    loss, vis, mi = az_rp.get_reproj_error_patch(L.to(DEV), R.to(DEV), dg, mask.to(DEV), ps=ps)
    loss.backward()
    np.testing.assert_allclose(loss.item(), rl.item(), rtol=1e-5)
    close(vis, rvis)
    assert torch.equal(mi.cpu(), rm)
    close(dg.grad, d32.grad)
    # forward only (no gradient buffers: the other template instance), and determinism
    with torch.no_grad():
        l2, v2, _ = az_rp.get_reproj_error_patch(L.to(DEV), R.to(DEV), disp.to(DEV), mask.to(DEV), ps=ps)
        l3, v3, _ = az_rp.get_reproj_error_patch(L.to(DEV), R.to(DEV), disp.to(DEV), mask.to(DEV), ps=ps)
    assert l2.item() == l3.item() and torch.equal(v2, v3)
    np.testing.assert_allclose(l2.item(), rl.item(), rtol=1e-5)
    close(v2, rvis)


# channels_last_3d volume: same values bit for bit, only the strides differ; the gradient is
# consumed in place when it arrives in channels_last_3d and through the NCDHW kernel otherwise.
@pytest.mark.parametrize("case", range(12))
def test_concat_volume_channels_last_3d(case):
    rng = np.random.default_rng(7000 + case)
    B, H = int(rng.integers(1, 4)), int(rng.integers(1, 9))
    W = int(rng.choice([1, 3, 7, 31, 32, 33, 64, 100]))
    C = int(rng.choice([4, 8, 12, 32]))
    Dq = int(rng.choice([1, 2, 5, 12, 48, 70]))
    torch.manual_seed(case)
    L = torch.randn(B, C, H, W, dtype=torch.float64, requires_grad=True)
    R = torch.randn(B, C, H, W, dtype=torch.float64, requires_grad=True)
    vol = so.concat_volume(L, R, Dq)
    g = torch.randn(vol.shape, dtype=torch.float64)
    vol.backward(g)
    for grad_fmt in (torch.channels_last_3d, torch.contiguous_format):
        Lg = L.detach().float().to(DEV).requires_grad_(True)
        Rg = R.detach().float().to(DEV).requires_grad_(True)
        out = ops.build_concat_volume(Lg, Rg, Dq, channels_last=True)
        assert out.shape == vol.shape and out.is_contiguous(memory_format=torch.channels_last_3d)
        assert torch.equal(out.detach().cpu(), so.concat_volume(L.detach().float(), R.detach().float(), Dq))
        out.backward(g.float().to(DEV).contiguous(memory_format=grad_fmt))
        close(Lg.grad, L.grad)
        close(Rg.grad, R.grad)
    # C not a multiple of 4: torch converts the layout, values unchanged
    L3, R3 = torch.randn(1, 3, 4, 9, device=DEV), torch.randn(1, 3, 4, 9, device=DEV)
    o3 = ops.build_concat_volume(L3, R3, 5, channels_last=True)
    assert o3.is_contiguous(memory_format=torch.channels_last_3d)
    assert torch.equal(o3, ops.build_concat_volume(L3, R3, 5))

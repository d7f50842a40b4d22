"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the
same seeded inputs, against the golden fixtures produced by the real reference,
and -- at BASELINE.json's full sizes -- through size-independent properties.

Gates (BASELINE.json north_star): concat volume and scatter warp bit-exact;
soft-argmin <= 1e-4 px; everything else rtol 1e-5 (fp32), with an absolute floor
of 1e-5 x the tensor's max magnitude for element-wise comparisons.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import stereo_oracle as so  # noqa: E402  (checker only)

if torch.cuda.is_available():
    from activezero_b200 import ops
    from activezero_b200.tools import temporal_ir as az_tir
    from activezero_b200.utils import reprojection as az_rp
    from activezero_b200.utils import warp_ops as az_wo

DEV = "cuda:0"
T = torch.from_numpy


def close(a, b, rtol=1e-5, floor=1e-5):
    """The gate of every "rtol 1e-5" row:  |a - b| <= rtol * |b| + floor * max|b|.
    The second term is an ABSOLUTE floor relative to the tensor's largest magnitude: fp32 results carry ~6e-8 * max|b|
    of rounding noise on every element, so a purely relative test would fail on elements that are ~0 by cancellation
    (a gradient of 1e-9 next to gradients of 1) without saying anything about the kernel.  It is therefore a
    "1e-5 of the tensor's scale" gate for small elements and a true relative 1e-5 gate for elements of typical size;
    the failure message reports the worst error and the worst relative error so a regression shows which one moved."""
    a = a.detach().cpu().double()
    b = b.detach().cpu().double() if isinstance(b, torch.Tensor) else torch.as_tensor(b).double()
    atol = floor * float(b.abs().max()) + 1e-30 if b.numel() else 0.0  # 1e-30: fp32 underflows where fp64 holds 1e-100
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    assert bool((err <= tol).all()), f"max err {float(err.max()):.3e} (atol {atol:.3e}), worst rel {float((err / (b.abs() + atol + 1e-300)).max()):.3e}"


def gpu(x):
    return x.to(DEV)


# --------------------------------------------------------------------------- a1/a2 concat volume
@pytest.mark.parametrize("shape,dq", [((1, 4, 5, 12), 4), ((2, 3, 7, 16), 16), ((1, 2, 3, 10), 7), ((1, 2, 4, 8), 12),
                                      ((2, 32, 20, 60), 48), ((1, 1, 1, 4), 1), ((1, 3, 9, 135), 48)])
def test_concat_volume_fwd_bit_exact(shape, dq):
    torch.manual_seed(0)
    L, R = torch.randn(shape), torch.randn(shape)
    ref = so.concat_volume(L, R, dq)
    out = ops.build_concat_volume(gpu(L), gpu(R), dq)
    assert out.shape == ref.shape
    assert torch.equal(out.cpu(), ref)


def test_concat_volume_golden(golden):
    g = golden("psmnet_inline")
    out = ops.build_concat_volume(gpu(T(g["feat_L"])), gpu(T(g["feat_R"])), int(g["num_disp"]))
    assert np.array_equal(out.cpu().numpy(), g["vol"])


def test_concat_volume_config1_bit_exact():
    """BASELINE config 1: one 256x512 pair -> features [1,32,64,128], D=192."""
    torch.manual_seed(0)
    L, R = torch.randn(1, 32, 64, 128), torch.randn(1, 32, 64, 128)
    out = ops.build_concat_volume(gpu(L), gpu(R), 48)
    assert torch.equal(out.cpu(), so.concat_volume(L, R, 48))


def test_concat_volume_full_size_properties():
    """BASELINE config 2 shape (B=2 of the 8): [2,32,136,240], Dq=48.  The
    restatement is run on the GPU by stock torch as the checker."""
    torch.manual_seed(1)
    L, R = torch.randn(2, 32, 136, 240, device=DEV), torch.randn(2, 32, 136, 240, device=DEV)
    out = ops.build_concat_volume(L, R, 48)
    assert torch.equal(out, so.concat_volume(L, R, 48))
    # size-independent properties: plane 0 is the plain concat; column sums telescope
    assert torch.equal(out[:, :32, 0], L) and torch.equal(out[:, 32:, 0], R)
    assert float(out[:, :, 47, :, :47].abs().max()) == 0.0
    # linearity: vol(aL1+L2, ..) == a*vol(L1,..)+vol(L2,..) holds exactly for a power of two
    out2 = ops.build_concat_volume(2 * L, 2 * R, 48)
    assert torch.equal(out2, 2 * out)


@pytest.mark.parametrize("shape,dq", [((1, 4, 5, 12), 4), ((2, 3, 7, 16), 16), ((1, 2, 4, 8), 12), ((2, 8, 20, 60), 48),
                                      ((1, 3, 9, 135), 48), ((1, 2, 6, 1028), 24)])
def test_concat_volume_bwd(shape, dq):
    torch.manual_seed(2)
    L = torch.randn(shape, dtype=torch.float64, requires_grad=True)
    R = torch.randn(shape, dtype=torch.float64, requires_grad=True)
    vol = so.concat_volume(L, R, dq)
    g = torch.randn(vol.shape, dtype=torch.float64)
    vol.backward(g)
    Lg = gpu(L.detach().float()).requires_grad_(True)
    Rg = gpu(R.detach().float()).requires_grad_(True)
    ops.build_concat_volume(Lg, Rg, dq).backward(gpu(g.float()))
    close(Lg.grad, L.grad)
    close(Rg.grad, R.grad)


def test_concat_volume_bwd_full_size_and_partial_grads():
    torch.manual_seed(3)
    L = torch.randn(2, 32, 136, 240, device=DEV, requires_grad=True)
    R = torch.randn(2, 32, 136, 240, device=DEV, requires_grad=True)
    g = torch.randn(2, 64, 48, 136, 240, device=DEV)
    ops.build_concat_volume(L, R, 48).backward(g)
    # closed form with torch ops: triangular sums
    i = torch.arange(48, device=DEV).view(1, 1, 48, 1, 1)
    x = torch.arange(240, device=DEV).view(1, 1, 1, 1, 240)
    gl_ref = (g[:, :32].double() * (x >= i)).sum(2)
    close(L.grad, gl_ref)
    gr_ref = torch.zeros_like(gl_ref)
    for k in range(48):
        gr_ref[..., : 240 - k] += g[:, 32:, k, :, k:].double()
    close(R.grad, gr_ref)
    # only one side requires grad
    L2 = L.detach().clone().requires_grad_(True)
    ops.build_concat_volume(L2, R.detach(), 48).backward(g)
    assert torch.equal(L2.grad, L.grad)


# --------------------------------------------------------------------------- a3 gwc volume (parity unpinned)
@pytest.mark.parametrize("shape,dq,G", [((1, 8, 5, 12), 4, 2), ((2, 32, 6, 40), 12, 8), ((1, 6, 4, 9), 5, 3),
                                        ((1, 32, 10, 60), 48, 8), ((1, 16, 3, 20), 24, 16), ((1, 32, 4, 24), 8, 2)])
def test_gwc_volume_fwd_bwd(shape, dq, G):
    torch.manual_seed(4)
    L = torch.randn(shape, dtype=torch.float64, requires_grad=True)
    R = torch.randn(shape, dtype=torch.float64, requires_grad=True)
    vol = so.gwc_volume(L, R, dq, G)
    g = torch.randn(vol.shape, dtype=torch.float64)
    vol.backward(g)
    Lg = gpu(L.detach().float()).requires_grad_(True)
    Rg = gpu(R.detach().float()).requires_grad_(True)
    out = ops.build_gwc_volume(Lg, Rg, dq, G)
    close(out, vol)
    close(out, so.gwc_volume(L.detach().float(), R.detach().float(), dq, G))
    out.backward(gpu(g.float()))
    close(Lg.grad, L.grad)
    close(Rg.grad, R.grad)


# --------------------------------------------------------------------------- a4/a5 soft-argmin
def _sa_ref64(cost):
    c = cost.double()
    p = torch.softmax(c, 1)
    d = torch.arange(c.shape[1], dtype=torch.float64, device=c.device).view(1, -1, 1, 1)
    return (p * d).sum(1, keepdim=True)


def test_soft_argmin_golden(golden):
    g = golden("psmnet_inline")
    out = ops.soft_argmin(gpu(T(g["logits"])))
    assert np.abs(out.cpu().numpy() - g["pred"]).max() <= 1e-4
    g = golden("soft_argmin_synth")
    for tag in ("s1", "s10", "d96"):
        cost = gpu(T(g[f"{tag}_cost"])).requires_grad_(True)
        out = ops.soft_argmin(cost)
        out.backward(gpu(T(g[f"{tag}_g"])))
        assert np.abs(out.detach().cpu().numpy() - g[f"{tag}_pred"]).max() <= 1e-4
        close(cost.grad, g[f"{tag}_gcost"])


@pytest.mark.parametrize("shape,scale", [((1, 192, 8, 32), 1.0), ((2, 192, 5, 28), 10.0), ((1, 96, 7, 20), 50.0),
                                         ((1, 288, 4, 16), 3.0), ((1, 50, 3, 7), 2.0), ((1, 5, 2, 3), 1.0),
                                         ((1, 192, 9, 33), 0.01)])
def test_soft_argmin_fwd_bwd(shape, scale):
    torch.manual_seed(5)
    cost = torch.randn(shape) * scale
    ref32 = so.soft_argmin(cost)
    c64 = cost.double().requires_grad_(True)
    ref64 = _sa_ref64(c64)
    g = torch.randn(ref64.shape)
    ref64.backward(g.double())
    cg = gpu(cost).requires_grad_(True)
    out = ops.soft_argmin(cg)
    out.backward(gpu(g))
    assert float((out.detach().cpu() - ref32).abs().max()) <= 1e-4  # the north-star gate vs the fp32 reference
    assert float((out.detach().cpu().double() - ref64.detach()).abs().max()) <= 2e-5  # and far inside it vs exact
    close(cg.grad, c64.grad, rtol=1e-5, floor=1e-5)


@pytest.mark.parametrize("scale,offset", [(1e3, 0.0), (30.0, 1e6), (1e5, -3e7)])
def test_soft_argmin_large_magnitude_logits(scale, offset):
    """Accuracy must not depend on the magnitude of the logits (exponents are formed from the
    exact difference to the running max): huge / shifted logits, fwd and bwd, vs fp64."""
    torch.manual_seed(21)
    cost = torch.randn(1, 192, 6, 20) * scale + offset
    c64 = cost.double().requires_grad_(True)
    ref64 = _sa_ref64(c64)
    g = torch.randn(ref64.shape)
    ref64.backward(g.double())
    cg = gpu(cost).requires_grad_(True)
    out = ops.soft_argmin(cg)
    out.backward(gpu(g))
    assert float((out.detach().cpu().double() - ref64.detach()).abs().max()) <= 2e-5
    close(cg.grad, c64.grad, rtol=1e-5, floor=1e-5)


def test_soft_argmin_edge_values():
    cost = torch.zeros(1, 192, 4, 8)
    out = ops.soft_argmin(gpu(cost))
    assert float((out.cpu() - 95.5).abs().max()) <= 1e-4  # flat distribution -> mean of 0..191
    cost = torch.full((1, 192, 4, 8), -1e4)
    cost[:, 137] = 50.0
    assert float((ops.soft_argmin(gpu(cost)).cpu() - 137.0).abs().max()) == 0.0  # one-hot
    cost = torch.randn(1, 192, 4, 8)
    cost[:, :16] = float("-inf")  # leading -inf planes (masked disparities)
    ref = so.soft_argmin(cost)
    assert float((ops.soft_argmin(gpu(cost)).cpu() - ref).abs().max()) <= 1e-4


def test_soft_argmin_full_size():
    """BASELINE config 2 head (B=2 of 8): [2,192,544,960] logits.  Over 10^6 pixels the fp32
    reference's own rounding noise reaches ~1e-4 px at the tail, so the kernel is held to 2e-5 px
    against the exact (fp64) value -- 5x inside the gate -- and its distance to the fp32
    reference (stock torch on the GPU as the checker) must be explained by that reference's
    distance to exact."""
    torch.manual_seed(6)
    cost = torch.randn(2, 192, 544, 960, device=DEV) * 4.0
    out = ops.soft_argmin(cost)
    ref32 = so.soft_argmin(cost)
    ref64 = torch.cat([_sa_ref64(cost[b:b + 1]) for b in range(2)])
    mine = float((out.double() - ref64).abs().max())
    theirs = float((ref32.double() - ref64).abs().max())
    assert mine <= 2e-5, mine
    assert float((out - ref32).abs().max()) <= theirs + 2e-5
    assert float(((out - ref32).abs() <= 1e-4).float().mean()) >= 0.99
    # shift invariance (softmax property)
    out_shift = ops.soft_argmin(cost + 8.0)
    assert float((out_shift - out).abs().max()) <= 1e-4
    assert float(out.min()) >= 0.0 and float(out.max()) <= 191.0


# --------------------------------------------------------------------------- a6 warp
def _rand_disp(B, H, W, seed, lo=-8.0, hi=64.0):
    gen = torch.Generator().manual_seed(seed)
    return torch.rand(B, 1, H, W, generator=gen) * (hi - lo) + lo


@pytest.mark.parametrize("shape", [(2, 3, 24, 40), (1, 1, 64, 128), (1, 2, 17, 135), (1, 1, 2, 2)])
def test_warp_fwd_bit_exact_and_bwd(shape):
    torch.manual_seed(7)
    B, C, H, W = shape
    img = torch.rand(shape)
    disp = _rand_disp(B, H, W, 8, -W / 3, W / 3)
    ref = so.apply_disparity(img, disp)
    out = ops.warp(gpu(img), gpu(disp))
    assert torch.equal(out.cpu(), ref), float((out.cpu() - ref).abs().max())
    i64 = img.double().requires_grad_(True)
    d64 = disp.double().requires_grad_(True)
    g = torch.rand(shape)
    so.apply_disparity(i64, d64).backward(g.double())
    ig, dg = gpu(img).requires_grad_(True), gpu(disp).requires_grad_(True)
    ops.warp(ig, dg).backward(gpu(g))
    # fp32 sample positions differ from fp64 ones at ~1e-5 px: compare where bilinear is smooth
    close(ig.grad, i64.grad, rtol=1e-3, floor=1e-3)
    i32 = img.clone().requires_grad_(True)
    d32 = disp.clone().requires_grad_(True)
    so.apply_disparity(i32, d32).backward(g)
    close(dg.grad, d32.grad, rtol=1e-5, floor=1e-5)
    close(ig.grad, i32.grad, rtol=1e-5, floor=1e-5)


def test_warp_golden(golden):
    g = golden("reprojection")
    d = gpu(T(g["disp"])).requires_grad_(True)
    img = gpu(T(g["R3"])).requires_grad_(True)
    w = az_rp.apply_disparity(img, -d)
    w.backward(gpu(T(g["warp3_gout"])))
    assert np.array_equal(w.detach().cpu().numpy(), g["warp3"])
    close(d.grad, g["warp3_gdisp"])
    close(img.grad, g["warp3_gimg"])
    assert np.array_equal(az_rp.apply_disparity(gpu(T(g["L1"])), gpu(T(g["disp_r"]))).cpu().numpy(), g["warp1_pos"])


# --------------------------------------------------------------------------- a7/a8/a9 reprojection losses
@pytest.mark.parametrize("tag,img,masked", [("old1m", "1", True), ("old3m", "3", True), ("old1", "1", False)])
def test_reproj_old_golden(golden, tag, img, masked):
    g = golden("reprojection")
    d = gpu(T(g["disp"])).requires_grad_(True)
    loss, warped, mi = az_rp.get_reprojection_error_old(gpu(T(g["L" + img])), gpu(T(g["R" + img])), d,
                                                        gpu(T(g["mask"])) if masked else None)
    loss.backward()
    assert loss.dim() == 0 and mi.dtype == torch.int32
    np.testing.assert_allclose(loss.item(), g[f"{tag}_loss"], rtol=1e-5)
    close(warped, g[f"{tag}_warped"])
    assert np.array_equal(mi.cpu().numpy(), g[f"{tag}_mask"])
    close(d.grad, g[f"{tag}_gdisp"])


def test_reproj_bidir_golden(golden):
    g = golden("reprojection")
    dl = gpu(T(g["disp"])).requires_grad_(True)
    dr = gpu(T(g["disp_r"])).requires_grad_(True)
    ll, lr, wl, wr, ml, mr = az_rp.get_reprojection_error(gpu(T(g["L3"])), gpu(T(g["R3"])), dl, dr, gpu(T(g["mask"])),
                                                          gpu(T(g["mask_r"])))
    (ll + 2 * lr).backward()
    np.testing.assert_allclose(ll.item(), g["bi_loss_l"], rtol=1e-5)
    np.testing.assert_allclose(lr.item(), g["bi_loss_r"], rtol=1e-5)
    close(wl, g["bi_warp_l"])
    close(wr, g["bi_warp_r"])
    assert np.array_equal(ml.cpu().numpy(), g["bi_mask_l"]) and np.array_equal(mr.cpu().numpy(), g["bi_mask_r"])
    close(dl.grad, g["bi_gdisp_l"])
    close(dr.grad, g["bi_gdisp_r"])


def test_reproj_bidir_auto_masks():
    """mask_l=None: masks come from the scatter warp (reprojection.py:50-65)."""
    torch.manual_seed(9)
    L, R = torch.rand(2, 3, 20, 48), torch.rand(2, 3, 20, 48)
    dl, dr = torch.rand(2, 1, 20, 48) * 10, torch.rand(2, 1, 20, 48) * 10
    ref = so.reproj_error_bidir(L, R, dl, dr)
    out = az_rp.get_reprojection_error(gpu(L), gpu(R), gpu(dl), gpu(dr))
    np.testing.assert_allclose(out[0].item(), ref[0].item(), rtol=1e-5)
    np.testing.assert_allclose(out[1].item(), ref[1].item(), rtol=1e-5)
    assert torch.equal(out[4].cpu(), ref[4]) and torch.equal(out[5].cpu(), ref[5])


@pytest.mark.parametrize("tag,img,masked,ps", [("p5m", "1", True, 5), ("p11m", "1", True, 11), ("p11", "1", False, 11),
                                                ("p3c3m", "3", True, 3), ("p1m", "1", True, 1)])
def test_reproj_patch_golden(golden, tag, img, masked, ps):
    g = golden("reprojection")
    d = gpu(T(g["disp"])).requires_grad_(True)
    loss, vis, mi = az_rp.get_reproj_error_patch(gpu(T(g["L" + img])), gpu(T(g["R" + img])), d,
                                                 gpu(T(g["mask"])) if masked else None, ps=ps)
    loss.backward()
    np.testing.assert_allclose(loss.item(), g[f"{tag}_loss"], rtol=1e-5)
    close(vis, g[f"{tag}_vis"])
    assert np.array_equal(mi.cpu().numpy(), g[f"{tag}_mask"])
    close(d.grad, g[f"{tag}_gdisp"])


@pytest.mark.parametrize("tag,ps", [("w11", 11), ("w7", 7)])
def test_reproj_patch_wide_golden(golden, tag, ps):
    """The real reference's get_reproj_error_patch on frames several strips / bands wide (reprojection_wide.npz)."""
    g = golden("reprojection_wide")
    d = gpu(T(g[f"{tag}_disp"])).requires_grad_(True)
    mask = gpu(T(g[f"{tag}_maskin"])) if f"{tag}_maskin" in g.files else None
    loss, vis, _ = az_rp.get_reproj_error_patch(gpu(T(g[f"{tag}_L"]).float()), gpu(T(g[f"{tag}_R"]).float()), d, mask, ps=ps)
    loss.backward()
    np.testing.assert_allclose(loss.item(), g[f"{tag}_loss"], rtol=1e-5)
    close(vis, g[f"{tag}_vis"])
    close(d.grad, g[f"{tag}_gdisp"])


@pytest.mark.parametrize("shape,ps", [((1, 1, 32, 64), 11), ((2, 1, 19, 45), 7), ((1, 2, 12, 33), 13), ((1, 1, 6, 5), 11)])
def test_reproj_patch_random(shape, ps):
    torch.manual_seed(10)
    B, C, H, W = shape
    L, R = (torch.rand(shape) > 0.5).float(), (torch.rand(shape) > 0.5).float()  # Bernoulli IR patterns
    disp = _rand_disp(B, H, W, 11, -4.0, W / 2)
    mask = torch.rand(B, 1, H, W) > 0.3
    d32 = disp.clone().requires_grad_(True)
    rl, rvis, rm = so.reproj_error_patch(L, R, d32, mask, ps=ps)
    rl.backward()
    dg = gpu(disp).requires_grad_(True)
    loss, vis, mi = az_rp.get_reproj_error_patch(gpu(L), gpu(R), dg, gpu(mask), ps=ps)
    loss.backward()
    np.testing.assert_allclose(loss.item(), rl.item(), rtol=1e-5)
    close(vis, rvis)
    assert torch.equal(mi.cpu(), rm)
    close(dg.grad, d32.grad)


def test_reproj_patch_empty_mask_nan_and_flags(golden):
    g = golden("reprojection")
    L, R, d = gpu(T(g["L1"])), gpu(T(g["R1"])), gpu(T(g["disp"]))
    loss, _, _ = az_rp.get_reproj_error_patch(L, R, d, torch.zeros(2, 1, 24, 40, dtype=torch.bool, device=DEV), ps=5)
    assert torch.isnan(loss)
    az_rp.RETURN_WARPED_PATCH_IMAGE = False
    try:
        loss2, vis, _ = az_rp.get_reproj_error_patch(L, R, d, None, ps=5)
        assert vis is None and torch.isfinite(loss2)
    finally:
        az_rp.RETURN_WARPED_PATCH_IMAGE = True


def test_reproj_full_size_properties():
    """544x960 (config 2 frame).  The CPU oracle checks one full-size pair;
    exact properties cover the batch: the loss scales by
    a^2 when both images scale by a = 2, and the disparity gradient is supported
    on the mask only."""
    torch.manual_seed(12)
    L = (torch.rand(2, 1, 544, 960, device=DEV) > 0.5).float()
    R = (torch.rand(2, 1, 544, 960, device=DEV) > 0.5).float()
    d = torch.rand(2, 1, 544, 960, device=DEV) * 64
    l1, _, _ = az_rp.get_reproj_error_patch(L, R, d, None, ps=11)
    l2, _, _ = az_rp.get_reproj_error_patch(2 * L, 2 * R, d, None, ps=11)
    np.testing.assert_allclose(l2.item(), 4 * l1.item(), rtol=1e-6)
    mask = torch.rand(2, 1, 544, 960, device=DEV) > 0.5
    mask[:, :, :100] = False
    dg = d.clone().requires_grad_(True)
    loss, vis, _ = az_rp.get_reproj_error_patch(L, R, dg, mask, ps=11)
    loss.backward()
    assert float(dg.grad[~mask].abs().max()) == 0.0 and float(dg.grad[mask].abs().max()) > 0.0
    # one full-size pair against the CPU oracle (same fp32 sample positions bit for bit)
    dr = d[:1].cpu().clone().requires_grad_(True)
    rl, rvis, _ = so.reproj_error_patch(L[:1].cpu(), R[:1].cpu(), dr, mask[:1].cpu(), ps=11)
    rl.backward()
    d0 = d[:1].clone().requires_grad_(True)
    l0, v0, _ = az_rp.get_reproj_error_patch(L[:1], R[:1], d0, mask[:1], ps=11)
    l0.backward()
    np.testing.assert_allclose(l0.item(), rl.item(), rtol=1e-5)
    close(v0, rvis)
    close(d0.grad, dr.grad)


def test_reproj_diff_ratio_golden(golden):
    g = golden("reprojection")
    d = gpu(T(g["disp"])).requires_grad_(True)
    tot, stages, ld = az_rp.get_reprojection_error_diff_ratio(gpu(T(g["L3"])), gpu(T(g["R3"])), d, gpu(T(g["mask"])))
    tot.backward()
    np.testing.assert_allclose(tot.item(), g["ms_loss"], rtol=1e-5)
    close(d.grad, g["ms_gdisp"], rtol=1e-4, floor=1e-4)  # F.interpolate backward on GPU uses atomics
    for k in range(3):
        np.testing.assert_allclose(ld[f"stage{k}"], g[f"ms_stage{k}_loss"], rtol=1e-5)
        close(stages[f"stage{k}"]["warped"], g[f"ms_stage{k}_warped"])
        assert np.array_equal(stages[f"stage{k}"]["mask"].cpu().numpy(), g[f"ms_stage{k}_mask"])


def test_reproj_image_gradients_use_the_warp_kernel():
    torch.manual_seed(13)
    L, R = torch.rand(1, 3, 16, 32), torch.rand(1, 3, 16, 32)
    d = torch.rand(1, 1, 16, 32) * 6
    R32 = R.clone().requires_grad_(True)
    ref, _, _ = so.reproj_error_old(L, R32, d)
    ref.backward()
    Rg = gpu(R).requires_grad_(True)
    loss, _, _ = az_rp.get_reprojection_error_old(gpu(L), Rg, gpu(d))
    loss.backward()
    np.testing.assert_allclose(loss.item(), ref.item(), rtol=1e-5)
    close(Rg.grad, R32.grad)


# --------------------------------------------------------------------------- a10 scatter warp
def test_scatter_warp_golden(golden):
    g = golden("scatter_warp")
    img = gpu(T(g["img"]))
    assert np.array_equal(az_wo.apply_disparity_cu(img, gpu(T(g["disp_pos"]))).cpu().numpy(), g["out_pos"])
    assert np.array_equal(az_wo.apply_disparity_cu(img, gpu(T(g["disp_neg"]))).cpu().numpy(), g["out_neg"])
    assert np.array_equal(az_wo.apply_disparity_cu(gpu(T(g["self_img"])), gpu(T(g["self_disp"]))).cpu().numpy(),
                          g["self_out"])


@pytest.mark.parametrize("shape", [(2, 1, 64, 240), (1, 3, 37, 135), (8, 1, 544, 960), (1, 1, 1, 1)])
@pytest.mark.parametrize("sign", [1, -1])
def test_scatter_warp_random_bit_exact(shape, sign):
    gen = torch.Generator().manual_seed(14)
    N, C, H, W = shape
    img = torch.rand(shape, generator=gen) * 100 + 1
    d = torch.randint(0, min(192, W), (N, 1, H, W), generator=gen, dtype=torch.int32)
    d[torch.rand(d.shape, generator=gen) < 0.3] = 0
    d = d * sign
    ref = so.scatter_warp(img, d)
    out = az_wo.apply_disparity_cu(gpu(img), gpu(d))
    assert torch.equal(out.cpu(), ref)
    out3 = az_wo.apply_disparity_cu(gpu(img), gpu(d.view(N, H, W)))  # (N,H,W) disparity is accepted too
    assert torch.equal(out3.cpu(), ref)


def test_scatter_warp_asserts_like_the_reference():
    img = torch.rand(1, 1, 4, 8, device=DEV)
    d = torch.zeros(1, 1, 4, 8, dtype=torch.int32, device=DEV)
    d[0, 0, 0, 0], d[0, 0, 1, 5] = 2, -2
    with pytest.raises(AssertionError):
        az_wo.apply_disparity_cu(img, d)
    with pytest.raises(AssertionError):
        az_wo.apply_disparity_cu(img, d.float())
    with pytest.raises(AssertionError):
        az_wo.apply_disparity_cu(img.cpu(), d.cpu())
    with pytest.raises(AssertionError):
        az_wo.apply_disparity_cu(img.expand(2, 1, 4, 8).transpose(2, 3), d)
    assert torch.equal(az_wo.apply_disparity_cu(img, torch.zeros_like(d)), img)  # all-zero disparity -> identity


# --------------------------------------------------------------------------- a11 temporal IR, a12 LCN
def test_temporal_ir_golden(golden):
    g = golden("temporal_ir")
    for side in ("L", "R"):
        pat = az_tir.extract_temporal_ir_pattern(gpu(T(g[f"frames_{side}"])))
        assert pat.shape == g[f"pattern_{side}"].shape
        assert np.array_equal(pat.cpu().numpy().astype(np.float64), g[f"pattern_{side}"])


@pytest.mark.parametrize("T_,H,W", [(7, 96, 128), (4, 60, 83), (7, 720, 1280), (9, 33, 47)])
def test_temporal_ir_random(T_, H, W):
    rng = np.random.default_rng(15)
    base = rng.integers(0, 256, size=(H, W)) * 0.5
    dots = rng.random((H, W)) < 0.1
    fr = np.stack([base + t * dots * 10.0 for t in range(T_)]) + rng.integers(0, 6, size=(T_, H, W))
    fr = np.clip(fr, 0, 255).astype(np.uint8)
    ref = so.temporal_ir_pattern(fr)
    batch = torch.from_numpy(np.stack([fr, fr[::-1].copy()]))
    pat = az_tir.extract_temporal_ir_pattern(gpu(batch)).cpu().numpy()
    # gate: exact except pixels within 1e-6 of the threshold -> allow a vanishing mismatch rate
    assert (pat[0] != ref).mean() <= 1e-5
    assert (pat[1] != so.temporal_ir_pattern(fr[::-1].copy())).mean() <= 1e-5
    assert 0.01 < pat[0].mean() < 0.6


def test_lcn(golden):
    g = golden("reprojection")
    img = gpu(T(g["lcn_img"]))
    n9, s9 = az_rp.local_contrast_norm(img, 9)
    n5, s5 = az_rp.local_contrast_norm(img, 5, eps=1e-3)
    close(n9, g["lcn9_norm"])
    close(s9, g["lcn9_std"])
    close(n5, g["lcn5_norm"])
    close(s5, g["lcn5_std"])
    big = torch.rand(2, 1, 544, 960)
    rn, rs = so.local_contrast_norm(big, 9)
    on, os_ = az_rp.local_contrast_norm(gpu(big), 9)
    close(on, rn)
    close(os_, rs)


# --------------------------------------------------------------------------- plumbing
def test_runs_on_a_side_stream_and_counts_launches():
    from activezero_b200 import _lib

    torch.manual_seed(16)
    L, R = torch.randn(1, 8, 16, 64, device=DEV), torch.randn(1, 8, 16, 64, device=DEV)
    ref = so.concat_volume(L, R, 12)
    before = _lib.launch_count
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        out = ops.build_concat_volume(L, R, 12)
    s.synchronize()
    assert torch.equal(out, ref)
    assert _lib.launch_count == before + 1


def test_int64_indexing_above_2_31_elements():
    """BASELINE config 5 top end: [4,64,72,272,480] = 2.4e9 elements (9.6 GB) needs 64-bit
    offsets; check the LAST pair (highest addresses) bit-exactly and the soft-argmin of a
    [4,288,1088,1920]-sized tensor on its last image."""
    torch.manual_seed(20)
    L, R = torch.randn(4, 32, 272, 480, device=DEV), torch.randn(4, 32, 272, 480, device=DEV)
    vol = ops.build_concat_volume(L, R, 72)
    assert vol.numel() > 2 ** 31
    assert torch.equal(vol[3:], so.concat_volume(L[3:], R[3:], 72))
    g = torch.zeros_like(vol)
    g[3, :, :, 100:104] = 1.0
    gL, gR = torch.empty_like(L), torch.empty_like(R)
    from activezero_b200 import _lib
    _lib.call("az_concat_volume_bwd", ops._ptr(g), ops._ptr(gL), ops._ptr(gR), 4, 32, 272, 480, 72, ops._stream())
    assert float(gL[:3].abs().max()) == 0.0 and float(gR[:3].abs().max()) == 0.0
    x = torch.arange(480, device=DEV)
    assert torch.equal(gL[3, 0, 101], torch.minimum(x + 1, torch.tensor(72, device=DEV)).float())
    assert torch.equal(gR[3, 5, 102], torch.minimum(480 - x, torch.tensor(72, device=DEV)).float())
    del vol, g
    cost = torch.randn(4, 288, 1088, 1920, device=DEV)
    assert cost.numel() > 2 ** 31
    out = ops.soft_argmin(cost)
    # outputs reach 287 here: half an fp32 ulp is already 1.5e-5
    assert float((out[3:].double() - _sa_ref64(cost[3:, :, :, :])).abs().max()) <= 4e-5


# --------------------------------------------------------------------------- §8f-1 fused upsample + soft-argmin
def _upsample_ref(low, size, dtype=torch.float32):
    """psmnet.py:186-217 for one head, restated with the oracle's soft-argmin."""
    up = torch.nn.functional.interpolate(low.to(dtype), size, mode="trilinear", align_corners=False)
    return so.soft_argmin(torch.squeeze(up, 1)) if dtype == torch.float32 else _sa_ref64(torch.squeeze(up, 1))


@pytest.mark.parametrize("shape,size,scale", [((1, 1, 12, 8, 16), (48, 32, 64), 3.0), ((2, 1, 48, 16, 32), (192, 64, 128), 8.0),
                                               ((1, 1, 5, 7, 9), (20, 28, 36), 1.0), ((1, 1, 6, 5, 7), (17, 13, 23), 5.0),
                                               ((1, 1, 4, 3, 3), (4, 3, 3), 2.0)])
def test_upsample_soft_argmin_fwd_bwd(shape, size, scale):
    torch.manual_seed(30)
    low = torch.randn(shape) * scale
    l64 = low.double().requires_grad_(True)
    ref64 = _upsample_ref(l64, size, torch.float64)
    g = torch.randn(ref64.shape)
    ref64.backward(g.double())
    ref32 = _upsample_ref(low, size)  # torch CPU: the reference's own path
    lg = gpu(low).requires_grad_(True)
    out = ops.upsample_soft_argmin(lg, size)
    out.backward(gpu(g))
    assert out.shape == ref32.shape
    mine = float((out.detach().cpu().double() - ref64.detach()).abs().max())
    theirs = float((ref32.double() - ref64.detach()).abs().max())
    assert mine <= 1e-4, mine  # the gate, against the exact value (fp32 interpolation noise included)
    assert float((out.detach().cpu() - ref32).abs().max()) <= 1e-4 + theirs
    close(lg.grad, l64.grad, rtol=2e-5, floor=2e-5)


def test_upsample_soft_argmin_matches_unfused_full_size():
    """One 544x960 pair: fused kernel vs F.interpolate + this repo's soft-argmin (stock torch
    interpolate on the GPU as the checker), forward and backward."""
    torch.manual_seed(31)
    low = (torch.randn(1, 1, 48, 136, 240, device=DEV) * 6).requires_grad_(True)
    low2 = low.detach().clone().requires_grad_(True)
    g = torch.randn(1, 1, 544, 960, device=DEV)
    out = ops.upsample_soft_argmin(low, (192, 544, 960))
    up = torch.nn.functional.interpolate(low2, (192, 544, 960), mode="trilinear", align_corners=False)
    ref = ops.soft_argmin(torch.squeeze(up, 1))
    # torch's fp32 interpolation and the in-kernel one round the logits differently (~1e-6), which
    # peaky pixels amplify; both sit within 1e-4 px of the fp64 value
    ref64 = _sa_ref64(torch.nn.functional.interpolate(low.detach().double(), (192, 544, 960), mode="trilinear",
                                                        align_corners=False).squeeze(1))
    assert float((out.detach().double() - ref64).abs().max()) <= 1e-4
    assert float((out.detach() - ref.detach()).abs().max()) <= 1e-4 + float((ref.detach().double() - ref64).abs().max())
    out.backward(g)
    ref.backward(g)
    close(low.grad, low2.grad, rtol=1e-4, floor=1e-4)  # torch's trilinear backward accumulates with float atomics


def test_psmnet_fused_upsample_flag():
    from activezero_b200.nets.psmnet.psmnet_3 import PSMNet

    torch.manual_seed(1)
    net = PSMNet(maxdisp=192).cuda().train()
    a, b = torch.rand(2, 3, 256, 256, device=DEV), torch.rand(2, 3, 256, 256, device=DEV)
    with torch.no_grad():
        p_ref = net(a, b)
        p_ref2 = net(a, b)
        net.fuse_upsample = True
        p_fused = net(a, b)
    assert all(torch.isfinite(p).all() for p in p_fused)
    for r, r2, f in zip(p_ref, p_ref2, p_fused):  # cuDNN run-to-run noise in the logits is the yardstick
        assert float((r - f).abs().max()) <= 3.0 * float((r - r2).abs().max()) + 2e-4


# --------------------------------------------------------------------------- §8f-4 fused error metrics
def test_err_metrics_golden_and_random(golden):
    from activezero_b200.utils.cascade_metrics import compute_err_metric

    g = golden("err_metrics")
    args = [gpu(T(g[k])) for k in ("disp_gt", "depth_gt", "disp_pred", "focal", "base", "mask")]
    m1 = compute_err_metric(*args)
    m2 = compute_err_metric(*args, depth_pred=gpu(T(g["depth_pred"])))
    for tag, m in (("m1_", m1), ("m2_", m2)):
        assert set(m) == {"epe", "bad1", "bad2", "depth_abs_err", "depth_err2", "depth_err4", "depth_err8"}
        for k, v in m.items():
            assert isinstance(v, float)
            if k in ("epe", "depth_abs_err"):
                np.testing.assert_allclose(v, float(g[tag + k]), rtol=1e-6)
            else:
                assert v == float(g[tag + k]), k  # counts are exact
    torch.manual_seed(40)
    B, H, W = 8, 544, 960
    disp_gt = torch.rand(B, 1, H, W) * 100 + 1
    disp_pred = disp_gt + torch.randn(B, 1, H, W) * 2
    focal, base = torch.rand(B, 1, 1, 1) * 50 + 400, torch.rand(B, 1, 1, 1) * 0.01 + 0.05
    depth_gt = focal * base / disp_gt + torch.randn(B, 1, H, W) * 2e-3
    mask = torch.rand(B, 1, H, W) > 0.4
    ref = so.compute_err_metric(disp_gt, depth_gt, disp_pred, focal, base, mask)
    out = compute_err_metric(gpu(disp_gt), gpu(depth_gt), gpu(disp_pred), gpu(focal), gpu(base), gpu(mask))
    for k in ref:
        if k in ("epe", "depth_abs_err"):
            np.testing.assert_allclose(out[k], ref[k], rtol=1e-5)
        else:
            assert out[k] == ref[k], k
    empty = compute_err_metric(gpu(disp_gt), gpu(depth_gt), gpu(disp_pred), gpu(focal), gpu(base),
                               torch.zeros_like(mask).to(DEV))
    assert all(np.isnan(v) for v in empty.values())


def test_ops_inside_autocast_compute_in_fp32():
    torch.manual_seed(50)
    cost = torch.randn(1, 48, 8, 16, device=DEV, requires_grad=True)
    L = torch.randn(1, 8, 6, 16, device=DEV, requires_grad=True)
    R = torch.randn(1, 8, 6, 16, device=DEV)
    ref = ops.soft_argmin(cost)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = ops.soft_argmin(cost.to(torch.bfloat16).float() * 0 + cost)  # fp32 stays fp32
        half_in = ops.soft_argmin(cost.half())                              # half input is cast up, not rejected
        vol = ops.build_concat_volume(L, R, 4)
    assert out.dtype == torch.float32 and torch.equal(out, ref)
    assert half_in.dtype == torch.float32
    assert float((half_in - ops.soft_argmin(cost.half().float())).detach().abs().max()) == 0.0
    assert vol.dtype == torch.float32 and torch.equal(vol, so.concat_volume(L.detach(), R, 4))
    (out.sum() + vol.sum()).backward()
    assert cost.grad is not None and L.grad is not None and cost.grad.dtype == torch.float32


# --------------------------------------------------------------------------- §8f-3 sim IR pattern
def test_sim_ir_pattern_golden_and_random(golden):
    from activezero_b200.datasets import dataset_utils as az_du

    g = golden("sim_ir_pattern")
    for tag in ("a", "b", "c"):
        ir, img = gpu(T(g[f"{tag}_ir_u8"])), gpu(T(g[f"{tag}_img_u8"]))
        assert np.array_equal(az_du.get_ir_pattern(ir, img).cpu().numpy(), g[f"{tag}_p1"])
        assert np.array_equal(az_du.get_smoothed_ir_pattern2(ir, img).cpu().numpy(), g[f"{tag}_p2"])
        assert np.array_equal(az_du.get_smoothed_ir_pattern2(ir, img, ks=5, threshold=0.01).cpu().numpy(), g[f"{tag}_p2_k5"])
        # float64 inputs (already /255) take the same path
        ir64, img64 = gpu(T(g[f"{tag}_ir_u8"] / 255)), gpu(T(g[f"{tag}_img_u8"] / 255))
        assert np.array_equal(az_du.get_smoothed_ir_pattern2(ir64, img64).cpu().numpy(), g[f"{tag}_p2"])
    rng = np.random.default_rng(60)
    for (h, w, ks) in ((540, 960, 11), (256, 512, 11), (121, 242, 11), (90, 75, 5)):
        base = rng.integers(0, 200, size=(2, h, w))
        ir = np.clip(base + (rng.random((2, h, w)) < 0.1) * 40 + rng.integers(0, 3, size=(2, h, w)), 0, 255).astype(np.uint8)
        img = base.astype(np.uint8)
        out = az_du.get_smoothed_ir_pattern2(gpu(T(ir)), gpu(T(img)), ks=ks).cpu().numpy()
        for b in range(2):
            ref = so.smoothed_ir_pattern2(ir[b] / 255, img[b] / 255, ks=ks)
            assert np.array_equal(out[b], ref), (h, w, ks, float((out[b] != ref).mean()))


def test_baseline_config1_cost_volume_and_regression():
    """BASELINE config 1, verbatim: cost volume + disparity regression forward on one synthetic 256x512
    pair, maxdisp 192, against the reference's torch path on the CPU (L,R = randn(1,32,64,128,
    seed 0), cost = randn(1,192,256,512))."""
    torch.manual_seed(0)
    L, R = torch.randn(1, 32, 64, 128), torch.randn(1, 32, 64, 128)
    cost = torch.randn(1, 192, 256, 512)
    vol = ops.build_concat_volume(gpu(L), gpu(R), 192 // 4)
    assert torch.equal(vol.cpu(), so.concat_volume(L, R, 48))          # bit-exact
    disp = ops.soft_argmin(gpu(cost))
    ref = so.soft_argmin(cost)                                          # torch CPU, fp32
    ref64 = _sa_ref64(cost)
    assert float((disp.cpu().double() - ref64).abs().max()) <= 2e-5
    assert float((disp.cpu() - ref).abs().max()) <= 1e-4


def test_hot_path_is_cuda_graph_capturable():
    """The C-ABI calls only enqueue work on the current stream (no allocation, no sync), so the whole
    forward hot path can be captured once in a CUDA graph and replayed (launch-bound small batches)."""
    torch.manual_seed(70)
    L, R = torch.randn(1, 32, 16, 64, device=DEV), torch.randn(1, 32, 16, 64, device=DEV)
    cost = torch.randn(1, 48, 64, 256, device=DEV)
    pL = (torch.rand(1, 1, 64, 256, device=DEV) > 0.5).float()
    pR = (torch.rand(1, 1, 64, 256, device=DEV) > 0.5).float()

    def step():
        vol = ops.build_concat_volume(L, R, 12)
        disp = ops.soft_argmin(cost)
        loss, vis = ops.reproj_loss(pL, pR, disp, None, ps=11, want_warped=True)
        return vol, disp, loss, vis

    with torch.no_grad():
        eager = [t.clone() for t in step()]  # also warms the linspace-table cache (an H2D copy)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            outs = step()
        for _ in range(3):
            cost.normal_()                      # new inputs in the captured buffers
            g.replay()
        torch.cuda.synchronize()
        ref = step()
    for a, b in zip(outs, ref):
        assert torch.equal(a, b)
    assert not torch.equal(outs[1], eager[1])   # the replay really consumed the new logits

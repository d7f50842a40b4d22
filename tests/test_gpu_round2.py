"""GPU parity tests added in round 2: the multi-scale loss's rescaling kernel (a9), cross-checks that tie the
group-wise correlation volume (a3, no reference operator) to pinned rows, argument validation of the
torch.library ops, and every kernel variant that `AZ_*` tuning knobs can select."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import stereo_oracle as so  # noqa: E402  (checker only)

if torch.cuda.is_available():
    from activezero_b200 import library_ops, ops
    from activezero_b200.utils import reprojection as az_rp

DEV = "cuda:0"


def close(a, b, rtol=1e-5, floor=1e-5):
    """|a - b| <= rtol * |b| + floor * max|b| (see tests/test_gpu_parity.py::close for why the floor is there)."""
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    atol = floor * float(b.abs().max()) + 1e-30
    err = (a - b).abs()
    assert bool((err <= atol + rtol * b.abs()).all()), f"max err {float(err.max()):.3e} (atol {atol:.3e})"


class _env:
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


# --------------------------------------------------------------------------- a9 rescaling (reprojection.py:153-158)
@pytest.mark.parametrize("shape", [(2, 1, 64, 96), (1, 3, 50, 70), (1, 1, 136, 240)])
@pytest.mark.parametrize("r", [0.25, 0.5, 1])
def test_rescale_matches_interpolate(shape, r):
    """Bit-identical to what the reference executes -- F.interpolate(scale_factor=r, mode="bilinear") on the GPU
    (stock torch as the checker) -- and within one rounding of torch's CPU kernel, which switches between two
    evaluation orders depending on the output size."""
    torch.manual_seed(3)
    B, C, H, W = shape
    L, R = torch.randn(shape), torch.randn(shape)
    d = torch.rand(B, 1, H, W) * 40
    m = torch.rand(B, 1, H, W) > 0.6
    Lg, Rg, dg, mg = L.to(DEV), R.to(DEV), d.to(DEV), m.to(DEV)
    Lr, Rr, dr, mr = ops.rescale_for_loss(Lg, Rg, dg, mg, r)
    assert torch.equal(Lr, F.interpolate(Lg, scale_factor=r, mode="bilinear"))
    assert torch.equal(Rr, F.interpolate(Rg, scale_factor=r, mode="bilinear"))
    assert torch.equal(dr, F.interpolate(dg, scale_factor=r, mode="bilinear") * r)
    assert torch.equal(mr.cpu(), F.interpolate(m.float(), scale_factor=r, mode="bilinear").type(torch.bool))
    for mine, src in ((Lr, L), (Rr, R)):
        cpu = F.interpolate(src, scale_factor=r, mode="bilinear")
        assert float((mine.cpu() - cpu).abs().max()) <= 2.4e-7 * float(cpu.abs().max())
    _, _, _, m_none = ops.rescale_for_loss(Lg, Rg, dg, None, r)
    assert bool(m_none.all())


@pytest.mark.parametrize("r", [0.25, 0.5, 1])
def test_rescale_disp_gradient(r):
    torch.manual_seed(4)
    d = (torch.rand(2, 1, 36, 52) * 30).double().requires_grad_(True)
    ref = F.interpolate(d, scale_factor=r, mode="bilinear") * r
    g = torch.randn(ref.shape)
    ref.backward(g.double())
    img = torch.randn(2, 1, 36, 52, device=DEV)
    dg = d.detach().float().to(DEV).requires_grad_(True)
    _, _, dr, _ = ops.rescale_for_loss(img, img, dg, None, r)
    dr.backward(g.to(DEV))
    close(dg.grad, d.grad)


def test_multiscale_loss_matches_oracle_with_tight_gradient():
    """a9 end to end: loss, per-stage outputs and d loss / d disp (the rescale adjoint is a deterministic gather,
    so the gate is rtol 1e-5 again -- round 1's 1e-4 came from torch's atomics in F.interpolate's backward)."""
    torch.manual_seed(7)
    L, R = torch.rand(2, 1, 64, 128), torch.rand(2, 1, 64, 128)
    d = (torch.rand(2, 1, 64, 128) * 20).requires_grad_(True)
    m = torch.rand(2, 1, 64, 128) > 0.3
    ref, ref_out, ref_dict = so.reproj_error_diff_ratio(L, R, d, m)
    ref.backward()
    dg = d.detach().to(DEV).requires_grad_(True)
    loss, out, ld = az_rp.get_reprojection_error_diff_ratio(L.to(DEV), R.to(DEV), dg, m.to(DEV))
    loss.backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    for k in ref_out:
        assert torch.equal(out[k]["mask"].cpu(), ref_out[k]["mask"])
        close(out[k]["pred_disp"], ref_out[k]["pred_disp"], floor=0.0, rtol=2.4e-7)
        close(out[k]["warped"], ref_out[k]["warped"])
        assert abs(ld[k] - ref_dict[k]) <= 1e-5 * abs(ref_dict[k]) + 1e-12
    close(dg.grad, d.grad.double())


# --------------------------------------------------------------------------- a3 gwc: cross-checks against pinned rows
def test_gwc_reduces_to_pinned_operators():
    """The reference has no gwc operator (parity unpinned).  Two identities tie it to rows that ARE pinned:
    (1) G = C: vol[b,c,i] = L[b,c] * shift_i(R[b,c]) -- the product of the two halves of the (bit-exact) concat volume;
    (2) L = 1: vol[b,g,i] = mean over the group of the concat volume's right half."""
    torch.manual_seed(11)
    B, C, H, W, Dq = 2, 32, 20, 64, 24
    L, R = torch.randn(B, C, H, W, device=DEV), torch.randn(B, C, H, W, device=DEV)
    cat = ops.build_concat_volume(L, R, Dq)
    full = ops.build_gwc_volume(L, R, Dq, C)
    assert torch.equal(full, cat[:, :C] * cat[:, C:])
    ones = torch.ones_like(L)
    g8 = ops.build_gwc_volume(ones, R, Dq, 8)
    cat1 = ops.build_concat_volume(ones, R, Dq)
    ref = cat1[:, C:].reshape(B, 8, C // 8, Dq, H, W).double().mean(2)
    close(g8, ref)


@pytest.mark.parametrize("knobs", [dict(AZ_GWC_DG=0, AZ_GWC_BWD=0), dict(AZ_GWC_DG=8, AZ_GWC_BWD=1), dict(AZ_GWC_DG=16, AZ_GWC_BWD=1),
                                   dict(AZ_GWC_DG=16, AZ_GWC_BWD=2), dict(AZ_GWC_BWD=2, AZ_GWC_BWD_NT=256, AZ_GWC_BWD_STAGES=5),
                                   dict(AZ_GWC_BWD=2, AZ_GWC_BWD_STAGES=5)])
@pytest.mark.parametrize("shape,dq,G", [((2, 32, 9, 60), 48, 8), ((1, 32, 136, 240), 48, 8), ((1, 16, 7, 36), 13, 16),
                                        ((1, 8, 5, 24), 30, 2), ((1, 64, 3, 16), 8, 8), ((2, 16, 70, 128), 47, 4),
                                        ((1, 8, 3, 1028), 9, 4), ((1, 4, 37, 4), 6, 4), ((1, 8, 11, 520), 50, 2)])
def test_gwc_kernel_variants(knobs, shape, dq, G):
    torch.manual_seed(12)
    L = torch.randn(shape, dtype=torch.float64, requires_grad=True)
    R = torch.randn(shape, dtype=torch.float64, requires_grad=True)
    vol = so.gwc_volume(L, R, dq, G)
    g = torch.randn(vol.shape, dtype=torch.float64)
    vol.backward(g)
    with _env(**knobs):
        Lg = L.detach().float().to(DEV).requires_grad_(True)
        Rg = R.detach().float().to(DEV).requires_grad_(True)
        out = ops.build_gwc_volume(Lg, Rg, dq, G)
        out.backward(g.float().to(DEV))
        torch.cuda.synchronize()
    close(out, vol)
    close(Lg.grad, L.grad)
    close(Rg.grad, R.grad)
    # one-sided gradients take the same kernels
    with _env(**knobs):
        L2 = L.detach().float().to(DEV).requires_grad_(True)
        ops.build_gwc_volume(L2, R.detach().float().to(DEV), dq, G).backward(g.float().to(DEV))
    assert torch.equal(L2.grad, Lg.grad)


# --------------------------------------------------------------------------- kernel variants behind the tuning knobs
@pytest.mark.parametrize("impl", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("shape,dq", [((2, 32, 136, 240), 48), ((1, 4, 33, 64), 12), ((1, 3, 5, 8), 8), ((1, 32, 64, 128), 47),
                                      ((1, 2, 40, 480), 72), ((1, 2, 3, 24), 29), ((1, 2, 3, 12), 5)])
def test_concat_fwd_variants_bit_exact(impl, shape, dq):
    """Register/LSU path, bulk-store path (cp.async.bulk shared->global), the hybrid, and the 256-bit load/store forms
    (3 / 4 / 5: CTAs of 128 / 256 / 64 threads; W % 8 != 0 falls through to the 128-bit kernel): all bit-exact."""
    torch.manual_seed(13)
    L, R = torch.randn(shape), torch.randn(shape)
    ref = so.concat_volume(L, R, dq)
    with _env(AZ_CONCAT_FWD=impl):
        out = ops.build_concat_volume(L.to(DEV), R.to(DEV), dq)
        torch.cuda.synchronize()
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("impl", [0, 1, 2, 3, 4])
def test_soft_argmin_fwd_variants(impl):
    torch.manual_seed(14)
    for shape, scale, off in [((1, 192, 16, 64), 4.0, 0.0), ((1, 192, 6, 20), 1e3, 0.0), ((1, 192, 6, 20), 30.0, 1e6),
                              ((2, 96, 5, 28), 10.0, 0.0), ((1, 50, 3, 8), 2.0, 0.0), ((1, 192, 4, 8), 0.0, 0.0)]:
        cost = torch.randn(shape) * scale + off
        c64 = cost.double()
        ref64 = (torch.softmax(c64, 1) * torch.arange(shape[1], dtype=torch.float64).view(1, -1, 1, 1)).sum(1, keepdim=True)
        with _env(AZ_SA_FWD=impl):
            cg = cost.to(DEV).requires_grad_(True)
            out = ops.soft_argmin(cg)
            out.backward(torch.ones_like(out))
            torch.cuda.synchronize()
        assert float((out.detach().cpu().double() - ref64).abs().max()) <= 2e-5, (impl, shape, scale)
        # the saved (max, log2 sum) must serve the backward whatever kernel produced them
        p = torch.softmax(c64, 1)
        gref = p * (torch.arange(shape[1], dtype=torch.float64).view(1, -1, 1, 1) - ref64)
        close(cg.grad, gref)
    cost = torch.randn(1, 192, 4, 8)
    cost[:, :16] = float("-inf")
    with _env(AZ_SA_FWD=impl):
        out = ops.soft_argmin(cost.to(DEV))
    assert float((out.cpu() - so.soft_argmin(cost)).abs().max()) <= 1e-4


@pytest.mark.parametrize("impl", [0, 1])
def test_upsample_soft_argmin_variants_full_width(impl):
    """The 4-pixel packed kernel (x4 in depth and width) against fp64 and against round 1's kernel, including the
    image borders (x = 0, 1, W-2, W-1 sample one low-res column) and a non-integer vertical factor."""
    torch.manual_seed(15)
    for (Dq, Hq, Wq), (H,), scale in [((48, 10, 40), (40,), 6.0), ((12, 7, 33), (19,), 3.0), ((48, 5, 9), (20,), 1e3)]:
        low = torch.randn(1, 1, Dq, Hq, Wq) * scale
        size = (4 * Dq, H, 4 * Wq)
        up = F.interpolate(low.double(), size, mode="trilinear", align_corners=False).squeeze(1)
        ref64 = (torch.softmax(up, 1) * torch.arange(size[0], dtype=torch.float64).view(1, -1, 1, 1)).sum(1, keepdim=True)
        with _env(AZ_USA_FWD=impl):
            lg = low.to(DEV).requires_grad_(True)
            out = ops.upsample_soft_argmin(lg, size)
            out.backward(torch.ones_like(out))
            torch.cuda.synchronize()
        err = float((out.detach().cpu().double() - ref64).abs().max())
        ref32 = so.soft_argmin(F.interpolate(low, size, mode="trilinear", align_corners=False).squeeze(1))
        theirs = float((ref32.double() - ref64).abs().max())
        # fp32 interpolation noise scales with the logits' magnitude: torch's own fp32 path is the yardstick there
        assert err <= 1e-4 + 2.0 * theirs, (impl, err, theirs)
        assert torch.isfinite(lg.grad).all()


# --------------------------------------------------------------------------- torch.library ops reject bad arguments
def test_library_ops_validate_before_launch():
    a = torch.randn(1, 8, 6, 12, device=DEV)
    with pytest.raises(ValueError):
        library_ops.concat_volume(a, a[:, :, :3], 4, False)          # tgt smaller than ref: would read out of bounds
    with pytest.raises(ValueError):
        library_ops.concat_volume(a, a.cpu(), 4, False)
    img = torch.rand(1, 1, 16, 24, device=DEV)
    d = torch.rand(1, 1, 16, 24, device=DEV)
    m = torch.ones(1, 1, 16, 24, device=DEV, dtype=torch.bool)
    with pytest.raises(ValueError):
        library_ops.reproj_loss(img, img[:, :, :8], d, m)
    with pytest.raises(ValueError):
        library_ops.reproj_loss(img, img, d[:, :, :8], m)
    with pytest.raises(ValueError):
        library_ops.reproj_loss(img, img, d, m[:, :, :8])
    with pytest.raises(ValueError):
        library_ops.reproj_loss(img, img, d, m.cpu())
    with pytest.raises(ValueError):
        ops.patch_fold(img, d[:, :, :8], 5)
    # fractional float masks select like ops.reproj_loss does (!= 0), not by truncation
    mf = torch.full((1, 1, 16, 24), 0.5, device=DEV)
    l1, _ = library_ops.reproj_loss(img, torch.rand_like(img), d, mf)
    l2, _ = ops.reproj_loss(img, torch.rand_like(img), d, mf)
    assert torch.isfinite(l1) and torch.isfinite(l2)
    torch.cuda.synchronize()


def test_patch_loss_with_image_gradient_falls_back_to_autograd():
    """The reference's get_reproj_error_patch is ordinary autograd code ("feature or image"): an input that requires
    grad must still get its gradient (Unfold + the differentiable warp kernel)."""
    torch.manual_seed(16)
    L = torch.rand(1, 1, 24, 40)
    R = torch.rand(1, 1, 24, 40).requires_grad_(True)
    d = torch.rand(1, 1, 24, 40) * 6
    ref, _, _ = so.reproj_error_patch(L, R, d, None, ps=5)
    ref.backward()
    Rg = R.detach().to(DEV).requires_grad_(True)
    loss, vis, _ = az_rp.get_reproj_error_patch(L.to(DEV), Rg, d.to(DEV), None, ps=5)
    loss.backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    close(Rg.grad, R.grad, rtol=1e-4, floor=1e-4)  # float atomics in the image gradient (as torch's grid_sampler)
    assert vis is not None


def test_err_metrics_propagate_nan():
    from activezero_b200.utils.cascade_metrics import compute_err_metric

    g = torch.rand(1, 1, 8, 8, device=DEV) + 1
    pred = g.clone()
    pred[0, 0, 0, 0] = float("nan")
    m = torch.ones(1, 1, 8, 8, device=DEV, dtype=torch.bool)
    f = torch.ones(1, 1, device=DEV)
    err = compute_err_metric(g, g, pred, f, f, m)
    assert np.isnan(err["epe"]) and np.isnan(err["depth_abs_err"])


# --------------------------------------------------------------------------- a7 patch loss + Fold, round-2 kernel (v3)
def _patch_case(B, C, H, W, ps, seed, field, masked=True):
    gen = torch.Generator().manual_seed(seed)
    L = torch.rand(B, C, H, W, generator=gen)
    R = torch.rand(B, C, H, W, generator=gen)
    if field == "random":
        d = torch.rand(B, 1, H, W, generator=gen) * min(64.0, W / 3)
    elif field == "smooth":
        d = (W / 10 + W / 16 * torch.sin(torch.arange(W).float() / 9).view(1, 1, 1, W)
             + 3 * torch.cos(torch.arange(H).float() / 5).view(1, 1, H, 1)).expand(B, 1, H, W).contiguous()
    else:  # samples leaving the image on both sides, boundary cells included
        d = (torch.rand(B, 1, H, W, generator=gen) - 0.5) * 3.0 * W
    m = (torch.rand(B, 1, H, W, generator=gen) > 0.3) if masked else None
    return L, R, d, m


@pytest.mark.parametrize("shape,ps,field", [((1, 1, 40, 64), 11, "random"), ((2, 1, 33, 100), 11, "smooth"),
                                            ((1, 2, 37, 72), 5, "random"), ((1, 1, 24, 90), 3, "wild"),
                                            ((1, 1, 30, 133), 11, "wild"), ((1, 3, 26, 61), 7, "random"),
                                            ((1, 1, 45, 260), 13, "random"), ((1, 1, 64, 256), 9, "smooth"),
                                            ((1, 1, 23, 12), 11, "random"), ((1, 1, 12, 8), 5, "wild"),
                                            ((1, 1, 33, 2610), 11, "smooth"), ((1, 1, 28, 1920), 11, "random")])
def test_patch_loss_fold_v3_vs_oracle(shape, ps, field):
    """The round-2 one-pass kernel (lanes over tap rows) against the oracle's restatement of
    reprojection.py:99-127: loss, d loss / d disp and the Fold image, ragged widths and out-of-image samples.  The
    last two shapes split the row into x-tiles (four at W = 2610, a width round 1's kernels cannot stage at all)."""
    B, C, H, W = shape
    L, R, d, m = _patch_case(B, C, H, W, ps, 40 + H + W, field)
    d64 = d.clone().requires_grad_(True)
    ref, ref_vis, _ = so.reproj_error_patch(L, R, d64, m, ps=ps)
    ref.backward()
    with _env(AZ_PATCH_IMPL=1):
        dg = d.to(DEV).requires_grad_(True)
        loss, vis, _ = az_rp.get_reproj_error_patch(L.to(DEV), R.to(DEV), dg, None if m is None else m.to(DEV), ps=ps)
        loss.backward()
        torch.cuda.synchronize()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item()), (loss.item(), ref.item())
    close(vis, ref_vis)
    close(dg.grad, d64.grad)


@pytest.mark.parametrize("H,W,field", [(544, 960, "random"), (720, 1280, "random"), (256, 512, "smooth"), (97, 1000, "wild"),
                                      (64, 1404, "random"), (70, 1920, "random"), (45, 1920, "wild"), (1088, 1920, "random")])
def test_patch_loss_fold_v3_vs_round1_kernel_large(H, W, field):
    """Sizes the oracle cannot afford in a test: the round-2 kernel against round 1's (itself pinned to the real
    reference by tests/golden/reprojection*.npz), forward with and without the gradient, deterministic.  W = 1920 and
    2610 make the round-2 kernel split the row into x-tiles (seam columns handed over through the side buffer)."""
    L, R, d, m = _patch_case(2, 1, H, W, 11, 7, field)
    Lg, Rg, mg = L.to(DEV), R.to(DEV), m.to(DEV)
    res = {}
    for impl in (0, 1):
        with _env(AZ_PATCH_IMPL=impl):
            dg = d.to(DEV).requires_grad_(True)
            loss, vis = ops.reproj_loss(Lg, Rg, dg, mg, ps=11, want_warped=True)
            loss.backward()
            loss2, vis2 = ops.reproj_loss(Lg, Rg, dg.detach(), mg, ps=11, want_warped=True)
            torch.cuda.synchronize()
        assert torch.equal(vis, vis2) and float(loss.detach()) == float(loss2.detach())  # run-to-run and grad / no-grad instances agree
        res[impl] = (loss.item(), vis, dg.grad)
    assert abs(res[1][0] - res[0][0]) <= 1e-6 * abs(res[0][0])
    close(res[1][1], res[0][1])
    close(res[1][2], res[0][2])


@pytest.mark.parametrize("impl", [0, 1])
def test_patch_golden_fixtures_under_both_kernels(golden, impl):
    """The real-reference fixtures (oracle/make_golden.py executed utils/reprojection.py) pin BOTH one-pass kernels:
    round 1's (AZ_PATCH_IMPL=0, still the fallback for widths the new plan cannot hold) and the round-2 one."""
    import test_gpu_parity as tp

    with _env(AZ_PATCH_IMPL=impl):
        for tag, img, masked, ps in [("p11m", "1", True, 11), ("p11", "1", False, 11)]:
            tp.test_reproj_patch_golden(golden, tag, img, masked, ps)
        for tag, ps in [("w11", 11), ("w7", 7)]:
            tp.test_reproj_patch_wide_golden(golden, tag, ps)
        tp.test_reproj_patch_empty_mask_nan_and_flags(golden)
        torch.cuda.synchronize()


@pytest.mark.parametrize("impl", [0, 1])
def test_patch_real_reference_720x1280(golden, impl):
    """One real-size IR frame (720x1280, datasets/messytable.py:325) against the REAL reference's output
    (oracle/make_golden.py:gen_reprojection_720): loss, strided samples of the Fold image and of d loss / d disp and
    their full-frame sums.  impl = 1 is the round-2 one-pass kernel (W = 1280 fits its single-buffer plan);
    impl = 0 is round 1's path, which at this width is the stand-alone loss + Fold kernel pair."""
    from oracle.make_golden import reprojection_720_inputs

    g = golden("reprojection_720")
    L, R, disp, mask = reprojection_720_inputs()
    with _env(AZ_PATCH_IMPL=impl):
        d = disp.to(DEV).requires_grad_(True)
        loss, vis, mi = az_rp.get_reproj_error_patch(L.to(DEV), R.to(DEV), d, mask.to(DEV), ps=11)
        loss.backward()
        torch.cuda.synchronize()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * float(g["loss"])
    assert int(mi.sum()) == int(g["mask_sum"])
    vmax, gmax = float(g["vis_abs_max"]), float(g["gdisp_abs_max"])
    assert float((vis.cpu()[:, :, ::7, ::5] - torch.from_numpy(g["vis_s"])).abs().max()) <= 1e-5 * vmax
    gs = d.grad.cpu()[:, :, ::7, ::5]
    ref = torch.from_numpy(g["gdisp_s"])
    assert bool(((gs - ref).abs() <= 1e-5 * ref.abs() + 1e-5 * gmax).all())
    assert abs(float(vis.double().sum()) - float(g["vis_sum"])) <= 1e-6 * float(g["vis_sum"])
    assert abs(float(d.grad.double().abs().sum()) - float(g["gdisp_abs_sum"])) <= 1e-5 * float(g["gdisp_abs_sum"])


# --------------------------------------------------------------------------- a6 backward: deterministic image gradient
@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("shape", [(2, 3, 24, 40), (1, 1, 17, 133), (1, 2, 64, 256)])
def test_warp_image_gradient_matches_autograd(impl, shape):
    """d out / d img of apply_disparity (reprojection.py:13-35) against fp64 autograd of the oracle; impl = 1 is the
    round-2 gather kernel (order-independent fixed-point sums: bit-identical from run to run, gimg fully written),
    impl = 0 round 1's float atomics."""
    torch.manual_seed(17)
    B, C, H, W = shape
    img = torch.randn(shape, dtype=torch.float64, requires_grad=True)
    d = (torch.rand(B, 1, H, W, dtype=torch.float64) - 0.3) * 0.6 * W  # both borders are crossed
    d.requires_grad_(True)
    g = torch.randn(shape, dtype=torch.float64)
    so.apply_disparity(img, d).backward(g)
    with _env(AZ_WARP_BWD_IMG=impl):
        grads = []
        for _ in range(2):
            ig = img.detach().float().to(DEV).requires_grad_(True)
            dg = d.detach().float().to(DEV).requires_grad_(True)
            ops.warp(ig, dg).backward(g.float().to(DEV))
            torch.cuda.synchronize()
            grads.append((ig.grad.clone(), dg.grad.clone()))
    close(grads[0][0], img.grad, rtol=1e-5, floor=1e-5)
    close(grads[0][1], d.grad, rtol=1e-4, floor=1e-4)  # fp32 sample positions: the disparity gradient jumps at integer positions
    if impl == 1:
        assert torch.equal(grads[0][0], grads[1][0])


def test_warp_image_gradient_degenerate_upstream():
    img = torch.randn(1, 1, 8, 16, device=DEV, requires_grad=True)
    d = torch.rand(1, 1, 8, 16, device=DEV) * 4
    ops.warp(img, d).backward(torch.zeros(1, 1, 8, 16, device=DEV))
    assert float(img.grad.abs().max()) == 0.0


def test_temporal_ir_integer_form_is_exact_on_ties_free_input():
    """The integer evaluation of tools/temporal_ir.py:93-114 against the oracle's float64 restatement: identical
    patterns (the two can only differ within float64 rounding of the threshold), several T, ks and a constant image."""
    rs = np.random.RandomState(3)
    for T_, H, W, ks in [(7, 90, 130, 11), (4, 64, 64, 11), (12, 40, 57, 5), (2, 33, 20, 3)]:
        fr = rs.randint(0, 256, (T_, H, W)).astype(np.uint8)
        ref = so.temporal_ir_pattern(fr, ks=ks, threshold=0.005)
        out = ops.temporal_ir_pattern(torch.from_numpy(fr).to(DEV), ks=ks, threshold=0.005).cpu().numpy()
        assert float((out != ref).mean()) <= 1e-5, (T_, H, W, ks)
    const = torch.full((7, 32, 32), 9, dtype=torch.uint8, device=DEV)
    assert float(ops.temporal_ir_pattern(const).abs().max()) == 0.0  # the reference divides 0/0 here: NaN > thr is False


# --------------------------------------------------------------------------- §8f-2: implicit-concat first convolution
def _conv_ref(L, R, w, dq, tf32):
    vol = ops.build_concat_volume(L, R, dq)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = tf32
    try:
        return F.conv3d(vol, w, padding=1)
    finally:
        torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("impl", [2, 1])
@pytest.mark.parametrize("shape,dq", [((1, 32, 9, 140), 12), ((2, 32, 6, 128), 8), ((1, 32, 5, 37), 48), ((1, 32, 34, 60), 5),
                                      ((1, 32, 3, 300), 20), ((1, 32, 4, 240), 48), ((1, 32, 1, 126), 3), ((1, 32, 2, 2), 1),
                                      ((2, 32, 3, 127), 55), ((1, 32, 2, 131), 56)])
def test_volume_conv0_matches_conv3d_of_materialised_volume(impl, shape, dq, monkeypatch):
    """conv3d(concat_volume) without the volume (tcgen05 TF32, fp32 accumulation in tensor memory) against stock
    cuDNN on the materialised volume: inside cuDNN's own TF32-vs-fp32 distance, and exact on one-hot probes.
    impl 2 = the shifted-coordinate, feature-stationary kernel (volume_conv_v2.cu; Dq = 56 falls back to impl 1),
    impl 1 = the per-tile gather kernel."""
    monkeypatch.setenv("AZ_VCONV", str(impl))
    torch.manual_seed(50)
    L, R = torch.randn(shape, device=DEV), torch.randn(shape, device=DEV)
    w = torch.randn(32, 64, 3, 3, 3, device=DEV) * 0.05
    out = ops.volume_conv0(L, R, ops.pack_volume_conv_weight(w), dq)
    ref32 = _conv_ref(L, R, w, dq, tf32=False)
    ref_tf = _conv_ref(L, R, w, dq, tf32=True)
    assert out.shape == ref32.shape
    scale = float(ref32.abs().max())
    mine, theirs = float((out - ref32).abs().max()), float((ref_tf - ref32).abs().max())
    assert mine <= max(2.0 * theirs, 2e-3 * scale), (mine, theirs, scale)
    # structure probe: values that TF32 represents exactly (small integers, one-hot weights) must come out exactly --
    # every tap, every channel of both halves, the diagonal mask and all four borders
    Li = torch.randint(-3, 4, shape, device=DEV).float()
    Ri = torch.randint(-3, 4, shape, device=DEV).float()
    wi = torch.randint(-2, 3, (32, 64, 3, 3, 3), device=DEV).float()
    outi = ops.volume_conv0(Li, Ri, ops.pack_volume_conv_weight(wi), dq)
    # exact reference: float64 convolution on the CPU (cuDNN picks FFT algorithms for H <= 2, which are not exact even
    # on integers: 3e-5 off in float32 AND in float64)
    voli = ops.build_concat_volume(Li, Ri, dq).double().cpu()
    assert torch.equal(outi.double().cpu(), F.conv3d(voli, wi.double().cpu(), padding=1))


def test_volume_conv0_epilogue_and_psmnet_flag():
    torch.manual_seed(51)
    L, R = torch.randn(1, 32, 8, 64, device=DEV), torch.randn(1, 32, 8, 64, device=DEV)
    w = torch.randn(32, 64, 3, 3, 3, device=DEV) * 0.05
    sc, sh = torch.rand(32, device=DEV) + 0.5, torch.randn(32, device=DEV)
    out = ops.volume_conv0(L, R, ops.pack_volume_conv_weight(w), 6, sc, sh, relu=True)
    ref = torch.relu(_conv_ref(L, R, w, 6, tf32=False) * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1))
    assert float((out - ref).abs().max()) <= 3e-3 * float(ref.abs().max())
    from activezero_b200.nets.psmnet.psmnet_3 import PSMNet

    net = PSMNet(maxdisp=192).cuda().eval()
    # non-trivial BatchNorm statistics, as a trained checkpoint has
    bn = net.dres0[0][1]
    bn.running_mean.normal_(0, 0.2)
    bn.running_var.uniform_(0.5, 2.0)
    bn.weight.data.uniform_(0.5, 1.5)
    bn.bias.data.normal_(0, 0.3)
    fl, fr = torch.randn(1, 32, 16, 64, device=DEV), torch.randn(1, 32, 16, 64, device=DEV)
    with torch.no_grad():
        want = net.dres0[1](net.dres0[0](ops.build_concat_volume(fl, fr, 48)))
        got = net._first_conv_implicit(fl, fr)
    assert float((got - want).abs().max()) <= 3e-3 * float(want.abs().max())
    # end to end: the flag only changes how dres0's first layer is computed (a random-initialised network's disparities
    # are arg-max plateaus that flip on 1e-3 logit noise, so only shape / finiteness / range are asserted here)
    a, b = torch.rand(1, 3, 256, 256, device=DEV), torch.rand(1, 3, 256, 256, device=DEV)
    with torch.no_grad():
        p_ref = net(a, b)
        net.fuse_volume_conv = True
        p_fused = net(a, b)
    assert p_fused.shape == p_ref.shape and torch.isfinite(p_fused).all()
    assert float(p_fused.min()) >= 0.0 and float(p_fused.max()) <= 191.0


# --------------------------------------------------------------------------- a10's caller: the trainer's GT chain, fused
@pytest.mark.parametrize("shape", [(2, 1, 64, 128), (1, 1, 1088, 1920), (3, 1, 7, 10), (1, 1, 2, 2), (1, 1, 33, 65)])
@pytest.mark.parametrize("sign", [1.0, -1.0])
def test_gt_chain_matches_reference_chain(shape, sign):
    """train.py:255-272: interpolate(nearest, 0.5) -> .type(int) -> apply_disparity_cu(r, r.int()) -> mask, as the
    reference spells it (torch's interpolate, the oracle's scatter warp = the reference kernel semantics) against the
    fused launch: bit-exact disparity, identical mask; zeros (holes in the rendered GT) and values above max_disp in."""
    torch.manual_seed(33)
    d2 = torch.rand(shape) * 230.0 * (torch.rand(shape) > 0.25).float() * sign
    r = F.interpolate(d2, scale_factor=0.5, mode="nearest", recompute_scale_factor=False)
    ref = so.scatter_warp(r, r.type(torch.int))
    ref_mask = (ref < 192) * (ref > 0)
    from activezero_b200.utils import warp_ops as az_wo
    out, mask = az_wo.disp_gt_from_right_view(d2.to(DEV), 192)
    assert out.shape == ref.shape and mask.dtype == torch.bool
    assert torch.equal(out.cpu(), ref)
    assert torch.equal(mask.cpu(), ref_mask)
    # and against the unfused calls of this library
    r_dev = F.interpolate(d2.to(DEV), scale_factor=0.5, mode="nearest", recompute_scale_factor=False)
    assert torch.equal(out, ops.scatter_warp(r_dev, r_dev.type(torch.int)))


# --------------------------------------------------------------------------- empty batch (B = 0)
def test_empty_batch_follows_the_reference():
    """The reference's torch code accepts an empty batch (empty tensors out, NaN for the MSE of nothing -- checked on
    the oracle below); the operators return the same shapes without enqueueing anything, forward and backward, through
    both the autograd Functions and the torch.library ops."""
    from activezero_b200 import _lib

    L0 = torch.zeros(0, 4, 5, 12)
    cost0 = torch.zeros(0, 8, 20, 48)
    img0, disp0 = torch.zeros(0, 1, 20, 48), torch.zeros(0, 1, 20, 48)
    mask0 = torch.zeros(0, 1, 20, 48, dtype=torch.bool)
    Lg, Rg = L0.to(DEV).requires_grad_(True), L0.to(DEV).requires_grad_(True)
    cg = cost0.to(DEV).requires_grad_(True)
    before = _lib.launch_count

    vol = ops.build_concat_volume(Lg, Rg, 3)
    assert vol.shape == so.concat_volume(L0, L0, 3).shape == (0, 8, 3, 5, 12)
    gvol = ops.build_gwc_volume(Lg, Rg, 3, 2)
    assert gvol.shape == so.gwc_volume(L0, L0, 3, 2).shape
    disp = ops.soft_argmin(cg)
    assert disp.shape == so.soft_argmin(cost0).shape == (0, 1, 20, 48)
    low = torch.zeros(0, 1, 2, 5, 12, device=DEV, requires_grad=True)
    assert ops.upsample_soft_argmin(low, (8, 20, 48)).shape == (0, 1, 20, 48)
    assert ops.warp(img0.to(DEV), disp).shape == so.apply_disparity(img0, disp0).shape

    loss, vis, m = az_rp.get_reproj_error_patch(img0.to(DEV), img0.to(DEV), disp, mask0.to(DEV), ps=5)
    rloss, rvis, rm = so.reproj_error_patch(img0, img0, disp0, mask0, ps=5)
    assert torch.isnan(loss) and torch.isnan(rloss)
    assert vis.shape == rvis.shape and m.shape == rm.shape and m.dtype == rm.dtype
    loss1, warped1, _ = az_rp.get_reprojection_error_old(img0.to(DEV), img0.to(DEV), disp, mask0.to(DEV))
    assert torch.isnan(loss1) and warped1.shape == img0.shape

    (vol.sum() + gvol.sum() + disp.sum() + loss).backward()
    assert Lg.grad.shape == L0.shape and Rg.grad.shape == L0.shape and cg.grad.shape == cost0.shape

    out = ops.scatter_warp(img0.to(DEV), torch.zeros(0, 1, 20, 48, dtype=torch.int32, device=DEV))
    assert out.shape == img0.shape
    assert ops.temporal_ir_pattern(torch.zeros(0, 7, 20, 48, dtype=torch.uint8, device=DEV)).shape == (0, 20, 48)
    normed, std = ops.local_contrast_norm(img0.to(DEV))
    assert normed.shape == std.shape == img0.shape
    sums = ops.error_metric_sums(disp0.to(DEV), disp0.to(DEV), disp0.to(DEV), mask0.to(DEV),
                                 depth_pred=disp0.to(DEV))
    assert torch.equal(sums.cpu(), torch.zeros(8, dtype=torch.float64))

    ns = torch.ops.az_stereo
    assert ns.concat_volume(Lg.detach(), Rg.detach(), 3, False).shape == (0, 8, 3, 5, 12)
    d2, _ = ns.soft_argmin(cg.detach())
    assert d2.shape == (0, 1, 20, 48)
    l2 = ns.reproj_loss(img0.to(DEV), img0.to(DEV), d2, mask0.to(DEV), 5, -1.0)[0]
    assert torch.isnan(l2)
    torch.cuda.synchronize()
    assert _lib.launch_count == before, "an empty batch must not enqueue a kernel"

"""torch.library registration of the hot-path operators (activezero_b200/library_ops.py): schema / fake kernel /
autograd registration checks, equality with the autograd.Function layer, and tracing through torch.compile
(aot_eager backend: the graph is captured and the custom ops stay opaque; no code generation involved)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from activezero_b200 import library_ops as lo
    from activezero_b200 import ops

DEV = "cuda:0"


def _inputs():
    torch.manual_seed(0)
    L = torch.randn(2, 8, 6, 20, device=DEV, requires_grad=True)
    R = torch.randn(2, 8, 6, 20, device=DEV, requires_grad=True)
    cost = torch.randn(2, 12, 24, 80, device=DEV, requires_grad=True)
    pL = (torch.rand(2, 1, 24, 80, device=DEV) > 0.5).float()
    pR = (torch.rand(2, 1, 24, 80, device=DEV) > 0.5).float()
    mask = torch.rand(2, 1, 24, 80, device=DEV) > 0.3
    return L, R, cost, pL, pR, mask


def test_ops_are_registered_and_match_the_function_layer():
    L, R, cost, pL, pR, mask = _inputs()
    for cl in (False, True):
        v1 = torch.ops.az_stereo.concat_volume(L, R, 5, cl)
        v2 = ops.build_concat_volume(L, R, 5, channels_last=cl)
        assert torch.equal(v1, v2) and v1.stride() == v2.stride()
    d1 = lo.soft_argmin(cost)
    d2 = ops.soft_argmin(cost)
    assert torch.equal(d1, d2)
    for ps in (1, 7):
        l1, w1 = lo.reproj_loss(pL, pR, d1, mask, ps=ps)
        l2, w2 = ops.reproj_loss(pL, pR, d2, mask, ps=ps, want_warped=True)
        assert torch.equal(l1, l2) and torch.equal(w1, w2)
    # gradients through the registered autograd formulas equal the Function layer's
    g = torch.autograd.grad(lo.reproj_loss(pL, pR, lo.soft_argmin(cost), mask, ps=7)[0]
                            + torch.ops.az_stereo.concat_volume(L, R, 5, True).square().mean(), [cost, L, R])
    h = torch.autograd.grad(ops.reproj_loss(pL, pR, ops.soft_argmin(cost), mask, ps=7)[0]
                            + ops.build_concat_volume(L, R, 5, channels_last=True).square().mean(), [cost, L, R])
    for a, b in zip(g, h):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-9)


def test_opcheck():
    L, R, cost, pL, pR, mask = _inputs()
    checks = ("test_schema", "test_faketensor", "test_autograd_registration")
    torch.library.opcheck(torch.ops.az_stereo.concat_volume.default, (L, R, 5, False), test_utils=checks)
    torch.library.opcheck(torch.ops.az_stereo.concat_volume.default, (L, R, 5, True), test_utils=checks)
    torch.library.opcheck(torch.ops.az_stereo.soft_argmin.default, (cost,), test_utils=checks)
    disp = lo.soft_argmin(cost).detach().requires_grad_(True)
    torch.library.opcheck(torch.ops.az_stereo.reproj_loss.default, (pL, pR, disp, mask, 5, -1.0), test_utils=checks)


def test_traces_through_torch_compile():
    L, R, cost, pL, pR, mask = _inputs()

    def step(L, R, cost):
        vol = torch.ops.az_stereo.concat_volume(L, R, 5, False)
        disp = lo.soft_argmin(cost)
        loss, _ = lo.reproj_loss(pL, pR, disp, mask, ps=5)
        return loss + vol.mean()

    eager = step(L, R, cost)
    compiled = torch.compile(step, backend="aot_eager", fullgraph=True)(L, R, cost)
    assert torch.allclose(eager, compiled, rtol=1e-6)
    ge = torch.autograd.grad(eager, [L, R, cost])
    gc = torch.autograd.grad(compiled, [L, R, cost])
    for a, b in zip(ge, gc):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-9)


def _more_inputs():
    torch.manual_seed(1)
    low = torch.randn(2, 1, 3, 6, 20, device=DEV, requires_grad=True)
    img = torch.rand(2, 3, 24, 80, device=DEV, requires_grad=True)
    disp = (8 * torch.rand(2, 1, 24, 80, device=DEV)).requires_grad_(True)
    frames = torch.randint(0, 256, (2, 7, 24, 80), dtype=torch.uint8, device=DEV)
    idisp = torch.randint(0, 9, (2, 1, 24, 80), dtype=torch.int32, device=DEV)
    return low, img, disp, frames, idisp


def test_remaining_ops_match_the_function_layer():
    """gwc volume, fused upsample + soft-argmin, warp (forward and gradients) and the three non-differentiable
    operators through torch.ops.az_stereo are the same C-ABI calls as activezero_b200.ops: identical bits."""
    L, R, *_ = _inputs()
    low, img, disp, frames, idisp = _more_inputs()
    assert torch.equal(torch.ops.az_stereo.gwc_volume(L, R, 5, 4), ops.build_gwc_volume(L, R, 5, 4))
    assert torch.equal(lo.upsample_soft_argmin(low, (12, 24, 80)), ops.upsample_soft_argmin(low, (12, 24, 80)))
    assert torch.equal(torch.ops.az_stereo.warp(img, disp), ops.warp(img, disp))
    assert torch.equal(lo.scatter_warp(disp.detach(), idisp), ops.scatter_warp(disp.detach(), idisp))
    assert torch.equal(torch.ops.az_stereo.temporal_ir_pattern(frames, 11, 0.005), ops.temporal_ir_pattern(frames))
    for a, b in zip(torch.ops.az_stereo.local_contrast_norm(img.detach(), 9, 1e-5), ops.local_contrast_norm(img.detach())):
        assert torch.equal(a, b)

    def total(ns):
        gwc, usa, wp = ((torch.ops.az_stereo.gwc_volume, lo.upsample_soft_argmin, torch.ops.az_stereo.warp) if ns
                        else (ops.build_gwc_volume, ops.upsample_soft_argmin, ops.warp))
        d = usa(low, (12, 24, 80))
        return gwc(L, R, 5, 4).square().sum() + (wp(img, disp + d) * img.detach()).sum()

    g = torch.autograd.grad(total(True), [L, R, low, img, disp])
    h = torch.autograd.grad(total(False), [L, R, low, img, disp])
    for a, b in zip(g, h):
        assert torch.equal(a, b)  # deterministic kernels on both routes
    with pytest.raises(ValueError):
        torch.ops.az_stereo.gwc_volume(L, R[:, :4], 5, 4)
    with pytest.raises(ValueError):
        torch.ops.az_stereo.warp(img, disp[:, :, :-1])
    with pytest.raises(AssertionError):
        lo.scatter_warp(disp.detach(), idisp - 4)  # mixed signs, as utils/warp_ops.py:73-77 asserts


def test_opcheck_remaining_ops():
    L, R, *_ = _inputs()
    low, img, disp, frames, idisp = _more_inputs()
    checks = ("test_schema", "test_faketensor", "test_autograd_registration")
    ns = torch.ops.az_stereo
    torch.library.opcheck(ns.gwc_volume.default, (L, R, 5, 4), test_utils=checks)
    torch.library.opcheck(ns.upsample_soft_argmin.default, (low, 12, 24, 80), test_utils=checks)
    torch.library.opcheck(ns.warp.default, (img, disp), test_utils=checks)
    torch.library.opcheck(ns.scatter_warp.default, (disp.detach(), idisp), test_utils=checks)
    torch.library.opcheck(ns.temporal_ir_pattern.default, (frames, 11, 0.005), test_utils=checks)
    torch.library.opcheck(ns.local_contrast_norm.default, (img.detach(), 9, 1e-5), test_utils=checks)


def test_remaining_ops_trace_through_torch_compile():
    L, R, *_ = _inputs()
    low, img, disp, frames, idisp = _more_inputs()

    def step(L, R, low, img):
        d = lo.upsample_soft_argmin(low, (12, 24, 80))
        pat = torch.ops.az_stereo.temporal_ir_pattern(frames, 11, 0.005).unsqueeze(1)
        w = torch.ops.az_stereo.warp(img, d)
        return (w * pat).mean() + torch.ops.az_stereo.gwc_volume(L, R, 5, 4).mean()

    eager = step(L, R, low, img)
    compiled = torch.compile(step, backend="aot_eager", fullgraph=True)(L, R, low, img)
    assert torch.allclose(eager, compiled, rtol=1e-6)
    for a, b in zip(torch.autograd.grad(eager, [L, R, low, img]), torch.autograd.grad(compiled, [L, R, low, img])):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-9)

"""Out-of-bounds WRITE detection without compute-sanitizer (closed on this pool): every output and workspace buffer of
the entry points below is a window inside a larger allocation whose margins hold a canary pattern; after the call the
margins must be untouched.  Shapes are ragged on purpose (rows that do not fill a CTA, widths that are not multiples of
the vector or tile sizes, tiles that overhang the image).  Reads cannot be checked this way; they are covered by the
bit-exact comparisons at the same shapes elsewhere in the suite."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from activezero_b200 import _lib, ops
    from activezero_b200.ops import _ptr, _stream

DEV = "cuda:0"
GUARD = 4096  # bytes on each side (keeps the window 16-byte aligned)


class Guarded:
    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(shape), dtype
        n = 1
        for s in self.shape:
            n *= s
        self.nbytes = n * torch.empty((), dtype=dtype).element_size()
        pad = (-self.nbytes) % 16
        self.raw = torch.full((GUARD + self.nbytes + pad + GUARD,), 0xA5, dtype=torch.uint8, device=DEV)
        self.t = self.raw[GUARD:GUARD + self.nbytes].view(dtype).view(self.shape)

    def check(self, what):
        lo = self.raw[:GUARD]
        hi = self.raw[GUARD + self.nbytes:]
        assert bool((lo == 0xA5).all()), f"{what}: bytes BEFORE the buffer were written"
        assert bool((hi == 0xA5).all()), f"{what}: bytes AFTER the buffer were written"


@pytest.mark.parametrize("shape,dq,G", [((1, 32, 5, 240), 48, 8), ((2, 8, 3, 36), 13, 4), ((1, 4, 7, 132), 9, 4), ((1, 16, 2, 520), 50, 4),
                                        ((1, 64, 3, 16), 8, 8), ((1, 8, 3, 1028), 9, 4), ((1, 6, 4, 9), 5, 3)])
def test_gwc_backward_writes_stay_inside(shape, dq, G):
    torch.manual_seed(41)
    B, C, H, W = shape
    L, R = torch.randn(shape, device=DEV), torch.randn(shape, device=DEV)
    g = torch.randn(B, G, dq, H, W, device=DEV)
    gL, gR = Guarded(shape, torch.float32), Guarded(shape, torch.float32)
    _lib.call("az_gwc_volume_bwd", _ptr(g), _ptr(L), _ptr(R), _ptr(gL.t), _ptr(gR.t), B, C, H, W, dq, G, _stream())
    torch.cuda.synchronize()
    gL.check("gL")
    gR.check("gR")
    vol = Guarded((B, G, dq, H, W), torch.float32)
    _lib.call("az_gwc_volume_fwd", _ptr(L), _ptr(R), _ptr(vol.t), B, C, H, W, dq, G, _stream())
    torch.cuda.synchronize()
    vol.check("gwc volume")
    # and the results are the ones the autograd path returns
    Lg, Rg = L.clone().requires_grad_(True), R.clone().requires_grad_(True)
    ops.build_gwc_volume(Lg, Rg, dq, G).backward(g)
    assert torch.equal(gL.t, Lg.grad) and torch.equal(gR.t, Rg.grad)


@pytest.mark.parametrize("shape", [(2, 14, 20), (1, 66, 130), (3, 2, 2), (1, 545, 962), (1, 33, 4100)])
def test_gt_chain_and_scatter_warp_writes_stay_inside(shape):
    torch.manual_seed(42)
    N, H2, W2 = shape
    H, W = H2 // 2, W2 // 2
    d2 = torch.rand(N, 1, H2, W2, device=DEV) * 90
    out, mask = Guarded((N, 1, H, W), torch.float32), Guarded((N, 1, H, W), torch.uint8)
    _lib.call("az_scatter_warp_gt", _ptr(d2), _ptr(out.t), _ptr(mask.t), ctypes.c_void_p(0), ctypes.c_float(192.0), N, H2, W2, _stream())
    torch.cuda.synchronize()
    out.check("disp_gt_l")
    mask.check("mask")
    img = torch.rand(N, 2, H, W, device=DEV)
    di = (torch.rand(N, 1, H, W, device=DEV) * 70).int()
    dst = Guarded((N, 2, H, W), torch.float32)
    _lib.call("az_scatter_warp", _ptr(img), _ptr(di), _ptr(dst.t), ctypes.c_void_p(0), N, 2, H, W, _stream())
    torch.cuda.synchronize()
    dst.check("scatter warp dst")
    assert torch.equal(dst.t, ops.scatter_warp(img, di, check_sign=False))


@pytest.mark.parametrize("shape,T_,ks", [((1, 37, 64), 7, 11), ((2, 33, 196), 4, 11), ((1, 5, 8), 3, 11), ((1, 70, 130), 7, 5),
                                         ((1, 3, 4), 2, 21), ((1, 100, 1000), 7, 11)])
def test_temporal_ir_and_lcn_writes_stay_inside(shape, T_, ks):
    torch.manual_seed(43)
    B, H, W = shape
    fr = torch.randint(0, 256, (B, T_, H, W), dtype=torch.uint8, device=DEV)
    pat = Guarded((B, H, W), torch.float32)
    ws = Guarded((_lib.query("az_temporal_ir_workspace_bytes", B, H, W),), torch.uint8)
    _lib.call("az_temporal_ir", _ptr(fr), _ptr(pat.t), _ptr(ws.t), B, T_, H, W, ks, ctypes.c_double(0.005), _stream())
    torch.cuda.synchronize()
    pat.check("pattern")
    ws.check("temporal IR workspace")
    assert torch.equal(pat.t, ops.temporal_ir_pattern(fr, ks=ks, threshold=0.005))
    img = torch.rand(B, 2, H, W, device=DEV)
    for lks in (9, 11, 3):
        normed, std = Guarded((B, 1, H, W), torch.float32), Guarded((B, 1, H, W), torch.float32)
        _lib.call("az_local_contrast_norm", _ptr(img), _ptr(normed.t), _ptr(std.t), B, 2, H, W, lks, ctypes.c_float(1e-5), _stream())
        torch.cuda.synchronize()
        normed.check(f"lcn normed ks={lks}")
        std.check(f"lcn std ks={lks}")


@pytest.mark.parametrize("shape", [(1, 1, 9, 40), (2, 3, 5, 124), (1, 1, 33, 964)])
def test_ps1_loss_writes_stay_inside(shape):
    torch.manual_seed(44)
    B, C, H, W = shape
    tgt, src = torch.rand(shape, device=DEV), torch.rand(shape, device=DEV)
    disp = (torch.rand(B, 1, H, W, device=DEV) * 1.4 - 0.2) * W
    mask = (torch.rand(B, 1, H, W, device=DEV) > 0.3).to(torch.uint8)
    lin_x, lin_y = ops.linspace_table(W, tgt.device), ops.linspace_table(H, tgt.device)
    warped, gpre = Guarded(shape, torch.float32), Guarded((B, 1, H, W), torch.float32)
    loss, stats = Guarded((1,), torch.float32), Guarded((2,), torch.float64)
    ws = Guarded((_lib.query("az_reproj_workspace_bytes", B, H),), torch.uint8)
    _lib.call("az_reproj_loss_fwd", _ptr(tgt), _ptr(src), _ptr(disp), ctypes.c_float(-1.0), _ptr(mask), _ptr(lin_x), _ptr(lin_y), 1,
              _ptr(warped.t), _ptr(gpre.t), _ptr(loss.t), _ptr(stats.t), _ptr(ws.t), B, C, H, W, _stream())
    torch.cuda.synchronize()
    for gbuf, name in ((warped, "warped"), (gpre, "gpre"), (loss, "loss"), (stats, "stats"), (ws, "workspace")):
        gbuf.check(name)
    ref, _ = ops.reproj_loss(tgt, src, disp, mask.bool(), ps=1, want_warped=True)
    assert float(loss.t[0]) == float(ref)


# --------------------------------------------------------------------------- the step kernels and the rest of the library
@pytest.mark.parametrize("shape,dq", [((1, 3, 5, 12), 7), ((2, 32, 6, 60), 48), ((1, 2, 3, 135), 48), ((1, 4, 9, 1028), 12), ((1, 1, 1, 4), 1)])
def test_concat_volume_writes_stay_inside(shape, dq):
    torch.manual_seed(45)
    B, C, H, W = shape
    L, R = torch.randn(shape, device=DEV), torch.randn(shape, device=DEV)
    for entry in ("az_concat_volume_fwd",) + (("az_concat_volume_fwd_ndhwc",) if C % 4 == 0 else ()):
        vol = Guarded((B, 2 * C, dq, H, W), torch.float32)
        _lib.call(entry, _ptr(L), _ptr(R), _ptr(vol.t), B, C, H, W, dq, _stream())
        torch.cuda.synchronize()
        vol.check(entry)
    g = torch.randn(B, 2 * C, dq, H, W, device=DEV)
    gL, gR = Guarded(shape, torch.float32), Guarded(shape, torch.float32)
    _lib.call("az_concat_volume_bwd", _ptr(g), _ptr(gL.t), _ptr(gR.t), B, C, H, W, dq, _stream())
    torch.cuda.synchronize()
    gL.check("concat gL")
    gR.check("concat gR")


@pytest.mark.parametrize("shape", [(1, 7, 3, 5), (2, 48, 9, 36), (1, 192, 5, 130), (1, 33, 17, 1026)])
def test_soft_argmin_writes_stay_inside(shape):
    torch.manual_seed(46)
    B, D, H, W = shape
    cost = torch.randn(shape, device=DEV) * 3
    disp, lse = Guarded((B, 1, H, W), torch.float32), Guarded((B, 2, H, W), torch.float32)
    _lib.call("az_soft_argmin_fwd", _ptr(cost), _ptr(disp.t), _ptr(lse.t), B, D, H, W, _stream())
    torch.cuda.synchronize()
    disp.check("disp")
    lse.check("lse")
    gcost = Guarded(shape, torch.float32)
    gd = torch.randn(B, 1, H, W, device=DEV)
    _lib.call("az_soft_argmin_bwd", _ptr(cost), _ptr(disp.t), _ptr(lse.t), _ptr(gd), _ptr(gcost.t), B, D, H, W, _stream())
    torch.cuda.synchronize()
    gcost.check("gcost")
    assert torch.equal(disp.t, ops.soft_argmin(cost))


@pytest.mark.parametrize("low,out", [((1, 5, 4, 6), (20, 16, 24)), ((2, 12, 9, 15), (48, 36, 60)), ((1, 3, 2, 3), (9, 7, 10))])
def test_upsample_soft_argmin_writes_stay_inside(low, out):
    torch.manual_seed(47)
    B, Dq, Hq, Wq = low
    D, H, W = out
    lowres = torch.randn(B, 1, Dq, Hq, Wq, device=DEV) * 3
    disp, stats = Guarded((B, 1, H, W), torch.float32), Guarded((B, 2, H, W), torch.float32)
    _lib.call("az_upsample_soft_argmin_fwd", _ptr(lowres), _ptr(disp.t), _ptr(stats.t), B, Dq, Hq, Wq, D, H, W, _stream())
    torch.cuda.synchronize()
    disp.check("fused disp")
    stats.check("fused stats")
    glow = Guarded((B, 1, Dq, Hq, Wq), torch.float32)
    ws = Guarded((max(16, _lib.query("az_upsample_soft_argmin_workspace_bytes", B, Dq, H, W, Wq)),), torch.uint8)
    gd = torch.randn(B, 1, H, W, device=DEV)
    _lib.call("az_upsample_soft_argmin_bwd", _ptr(lowres), _ptr(disp.t), _ptr(stats.t), _ptr(gd), _ptr(glow.t), _ptr(ws.t),
              B, Dq, Hq, Wq, D, H, W, _stream())
    torch.cuda.synchronize()
    glow.check("glow")
    ws.check("fused backward workspace")


@pytest.mark.parametrize("shape,ps", [((1, 1, 13, 40), 11), ((2, 1, 30, 124), 11), ((1, 1, 64, 1412), 11), ((1, 2, 9, 36), 5), ((1, 1, 25, 70), 3)])
def test_patch_loss_and_warp_writes_stay_inside(shape, ps):
    torch.manual_seed(48)
    B, C, H, W = shape
    tgt, src = torch.rand(shape, device=DEV), torch.rand(shape, device=DEV)
    disp = (torch.rand(B, 1, H, W, device=DEV) * 1.3 - 0.15) * min(W, 100)
    mask = (torch.rand(B, 1, H, W, device=DEV) > 0.3).to(torch.uint8)
    lin_x, lin_y = ops.linspace_table(W, tgt.device), ops.linspace_table(H, tgt.device)
    vis, gpre = Guarded(shape, torch.float32), Guarded((B, 1, H, W), torch.float32)
    vis.t.zero_()  # the Fold image is accumulated onto zero-filled memory (header contract)
    loss, stats = Guarded((1,), torch.float32), Guarded((2,), torch.float64)
    ws = Guarded((_lib.query("az_reproj_workspace_bytes", B, H),), torch.uint8)
    _lib.call("az_reproj_loss_fwd", _ptr(tgt), _ptr(src), _ptr(disp), ctypes.c_float(-1.0), _ptr(mask), _ptr(lin_x), _ptr(lin_y), ps,
              _ptr(vis.t), _ptr(gpre.t), _ptr(loss.t), _ptr(stats.t), _ptr(ws.t), B, C, H, W, _stream())
    torch.cuda.synchronize()
    for gbuf, name in ((vis, "fold image"), (gpre, "gpre"), (loss, "loss"), (stats, "stats"), (ws, "workspace")):
        gbuf.check(name)
    out = Guarded(shape, torch.float32)
    _lib.call("az_warp_fwd", _ptr(src), _ptr(disp), _ptr(lin_x), _ptr(lin_y), _ptr(out.t), B, C, H, W, _stream())
    gimg, gdisp = Guarded(shape, torch.float32), Guarded((B, 1, H, W), torch.float32)
    gout = torch.randn(shape, device=DEV)
    _lib.call("az_warp_bwd", _ptr(src), _ptr(disp), _ptr(lin_x), _ptr(lin_y), _ptr(gout), _ptr(gimg.t), _ptr(gdisp.t), B, C, H, W, _stream())
    torch.cuda.synchronize()
    out.check("warped")
    gimg.check("warp gimg")
    gdisp.check("warp gdisp")


@pytest.mark.parametrize("shape,dq", [((1, 32, 3, 130), 7), ((2, 32, 2, 37), 48), ((1, 32, 5, 254), 3), ((1, 32, 2, 60), 56)])
def test_volume_conv0_writes_stay_inside(shape, dq):
    """Both tcgen05 kernels (dq = 56 takes the first form): output and packed-weight buffers between canaries."""
    torch.manual_seed(49)
    B, C, H, W = shape
    L, R = torch.randn(shape, device=DEV), torch.randn(shape, device=DEV)
    w = torch.randn(32, 64, 3, 3, 3, device=DEV) * 0.05
    wp = Guarded((27 * 3072,), torch.float32)
    _lib.call("az_volume_conv0_pack", _ptr(w), _ptr(wp.t), _stream())
    out = Guarded((B, 32, dq, H, W), torch.float32)
    _lib.call("az_volume_conv0_fwd", _ptr(L), _ptr(R), _ptr(wp.t), ctypes.c_void_p(0), ctypes.c_void_p(0), _ptr(out.t), B, C, H, W, dq, 0,
              _stream())
    torch.cuda.synchronize()
    wp.check("packed weights")
    out.check("conv output")
    assert torch.equal(out.t, ops.volume_conv0(L, R, wp.t, dq))

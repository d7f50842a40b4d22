"""CPU restatement of the ActiveZero stereo hot path (TEST INFRASTRUCTURE ONLY).

Every function cites the reference lines (relative to ``/root/reference``) it
restates.  The arithmetic that lives in third-party code (``F.grid_sample``,
``F.softmax``, ``nn.Unfold``/``nn.Fold``, ``F.mse_loss``, ``F.interpolate`` from
PyTorch -- reference pins torch 1.10.0, this image has 2.11.0 with the same
defaults -- and ``cv2.blur`` from OpenCV, pinned 4.5.5.62, image has 4.13) is
called, not re-derived, so the oracle is "container torch/cv2 executing the
reference's algorithm".  Device-agnostic: the reference's hard-coded ``.cuda()``
calls are the only thing dropped.

Pinning: ``oracle/make_golden.py`` runs the *real* reference modules on seeded
inputs and stores input/output pairs in ``tests/golden``;
``tests/test_oracle_golden.py`` checks every function below against them.
``gwc_volume`` has no reference counterpart: parity unpinned.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# a1 / a2  concat cost volume           nets/psmnet/psmnet.py:151-165
# ----------------------------------------------------------------------------

def concat_volume(ref_feat: torch.Tensor, tgt_feat: torch.Tensor, num_disp: int) -> torch.Tensor:
    """[B,C,H,W] x2 -> [B,2C,num_disp,H,W].

    psmnet.py:152-156 zero-fills the volume; :158-164 copies, per disparity
    plane ``i``, the left features into channels ``[0,C)`` and the right
    features shifted right by ``i`` into channels ``[C,2C)``, only for
    columns ``x >= i``.  (= psmnet_3.py:149-163.)
    """
    B, C, H, W = ref_feat.shape
    vol = torch.zeros(B, 2 * C, num_disp, H, W, dtype=ref_feat.dtype, device=ref_feat.device)
    for i in range(num_disp):
        if i == 0:
            vol[:, :C, 0] = ref_feat
            vol[:, C:, 0] = tgt_feat
        elif i < W:
            vol[:, :C, i, :, i:] = ref_feat[..., i:]
            vol[:, C:, i, :, i:] = tgt_feat[..., : W - i]
    return vol.contiguous()


# ----------------------------------------------------------------------------
# a3  group-wise correlation volume -- NOT IN THE REFERENCE (parity unpinned)
# ----------------------------------------------------------------------------

def gwc_volume(ref_feat: torch.Tensor, tgt_feat: torch.Tensor, num_disp: int, num_groups: int) -> torch.Tensor:
    """GwcNet-style volume, the definition SURVEY.md §8a row a3 adopts:
    ``vol[b,g,i,y,x] = mean_{c in group g} L[b,c,y,x] * R[b,c,y,x-i]`` for
    ``x >= i`` else 0.  No reference implementation exists.
    """
    B, C, H, W = ref_feat.shape
    assert C % num_groups == 0
    cpg = C // num_groups
    vol = torch.zeros(B, num_groups, num_disp, H, W, dtype=ref_feat.dtype, device=ref_feat.device)
    for i in range(min(num_disp, W)):
        prod = ref_feat[..., i:] * tgt_feat[..., : W - i]
        vol[:, :, i, :, i:] = prod.view(B, num_groups, cpg, H, W - i).mean(dim=2)
    return vol


# ----------------------------------------------------------------------------
# a4 / a5  soft-argmin     psmnet.py:200-201 + psmnet_submodule.py:80-89
# ----------------------------------------------------------------------------

def disparity_regression(prob: torch.Tensor) -> torch.Tensor:
    """psmnet_submodule.py:83-89: ``sum_d prob[:,d] * d`` with ``d`` a float32
    ``range(maxdisp)`` reshaped ``[1,D,1,1]`` (numpy int range -> torch.Tensor)."""
    D = prob.shape[1]
    idx = torch.tensor(np.arange(D), dtype=torch.float32).view(1, D, 1, 1).to(prob.device).type_as(prob)
    return torch.sum(prob * idx, 1, keepdim=True)


def soft_argmin(cost: torch.Tensor) -> torch.Tensor:
    """[B,D,H,W] logits -> [B,1,H,W].  psmnet.py:200-201 (softmax over dim 1,
    then DisparityRegression)."""
    return disparity_regression(F.softmax(cost, dim=1))


# ----------------------------------------------------------------------------
# a6  bilinear disparity warp          utils/reprojection.py:13-35
# ----------------------------------------------------------------------------

def apply_disparity(img: torch.Tensor, disp: torch.Tensor) -> torch.Tensor:
    """Pull ``img`` along x by ``disp`` with ``F.grid_sample``.

    reprojection.py:15 scales disp by 1/W; :18-24 build the base grid from
    ``torch.linspace(0,1,n)`` (fp32, then ``type_as(img)``); :27-28 add the
    shift to x only; :31-33 sample at ``2*grid-1`` with bilinear / zeros and the
    default ``align_corners=False``.  The op order is kept because it decides
    the fp32 rounding of the sample position (SURVEY.md §8a closed forms).
    """
    B, _, H, W = img.shape
    shift = disp / W
    lin_x = torch.linspace(0, 1, W).type_as(img)
    lin_y = torch.linspace(0, 1, H).type_as(img)
    gx = lin_x.view(1, 1, W).expand(B, H, W) + shift[:, 0]
    gy = lin_y.view(1, H, 1).expand(B, H, W)
    grid = torch.stack((gx, gy), dim=3)
    return F.grid_sample(img, 2 * grid - 1, mode="bilinear", padding_mode="zeros", align_corners=False)


# ----------------------------------------------------------------------------
# a10  integer scatter warp            utils/warp_ops.py:20-47, 55-95
# ----------------------------------------------------------------------------

def scatter_warp(img: torch.Tensor, disp: torch.Tensor) -> torch.Tensor:
    """CPU execution of the two NVRTC kernels' row loops.

    warp_ops.py:24-32 (all disp >= 0): ``for j = w-1..0: idx = j+disp[j];
    if idx < w: dst[idx] = src[j]``; :36-44 (all disp <= 0): ``for j = 0..w-1:
    if idx > -1``.  One kernel thread owns one (n,c,y) row; the disparity is
    ``[N,1,H,W]`` (or ``[N,H,W]``) int32 shared by all channels
    (``dbase=(i/h/c*h+i%h)*w``, :27).  Holes stay 0 (:83).  The host asserts of
    :69-77 are kept.  Rows are vectorised with numpy; the column loop is
    sequential, in the kernels' order.
    """
    assert img.is_contiguous() and disp.is_contiguous()
    assert disp.dtype == torch.int32
    src = img.detach().cpu().numpy()
    N, C, H, W = src.shape
    d = disp.detach().cpu().numpy().reshape(N, 1, H, W)
    if (d >= 0).all():
        order = range(W - 1, -1, -1)
        positive = True
    else:
        assert (d <= 0).all()
        order = range(W)
        positive = False
    d = np.broadcast_to(d, (N, C, H, W)).reshape(-1, W)
    s = src.reshape(-1, W)
    out = np.zeros_like(s)
    rows = np.arange(s.shape[0])
    for j in order:
        idx = j + d[:, j]
        ok = (idx < W) if positive else (idx > -1)
        out[rows[ok], idx[ok]] = s[ok, j]
    return torch.from_numpy(out.reshape(N, C, H, W)).to(img.device)


# ----------------------------------------------------------------------------
# a8  single-scale / bidirectional masked MSE   reprojection.py:38-96
# ----------------------------------------------------------------------------

def _masked_mse(a: torch.Tensor, b: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    # reprojection.py:69-70, 95, 118: F.mse_loss over the boolean-selected elements
    return F.mse_loss(a[mask], b[mask])


def reproj_error_old(input_L, input_R, pred_disp_l, mask=None):
    """reprojection.py:81-96 -> (loss, warped [B,C,H,W], mask int [B,C,H,W])."""
    warped = apply_disparity(input_R, -pred_disp_l)
    if mask is not None:
        mask = mask.repeat(1, input_L.shape[1], 1, 1)
    else:
        mask = torch.ones_like(warped).type(torch.bool)
    return _masked_mse(warped, input_L, mask), warped, mask.type(torch.int)


def reproj_error_bidir(input_L, input_R, pred_disp_l, pred_disp_r, mask_l=None, mask_r=None):
    """reprojection.py:38-78.  When no masks are given they are built from the
    integer scatter warp of the truncated disparities (:50-65)."""
    L_w = apply_disparity(input_R, -pred_disp_l)
    R_w = apply_disparity(input_L, pred_disp_r)
    if mask_l is None:
        gt_l = scatter_warp(pred_disp_r.contiguous(), pred_disp_r.type(torch.int).contiguous())
        gt_r = scatter_warp(pred_disp_l.contiguous(), (-pred_disp_l.type(torch.int)).contiguous())
        mask_l = ((gt_l < 192) * (gt_l > 0)).detach()
        mask_r = ((gt_r < 192) * (gt_r > 0)).detach()
    c = input_L.shape[1]
    mask_l = mask_l.repeat(1, c, 1, 1)
    mask_r = mask_r.repeat(1, c, 1, 1)
    return (
        _masked_mse(L_w, input_L, mask_l),
        _masked_mse(R_w, input_R, mask_r),
        L_w,
        R_w,
        mask_l.type(torch.int),
        mask_r.type(torch.int),
    )


# ----------------------------------------------------------------------------
# a7  patch reprojection loss          reprojection.py:99-127
# ----------------------------------------------------------------------------

def reproj_error_patch(input_L, input_R, pred_disp_l, mask=None, ps=5):
    """reprojection.py:99-127 -> (loss, warped [B,C,H,W], mask int [B,C,H,W]).

    :102-111 zero-padded im2col of both images to ``[B, C*ps*ps, H, W]``;
    :112 warp the unfolded right image by ``-disp``; :113-118 masked MSE over
    all ``C*ps*ps`` planes; :120-125 Fold (overlap-sum) the warped planes back
    and crop the ``(ps-1)/2`` border -- visualisation only.
    """
    assert ps % 2 == 1
    B, C, H, W = input_L.shape
    half = (ps - 1) // 2
    K = C * ps * ps
    Lu = F.unfold(input_L, kernel_size=(ps, ps), padding=half).reshape(B, K, H, W)
    Ru = F.unfold(input_R, kernel_size=(ps, ps), padding=half).reshape(B, K, H, W)
    Wu = apply_disparity(Ru, -pred_disp_l)
    if mask is not None:
        mask = mask.repeat(1, K, 1, 1)
    else:
        mask = torch.ones_like(Wu).type(torch.bool)
    loss = _masked_mse(Wu, Lu, mask)
    vis = F.fold(Wu.reshape(B, K, H * W), output_size=(H + ps - 1, W + ps - 1), kernel_size=(ps, ps))
    if ps > 1:
        vis = vis[:, :, half:-half, half:-half]
    return loss, vis, mask[:, :C].type(torch.int)


# ----------------------------------------------------------------------------
# a9  multi-scale loss                 reprojection.py:130-173
# ----------------------------------------------------------------------------

def reproj_error_diff_ratio(input_L, input_R, pred_disp_l, mask=None):
    """reprojection.py:130-173: scales 0.25/0.5/1 (:138), weights 0.3/0.5/0.2
    (:139); bilinear ``F.interpolate`` of images, disparity (times the ratio,
    :155-157) and the float mask cast back to bool (:158)."""
    ratios, weights = (0.25, 0.5, 1), (0.3, 0.5, 0.2)
    if mask is not None:
        mask = mask.repeat(1, input_L.shape[1], 1, 1)
    else:
        mask = torch.ones_like(input_L)
    mask = mask.type(torch.float32).detach()
    total, stages, losses = 0, {}, {}
    for k, (r, wt) in enumerate(zip(ratios, weights)):
        L_s = F.interpolate(input_L, scale_factor=r, mode="bilinear")
        R_s = F.interpolate(input_R, scale_factor=r, mode="bilinear")
        d_s = F.interpolate(pred_disp_l, scale_factor=r, mode="bilinear") * r
        m_s = F.interpolate(mask, scale_factor=r, mode="bilinear").type(torch.bool)
        w_s = apply_disparity(R_s, -d_s)
        l_s = _masked_mse(w_s, L_s, m_s)
        stages[f"stage{k}"] = {"target": L_s, "warped": w_s, "pred_disp": d_s, "mask": m_s.type(torch.int)}
        losses[f"stage{k}"] = l_s.item()
        total = total + l_s * wt
    return total, stages, losses


# ----------------------------------------------------------------------------
# a12  local contrast normalisation    reprojection.py:175-200
# ----------------------------------------------------------------------------

def local_contrast_norm(image, kernel_size=9, eps=1e-5):
    """reprojection.py:175-200: first channel only (:184-185); zero-padded
    ks x ks window mean (:190-192) and population std (:193-197);
    ``(img-mean)/(std+eps)`` (:199).  Returns (normed, std)."""
    assert kernel_size % 2 == 1
    image = image[:, :1]
    B, _, H, W = image.shape
    cols = F.unfold(image, kernel_size, padding=(kernel_size - 1) // 2)
    avg = cols.mean(dim=1).view(B, 1, H, W)
    std = cols.std(dim=1, unbiased=False).view(B, 1, H, W)
    return (image - avg) / (std + eps), std


# ----------------------------------------------------------------------------
# a11  temporal IR pattern             tools/temporal_ir.py:35-40, 93-114
# ----------------------------------------------------------------------------

def box_blur_reflect101(img: np.ndarray, ks: int) -> np.ndarray:
    """Mean over a ks x ks window with BORDER_REFLECT_101 padding -- what
    ``cv2.blur(img, (ks, ks))`` (temporal_ir.py:37) computes with its default
    border.  Used when cv2 is unavailable and to cross-check cv2."""
    h = ks // 2
    pad = np.pad(img, h, mode="reflect")
    acc = np.zeros_like(img, dtype=np.float64)
    H, W = img.shape
    for dy in range(ks):
        for dx in range(ks):
            acc += pad[dy : dy + H, dx : dx + W]
    return acc / (ks * ks)


def temporal_ir_pattern(frames: np.ndarray, ks: int = 11, threshold: float = 0.005, use_cv2: bool = True) -> np.ndarray:
    """frames [T,H,W] uint8 -> pattern [H,W] float64 in {0,1}.

    temporal_ir.py:93-107: per-pixel least-squares line over t = 0..T-1 (the
    reference stacks frames on the last axis, :78-89; T = 7 there);
    :110-111 ``|fit[T-1]-fit[0]|/255``; :113 min-max normalise; :114 + :35-40
    ``|diff| - blur(|diff|) > 0.005``.  numpy promotes uint8 to float64.
    """
    y = np.moveaxis(np.asarray(frames), 0, -1)  # [H,W,T] like the reference
    H, W, T = y.shape
    t = np.linspace(0, T - 1, num=T, dtype=int).reshape(1, 1, -1)
    t = np.repeat(np.repeat(t, H, axis=0), W, axis=1)
    t_avg = np.average(t, axis=-1).reshape(H, W, 1)
    y_avg = np.average(y, axis=-1).reshape(H, W, 1)
    slope = (np.sum((y - y_avg) * (t - t_avg), axis=-1) / np.sum((t - t_avg) ** 2, axis=-1))[:, :, None]
    fit = slope * t + (y_avg - slope * t_avg)
    diff = np.abs((fit[:, :, -1] - fit[:, :, 0]) / 255)
    diff = (diff - np.min(diff)) / (np.max(diff) - np.min(diff))
    diff = np.abs(diff)
    blurred = None
    if use_cv2:
        try:
            import cv2

            blurred = cv2.blur(diff, (ks, ks))
        except ImportError:  # pragma: no cover
            blurred = None
    if blurred is None:
        blurred = box_blur_reflect101(diff, ks)
    pattern = np.zeros_like(diff)
    pattern[diff - blurred > threshold] = 1
    return pattern


# ----------------------------------------------------------------------------
# §8f rank 3  sim-domain IR pattern     datasets/dataset_utils.py:12-46
# ----------------------------------------------------------------------------

def ir_pattern(img_ir: np.ndarray, img: np.ndarray, threshold: float = 0.005) -> np.ndarray:
    """dataset_utils.py:12-17 (get_ir_pattern): min-max normalised |ir - no_ir| > threshold."""
    diff = np.abs(img_ir - img)
    diff = (diff - np.min(diff)) / (np.max(diff) - np.min(diff))
    out = np.zeros_like(diff)
    out[diff > threshold] = 1
    return out


def smoothed_ir_pattern2(img_ir: np.ndarray, img: np.ndarray, ks: int = 11, threshold: float = 0.005,
                         use_cv2: bool = True) -> np.ndarray:
    """dataset_utils.py:33-46 (get_smoothed_ir_pattern2): the normalised difference minus its
    INTER_AREA down(//ks)-then-up resampling, thresholded.  ``use_cv2=False`` runs the numpy
    restatement of OpenCV's two resize paths below (bit-identical to cv2 4.13 in this image)."""
    h, w = img_ir.shape
    hs, ws = int(h // ks), int(w // ks)
    diff = np.abs(img_ir - img)
    diff = (diff - np.min(diff)) / (np.max(diff) - np.min(diff))
    if use_cv2:
        import cv2

        avg = cv2.resize(diff, (ws, hs), interpolation=cv2.INTER_AREA)
        avg = cv2.resize(avg, (w, h), interpolation=cv2.INTER_AREA)
    else:
        avg = resize_area_up(resize_area_down(diff, ws, hs), w, h)
    out = np.zeros_like(diff)
    out[diff - avg > threshold] = 1
    return out


def _area_tab(ssize: int, dsize: int):
    """OpenCV computeResizeAreaTab (imgproc/resize.cpp): fractional-area weights, stored as float."""
    import math

    scale = 1.0 / (dsize / ssize)
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def resize_area_down(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), INTER_AREA) for a float64 image being SHRUNK: the integer-ratio fast
    path (block sum times a float 1/area) or the general fractional-area path, in OpenCV's summation order."""
    H, W = src.shape
    sx_f, sy_f = 1.0 / (dw / W), 1.0 / (dh / H)
    if abs(sx_f - round(sx_f)) < np.finfo(np.float64).eps and abs(sy_f - round(sy_f)) < np.finfo(np.float64).eps:
        ix, iy = int(round(sx_f)), int(round(sy_f))
        scale = np.float64(np.float32(1.0) / np.float32(ix * iy))
        # resizeAreaFast_: taps in row-major order, summed four at a time (CV_ENABLE_UNROLLED)
        taps = [src[ky:ky + dh * iy:iy, kx:kx + dw * ix:ix] for ky in range(iy) for kx in range(ix)]
        dst = np.zeros((dh, dw))
        k = 0
        while k + 4 <= len(taps):
            dst = dst + (((taps[k] + taps[k + 1]) + taps[k + 2]) + taps[k + 3])
            k += 4
        while k < len(taps):
            dst = dst + taps[k]
            k += 1
        return dst * scale
    xt, yt = _area_tab(W, dw), _area_tab(H, dh)
    sums = {}
    for dy, sy, beta in yt:
        buf = np.zeros(dw)
        for dx, sx, a in xt:
            buf[dx] = buf[dx] + src[sy, sx] * np.float64(a)
        sums[dy] = np.float64(beta) * buf if dy not in sums else sums[dy] + np.float64(beta) * buf
    dst = np.zeros((dh, dw))
    for dy, row in sums.items():
        dst[dy] = row
    return dst


def _area_up_coef(ssize: int, dsize: int):
    """OpenCV's linear-resize coefficient set-up in area mode (INTER_AREA while ENLARGING)."""
    import math

    inv = dsize / ssize
    scale = 1.0 / inv
    ofs, alpha, xmax = [], [], dsize
    for dx in range(dsize):
        sx = math.floor(dx * scale)
        fx = np.float32((dx + 1) - (sx + 1) * inv)
        fx = np.float32(0) if fx <= 0 else np.float32(fx - np.floor(fx))
        if sx < 0:
            fx, sx = np.float32(0), 0
        if sx + 1 >= ssize:
            xmax = min(xmax, dx)
            if sx >= ssize - 1:
                fx, sx = np.float32(0), ssize - 1
        ofs.append(sx)
        alpha.append((np.float32(1) - fx, fx))
    return ofs, alpha, xmax


def resize_area_up(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), INTER_AREA) for a float64 image being ENLARGED (two-tap linear with
    area-mode coefficients; horizontal pass then vertical pass)."""
    H, W = src.shape
    xo, xa, xmax = _area_up_coef(W, dw)
    yo, ya, _ = _area_up_coef(H, dh)
    hb = np.zeros((H, dw))
    for dx in range(dw):
        if dx < xmax:
            hb[:, dx] = src[:, xo[dx]] * np.float64(xa[dx][0]) + src[:, xo[dx] + 1] * np.float64(xa[dx][1])
        else:
            hb[:, dx] = src[:, xo[dx]] * 1.0
    dst = np.zeros((dh, dw))
    for dy in range(dh):
        s0, s1 = min(max(yo[dy], 0), H - 1), min(max(yo[dy] + 1, 0), H - 1)
        dst[dy] = hb[s0] * np.float64(ya[dy][0]) + hb[s1] * np.float64(ya[dy][1])
    return dst


# ----------------------------------------------------------------------------
# §8f rank 4  error metrics            utils/cascade_metrics.py:16-57
# ----------------------------------------------------------------------------

def compute_err_metric(disp_gt, depth_gt, disp_pred, focal_length, baseline, mask, depth_pred=None):
    """cascade_metrics.py:30-57: masked EPE (:30), |diff| > 1 / 2 px rates (:31-33), depth from
    disparity (:36-37), mean |dz| in mm clipped to 100 (:39-42), 2/4/8 mm outlier rates (:45-48)."""
    with torch.no_grad():
        sel_gt, sel_pred = disp_gt[mask], disp_pred[mask]
        adiff = torch.abs(sel_gt - sel_pred)
        n = adiff.numel()
        if depth_pred is None:
            depth_pred = focal_length * baseline / disp_pred
        mm = torch.clip(torch.abs(depth_gt[mask] * 1000 - depth_pred[mask] * 1000), min=0, max=100)
        dz = torch.abs(depth_gt[mask] - depth_pred[mask])
        return {
            "epe": F.l1_loss(sel_pred, sel_gt, reduction="mean").item(),
            "bad1": int((adiff > 1).sum()) / n,
            "bad2": int((adiff > 2).sum()) / n,
            "depth_abs_err": torch.mean(mm).item(),
            "depth_err2": int((dz > 2e-3).sum()) / dz.numel(),
            "depth_err4": int((dz > 4e-3).sum()) / dz.numel(),
            "depth_err8": int((dz > 8e-3).sum()) / dz.numel(),
        }


# ----------------------------------------------------------------------------
# Closed forms (derived from the lines above; used to check kernel formulas on
# tiny shapes in float64 -- SURVEY.md §8a "closed forms certified")
# ----------------------------------------------------------------------------

def sample_coords(disp: torch.Tensor, H: int, W: int, dtype=torch.float32):
    """Pixel-space sample position of ``apply_disparity(img, disp)`` in the
    reference's op order: f = lin + disp/W; g = 2f-1; xs = ((g+1)*W-1)/2."""
    lin_x = torch.linspace(0, 1, W).to(dtype)
    lin_y = torch.linspace(0, 1, H).to(dtype)
    d = disp[:, 0].to(dtype)
    fx = lin_x.view(1, 1, W) + d / W
    gx = 2 * fx - 1
    xs = ((gx + 1) * W - 1) / 2
    gy = 2 * lin_y - 1
    ys = ((gy + 1) * H - 1) / 2
    return xs, ys.view(1, H, 1).expand_as(xs)


def patch_loss_closed_form(L: np.ndarray, R: np.ndarray, disp: np.ndarray, mask: np.ndarray | None, ps: int):
    """float64 loop restatement of a7 (loss, dloss/ddisp) for one-channel
    images ``[B,1,H,W]``; ``disp`` is the *left* disparity (warp uses -disp).
    Two border rules compose: taps outside the unfolded plane contribute 0,
    and in-plane taps read the zero-padded image."""
    B, C, H, W = L.shape
    assert C == 1
    p = (ps - 1) // 2
    xs, ys = sample_coords(torch.from_numpy(-disp), H, W, torch.float64)
    xs, ys = xs.numpy(), ys.numpy()
    m = np.ones((B, H, W), bool) if mask is None else mask[:, 0].astype(bool)
    Lp = np.pad(L[:, 0].astype(np.float64), ((0, 0), (p + 1, p + 1), (p + 1, p + 1)))
    Rp = np.pad(R[:, 0].astype(np.float64), ((0, 0), (p + 1, p + 1), (p + 1, p + 1)))
    total, grad = 0.0, np.zeros((B, H, W))
    for b in range(B):
        for i in range(H):
            for j in range(W):
                if not m[b, i, j]:
                    continue
                x0 = int(np.floor(xs[b, i, j])); wx = xs[b, i, j] - x0
                y0 = int(np.floor(ys[b, i, j])); wy = ys[b, i, j] - y0
                acc = g = 0.0
                for ky in range(ps):
                    for kx in range(ps):
                        v = np.zeros((2, 2))
                        for dy in (0, 1):
                            for dx in (0, 1):
                                yy, xx = y0 + dy, x0 + dx
                                if 0 <= yy < H and 0 <= xx < W:
                                    v[dy, dx] = Rp[b, yy + ky + 1, xx + kx + 1]
                        wu = (1 - wy) * ((1 - wx) * v[0, 0] + wx * v[0, 1]) + wy * ((1 - wx) * v[1, 0] + wx * v[1, 1])
                        dwu = (1 - wy) * (v[0, 1] - v[0, 0]) + wy * (v[1, 1] - v[1, 0])
                        r = wu - Lp[b, i + ky + 1, j + kx + 1]
                        acc += r * r
                        g += r * dwu
                total += acc
                grad[b, i, j] = g
    n = ps * ps * m.sum()
    return total / n, (-2.0 / n) * grad[:, None]

"""``torch.library`` registration of the hot-path operators ("registered via torch.library so
DDP/torch.compile see them").

``activezero_b200.ops`` exposes the operators as ``torch.autograd.Function``s -- all the reference's eager
trainer needs.  This module registers the same C-ABI calls as custom ops in the ``az_stereo`` namespace, with
fake (meta) kernels and autograd formulas, so that a traced / compiled program sees opaque, shape-checked
operators instead of Python:

    torch.ops.az_stereo.concat_volume(ref, tgt, num_disp, channels_last) -> vol      (psmnet.py:151-165)
    torch.ops.az_stereo.soft_argmin(cost) -> (disp, lse)                             (psmnet.py:200-217)
    torch.ops.az_stereo.reproj_loss(tgt, src, disp, mask, ps, sign) -> (loss, warped, gpre, stats)
                                                                                    (reprojection.py:81-127)

The convenience wrappers below return what the reference-named functions return.  Nothing here computes on the
host: every implementation enqueues the CUDA kernels of ``libaz_stereo.so`` (no CPU fallback).
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from . import _lib
from .ops import _cuda_f32, _mask_u8, _ptr, _stream, check_feature_pair, check_reproj_inputs, linspace_table

NS = "az_stereo"


# ---------------------------------------------------------------------------------------------
# concat volume
# ---------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{NS}::concat_volume", mutates_args=(), device_types="cuda")
def concat_volume(ref: Tensor, tgt: Tensor, num_disp: int, channels_last: bool) -> Tensor:
    L, R = check_feature_pair(ref, tgt, "az_stereo::concat_volume")
    if num_disp <= 0:
        raise ValueError("az_stereo::concat_volume: num_disp must be positive")
    B, C, H, W = L.shape
    ndhwc = channels_last and C % 4 == 0
    vol = torch.empty((B, 2 * C, num_disp, H, W), dtype=torch.float32, device=L.device,
                      memory_format=torch.channels_last_3d if ndhwc else torch.contiguous_format)
    with torch.cuda.device(L.device):
        _lib.call("az_concat_volume_fwd_ndhwc" if ndhwc else "az_concat_volume_fwd", _ptr(L), _ptr(R), _ptr(vol),
                  B, C, H, W, num_disp, _stream())
    return vol.contiguous(memory_format=torch.channels_last_3d) if channels_last and not ndhwc else vol


@concat_volume.register_fake
def _(ref, tgt, num_disp, channels_last):
    B, C, H, W = ref.shape
    return torch.empty((B, 2 * C, num_disp, H, W), dtype=ref.dtype, device=ref.device,
                       memory_format=torch.channels_last_3d if channels_last else torch.contiguous_format)


@torch.library.custom_op(f"{NS}::concat_volume_backward", mutates_args=(), device_types="cuda")
def concat_volume_backward(gvol: Tensor, C: int) -> Tuple[Tensor, Tensor]:
    if gvol.dim() != 5 or gvol.shape[1] != 2 * C or not gvol.is_cuda or gvol.dtype != torch.float32:
        raise ValueError("az_stereo::concat_volume_backward: gvol must be a CUDA float32 [B,2C,Dq,H,W] tensor")
    B, _, Dq, H, W = gvol.shape
    ndhwc = C % 4 == 0 and gvol.is_contiguous(memory_format=torch.channels_last_3d) and not gvol.is_contiguous()
    g = gvol if ndhwc else _cuda_f32(gvol, "grad_volume")
    gL = torch.empty((B, C, H, W), dtype=torch.float32, device=g.device)
    gR = torch.empty_like(gL)
    with torch.cuda.device(g.device):
        _lib.call("az_concat_volume_bwd_ndhwc" if ndhwc else "az_concat_volume_bwd", _ptr(g), _ptr(gL), _ptr(gR),
                  B, C, H, W, Dq, _stream())
    return gL, gR


@concat_volume_backward.register_fake
def _(gvol, C):
    B, _, _, H, W = gvol.shape
    return gvol.new_empty((B, C, H, W)), gvol.new_empty((B, C, H, W))


def _concat_setup(ctx, inputs, output):
    ctx.C = inputs[0].shape[1]


def _concat_backward(ctx, gvol):
    gL, gR = concat_volume_backward(gvol, ctx.C)
    return gL, gR, None, None


concat_volume.register_autograd(_concat_backward, setup_context=_concat_setup)


# ---------------------------------------------------------------------------------------------
# soft-argmin
# ---------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{NS}::soft_argmin", mutates_args=(), device_types="cuda")
def soft_argmin_op(cost: Tensor) -> Tuple[Tensor, Tensor]:
    c = _cuda_f32(cost, "cost")
    if c.dim() != 4:
        raise ValueError("az_stereo::soft_argmin: cost must be [B,D,H,W]")
    B, D, H, W = c.shape
    disp = torch.empty((B, 1, H, W), dtype=torch.float32, device=c.device)
    lse = torch.empty((B, 2, H, W), dtype=torch.float32, device=c.device)
    with torch.cuda.device(c.device):
        _lib.call("az_soft_argmin_fwd", _ptr(c), _ptr(disp), _ptr(lse), B, D, H, W, _stream())
    return disp, lse


@soft_argmin_op.register_fake
def _(cost):
    B, _, H, W = cost.shape
    return cost.new_empty((B, 1, H, W)), cost.new_empty((B, 2, H, W))


@torch.library.custom_op(f"{NS}::soft_argmin_backward", mutates_args=(), device_types="cuda")
def soft_argmin_backward(cost: Tensor, disp: Tensor, lse: Tensor, gdisp: Tensor) -> Tensor:
    c, g = _cuda_f32(cost, "cost"), _cuda_f32(gdisp, "grad_disp")
    if c.dim() != 4:
        raise ValueError("az_stereo::soft_argmin_backward: cost must be [B,D,H,W]")
    B, D, H, W = c.shape
    d_, l_ = _cuda_f32(disp, "disp"), _cuda_f32(lse, "lse")
    if tuple(d_.shape) != (B, 1, H, W) or tuple(g.shape) != (B, 1, H, W) or tuple(l_.shape) != (B, 2, H, W) \
            or not (d_.device == l_.device == g.device == c.device):
        raise ValueError("az_stereo::soft_argmin_backward: disp/gdisp [B,1,H,W] and lse [B,2,H,W] on cost's device expected")
    disp, lse = d_, l_
    gcost = torch.empty_like(c)
    with torch.cuda.device(c.device):
        _lib.call("az_soft_argmin_bwd", _ptr(c), _ptr(disp), _ptr(lse), _ptr(g), _ptr(gcost), B, D, H, W, _stream())
    return gcost


@soft_argmin_backward.register_fake
def _(cost, disp, lse, gdisp):
    return torch.empty_like(cost)


def _sa_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], output[0], output[1])


def _sa_backward(ctx, gdisp, _glse):
    cost, disp, lse = ctx.saved_tensors
    return soft_argmin_backward(cost, disp, lse, gdisp)


soft_argmin_op.register_autograd(_sa_backward, setup_context=_sa_setup)


def soft_argmin(cost: Tensor) -> Tensor:
    """[B,D,H,W] logits -> [B,1,H,W] expected disparity (softmax + DisparityRegression)."""
    return soft_argmin_op(cost)[0]


# ---------------------------------------------------------------------------------------------
# fused warp + masked-MSE reprojection loss (ps = 1 and patch), with the warped / Fold image
# ---------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{NS}::reproj_loss", mutates_args=(), device_types="cuda")
def reproj_loss_op(tgt: Tensor, src: Tensor, disp: Tensor, mask: Tensor, ps: int, sign: float
                   ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    t, s, d = check_reproj_inputs(tgt, src, disp, "az_stereo::reproj_loss")
    m = _mask_u8(mask, t)  # [B,1,H,W] on t's device; non-zero = selected (as ops.reproj_loss)
    if ps < 1 or ps % 2 != 1:
        raise ValueError("az_stereo::reproj_loss: ps must be odd")
    B, C, H, W = t.shape
    dev = t.device
    warped, gpre = torch.empty_like(t), torch.empty_like(d)
    loss = torch.empty((1,), dtype=torch.float32, device=dev)
    stats = torch.empty((2,), dtype=torch.float64, device=dev)
    ws = torch.empty((_lib.query("az_reproj_workspace_bytes", B, H),), dtype=torch.uint8, device=dev)
    lx, ly = linspace_table(W, dev), linspace_table(H, dev)
    with torch.cuda.device(dev):
        _lib.call("az_reproj_loss_fwd", _ptr(t), _ptr(s), _ptr(d), float(sign), _ptr(m), _ptr(lx), _ptr(ly), int(ps),
                  _ptr(warped), _ptr(gpre), _ptr(loss), _ptr(stats), _ptr(ws), B, C, H, W, _stream())
    return loss.reshape(()), warped, gpre, stats


@reproj_loss_op.register_fake
def _(tgt, src, disp, mask, ps, sign):
    return (tgt.new_empty(()), torch.empty_like(tgt), torch.empty_like(disp),
            tgt.new_empty((2,), dtype=torch.float64))


@torch.library.custom_op(f"{NS}::reproj_loss_backward", mutates_args=(), device_types="cuda")
def reproj_loss_backward(gpre: Tensor, stats: Tensor, gloss: Tensor, sign: float, C: int, ps: int) -> Tensor:
    if gpre.dim() != 4 or gpre.shape[1] != 1 or not gpre.is_cuda or gpre.dtype != torch.float32 \
            or stats.numel() != 2 or stats.dtype != torch.float64 or stats.device != gpre.device:
        raise ValueError("az_stereo::reproj_loss_backward: gpre [B,1,H,W] float32 and stats float64[2] on one CUDA device expected")
    gpre = gpre.contiguous()
    B, _, H, W = gpre.shape
    gl = gloss.reshape(1).to(torch.float32).contiguous()
    gdisp = torch.empty_like(gpre)
    with torch.cuda.device(gpre.device):
        _lib.call("az_reproj_loss_bwd", _ptr(gpre), _ptr(stats), _ptr(gl), float(sign), _ptr(gdisp), B, C, H, W, int(ps),
                  _stream())
    return gdisp


@reproj_loss_backward.register_fake
def _(gpre, stats, gloss, sign, C, ps):
    return torch.empty_like(gpre)


def _rl_setup(ctx, inputs, output):
    ctx.save_for_backward(output[2], output[3])
    ctx.meta = (inputs[0].shape[1], int(inputs[4]), float(inputs[5]))


def _rl_backward(ctx, gloss, _gw, _gp, _gs):
    gpre, stats = ctx.saved_tensors
    C, ps, sign = ctx.meta
    return None, None, reproj_loss_backward(gpre, stats, gloss, sign, C, ps), None, None, None


reproj_loss_op.register_autograd(_rl_backward, setup_context=_rl_setup)


def reproj_loss(tgt: Tensor, src: Tensor, disp: Tensor, mask: Tensor, ps: int = 1, sign: float = -1.0):
    """-> (loss, warped): masked MSE between ``tgt`` and ``apply_disparity(unfold(src), sign*disp)`` and the warped
    image (ps = 1) / Fold image (ps > 1); differentiable w.r.t. ``disp``."""
    if tgt.requires_grad or src.requires_grad:
        raise ValueError("az_stereo::reproj_loss is differentiable w.r.t. disp only; use ops.warp() for image gradients")
    loss, warped, _, _ = reproj_loss_op(tgt, src, disp, mask, ps, sign)
    return loss, warped

"""CPU-side checks of the torch.library registration (activezero_b200/library_ops.py): every
operator of the path is registered in the ``az_stereo`` namespace, its fake (meta) kernel returns the shapes /
dtypes the CUDA implementation allocates, and the differentiable ones carry an autograd formula.  No compute
call is made (fake tensors only), so this runs without a GPU."""
import torch
from torch._subclasses.fake_tensor import FakeTensorMode

from activezero_b200 import library_ops as lo  # importing registers the ops

B, C, Hq, Wq, Dq, G = 2, 8, 6, 20, 5, 4
D, H, W = 4 * Dq, 4 * Hq, 4 * Wq

FORWARD_OPS = ["concat_volume", "soft_argmin", "reproj_loss", "gwc_volume", "upsample_soft_argmin", "warp"]
BACKWARD_OPS = [n + "_backward" for n in FORWARD_OPS]
PLAIN_OPS = ["scatter_warp", "temporal_ir_pattern", "local_contrast_norm"]


def test_every_operator_of_the_path_is_registered():
    for name in FORWARD_OPS + BACKWARD_OPS + PLAIN_OPS:
        op = getattr(torch.ops.az_stereo, name).default
        assert op._schema.name == f"az_stereo::{name}"
        assert not op._schema.is_mutable, name  # functional: outputs are fresh tensors, nothing is mutated


def test_differentiable_operators_have_autograd_formulas():
    defs = {d._qualname.split("::")[1]: d for d in vars(lo).values() if isinstance(d, torch.library.CustomOpDef)}
    assert sorted(defs) == sorted(FORWARD_OPS + BACKWARD_OPS + PLAIN_OPS)
    for name in FORWARD_OPS:
        assert defs[name]._backward_fn is not None and defs[name]._setup_context_fn is not None, name
    for name in BACKWARD_OPS + PLAIN_OPS:  # once-differentiable backwards; data-preparation operators
        assert defs[name]._backward_fn is None, name


def _shapes(out):
    out = out if isinstance(out, (tuple, list)) else (out,)
    return [(tuple(t.shape), t.dtype) for t in out]


def test_fake_kernels_describe_the_outputs():
    f32, i32, u8, f64 = torch.float32, torch.int32, torch.uint8, torch.float64
    ns = torch.ops.az_stereo
    with FakeTensorMode():
        e = lambda *s, dtype=f32: torch.empty(*s, device="cuda", dtype=dtype)
        L, img, disp = e(B, C, Hq, Wq), e(B, 3, H, W), e(B, 1, H, W)
        assert _shapes(ns.concat_volume(L, L, Dq, False)) == [((B, 2 * C, Dq, Hq, Wq), f32)]
        assert _shapes(ns.concat_volume_backward(e(B, 2 * C, Dq, Hq, Wq), C)) == [((B, C, Hq, Wq), f32)] * 2
        assert _shapes(ns.gwc_volume(L, L, Dq, G)) == [((B, G, Dq, Hq, Wq), f32)]
        assert _shapes(ns.gwc_volume_backward(e(B, G, Dq, Hq, Wq), L, L)) == [((B, C, Hq, Wq), f32)] * 2
        assert _shapes(ns.soft_argmin(e(B, D, H, W))) == [((B, 1, H, W), f32), ((B, 2, H, W), f32)]
        assert _shapes(ns.soft_argmin_backward(e(B, D, H, W), disp, e(B, 2, H, W), disp)) == [((B, D, H, W), f32)]
        low = e(B, 1, Dq, Hq, Wq)
        assert _shapes(ns.upsample_soft_argmin(low, D, H, W)) == [((B, 1, H, W), f32), ((B, 2, H, W), f32)]
        assert _shapes(ns.upsample_soft_argmin_backward(low, disp, e(B, 2, H, W), disp, D)) == [((B, 1, Dq, Hq, Wq), f32)]
        assert _shapes(ns.warp(img, disp)) == [((B, 3, H, W), f32)]
        assert _shapes(ns.warp_backward(img, disp, img)) == [((B, 3, H, W), f32), ((B, 1, H, W), f32)]
        out = ns.reproj_loss(disp, disp, disp, e(B, 1, H, W, dtype=torch.bool), 11, -1.0)
        assert _shapes(out) == [((), f32), ((B, 1, H, W), f32), ((B, 1, H, W), f32), ((2,), f64)]
        assert _shapes(ns.reproj_loss_backward(disp, e(2, dtype=f64), e(()), -1.0, 1, 11)) == [((B, 1, H, W), f32)]
        assert _shapes(ns.scatter_warp(disp, e(B, 1, H, W, dtype=i32))) == [((B, 1, H, W), f32), ((1,), i32)]
        assert _shapes(ns.temporal_ir_pattern(e(B, 7, H, W, dtype=u8), 11, 0.005)) == [((B, H, W), f32)]
        assert _shapes(ns.local_contrast_norm(img, 9, 1e-5)) == [((B, 1, H, W), f32)] * 2


def test_autograd_graph_builds_on_fake_tensors():
    """Forward + backward of the differentiable ops run end to end on meta tensors (shape propagation only): the
    backward of each op is the registered ``*_backward`` op, with gradients of the inputs' shapes."""
    ns = torch.ops.az_stereo
    L = torch.empty(B, C, Hq, Wq, device="meta", requires_grad=True)
    R = torch.empty(B, C, Hq, Wq, device="meta", requires_grad=True)
    low = torch.empty(B, 1, Dq, Hq, Wq, device="meta", requires_grad=True)
    img = torch.empty(B, 1, H, W, device="meta", requires_grad=True)
    disp, _ = ns.upsample_soft_argmin(low, D, H, W)
    total = ns.gwc_volume(L, R, Dq, G).sum() + ns.warp(img, disp).sum() + ns.concat_volume(L, R, Dq, False).sum()
    grads = torch.autograd.grad(total, [L, R, low, img])
    assert [tuple(g.shape) for g in grads] == [tuple(t.shape) for t in (L, R, low, img)]

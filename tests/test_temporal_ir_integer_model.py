"""CPU model of the exact integer form behind ``az_temporal_ir`` (activezero_b200/csrc/temporal_ir.cu; reference:
/root/reference/tools/temporal_ir.py:93-114 with :35-40).  The reference runs the per-pixel least-squares slope, the
min-max normalisation and ``diff - cv2.blur(diff) > 0.005`` in float64; with ``a = |sum_t y_t (2t - (T-1))|`` (an integer)
and ``R = max a - min a`` the same decision is ``ks^2 a - boxsum_ks(a) > thr ks^2 R`` with BORDER_REFLECT_101 -- integer
sums and one float64 product per image, which is what the kernels evaluate.  The model below replays that with numpy
integers (including the kernel's integer threshold ``floor(rhs) + 1``) and must agree with the oracle's float64
restatement of the reference on every pixel whose margin to the threshold exceeds 1e-6 (the parity gate of SURVEY.md
§8a row a11); on these random inputs it agrees on ALL pixels.  Runs without a GPU."""
import numpy as np
import pytest

from oracle import stereo_oracle as so


def integer_pattern(frames, ks=11, threshold=0.005):
    T, H, W = frames.shape
    w = 2 * np.arange(T, dtype=np.int64) - (T - 1)
    a = np.abs(np.tensordot(w, frames.astype(np.int64), axes=(0, 0)))          # [H,W] int64
    R = int(a.max() - a.min())
    p = ks // 2
    ap = np.pad(a, p, mode="reflect")                                          # numpy "reflect" == BORDER_REFLECT_101
    c = np.zeros((H + 2 * p + 1, W + 2 * p + 1), dtype=np.int64)
    c[1:, 1:] = ap.cumsum(0).cumsum(1)
    box = c[ks:, ks:] - c[:-ks, ks:] - c[ks:, :-ks] + c[:-ks, :-ks]            # ks x ks sums
    lhs = ks * ks * a - box
    thr = int(np.floor(threshold * float(ks * ks) * float(R))) + 1             # lhs > rhs  <=>  lhs >= floor(rhs) + 1
    return (lhs >= thr).astype(np.float64), lhs, threshold * ks * ks * R, R


@pytest.mark.parametrize("T,H,W,seed", [(7, 48, 64, 0), (4, 33, 50, 1), (7, 21, 23, 2), (2, 16, 16, 3)])
def test_integer_form_matches_the_float64_reference(T, H, W, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(H, W)) * 0.5
    dots = rng.random((H, W)) < 0.1
    frames = np.stack([np.clip(base + t * dots * 10 + rng.integers(0, 3, size=(H, W)), 0, 255) for t in range(T)]).astype(np.uint8)
    ref = so.temporal_ir_pattern(frames, ks=11, threshold=0.005)
    out, lhs, rhs, R = integer_pattern(frames)
    margin = np.abs(lhs - rhs) / (121.0 * R)  # |diff_n - blur(diff_n) - thr|, the reference's own quantity
    decided = margin > 1e-6
    assert decided.mean() > 0.99
    assert np.array_equal(out[decided], ref[decided])
    assert np.array_equal(out, ref)  # no pixel of these inputs sits inside the excluded band
    assert 0 < out.mean() < 1


def test_constant_stack_has_no_pattern():
    frames = np.full((7, 16, 20), 93, dtype=np.uint8)
    out, _, _, _ = integer_pattern(frames)  # R = 0: 0 > 0 is false everywhere (the float64 chain divides 0 by 0 here)
    assert out.sum() == 0

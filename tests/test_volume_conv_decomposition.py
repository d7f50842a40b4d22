"""CPU model of the decomposition behind `volume_conv0_v2_kernel` (csrc/volume_conv_v2.cu), in float64 on tiny shapes,
against conv3d of the materialised concat volume (oracle restatement of psmnet.py:151-168):

  out[d][y][xt + d] =   sum_{kd,ky,kx} W_L[kd,ky,kx] . L[y'][xt + d + kx - 1]            (per-plane GEMM, NO mask)
                      + sum_{kd valid} Q_kd[y][xt],   Q_kd = sum_{ky,kx} W_R[kd,ky,kx] . R[y'][xt + kx - kd]   (once per row)
                      - [left-mask fix: taps with xt + kx < kd, rows xt = 0, 1;  rows xt = -2, -1: only the allowed taps]
                      - [right-border fix: column x = W - 1, taps kx = 2 of the right half read R[W - d'] where the volume is 0]
  and out = 0 for x <= d - 3."""
import pytest
import torch
import torch.nn.functional as F

from oracle import stereo_oracle as so


def decomposed(L, R, w, Dq):
    C, H, W = L.shape
    out = torch.zeros(w.shape[0], Dq, H, W, dtype=torch.float64)
    wl, wr = w[:, :C], w[:, C:]

    def Lat(yp, xp):
        return L[:, yp, xp] if 0 <= yp < H and 0 <= xp < W else torch.zeros(C, dtype=torch.float64)

    def Rat(yp, xr):
        return R[:, yp, xr] if 0 <= yp < H and 0 <= xr < W else torch.zeros(C, dtype=torch.float64)

    for y in range(H):
        for xt in range(-2, W):
            # right half: independent of d (zero prefix of R = the mask xt + kx >= kd)
            Q = [sum(wr[:, :, kd, ky, kx] @ Rat(y + ky - 1, xt + kx - kd) for ky in range(3) for kx in range(3)) for kd in range(3)]
            for d in range(Dq):
                x = xt + d
                if not 0 <= x < W:
                    continue
                kds = [kd for kd in range(3) if 0 <= d + kd - 1 < Dq]
                # left half, every tap of the valid planes, unmasked -- what the tensor core accumulates
                acc = sum(wl[:, :, kd, ky, kx] @ Lat(y + ky - 1, x + kx - 1) for kd in kds for ky in range(3) for kx in range(3))
                masked = sum(wl[:, :, kd, ky, kx] @ Lat(y + ky - 1, x + kx - 1) for kd in kds for ky in range(3) for kx in range(3)
                             if xt + kx < kd)
                val = acc - masked + sum(Q[kd] for kd in kds)      # (rows xt < 0: acc - masked = the allowed taps only)
                if x == W - 1:                                     # the volume is zero at x' = W, the shifted operand is not
                    val = val - sum(wr[:, :, kd, ky, 2] @ Rat(y + ky - 1, W - (d + kd - 1)) for kd in kds for ky in range(3)
                                    if d + kd - 1 >= 1)
                out[:, d, y, x] = val
    return out  # columns x <= d - 3 stay zero


@pytest.mark.parametrize("C,H,W,Dq", [(3, 4, 9, 5), (2, 1, 6, 3), (2, 3, 4, 6), (1, 2, 12, 1)])
def test_shifted_coordinate_decomposition_equals_conv3d_of_the_volume(C, H, W, Dq):
    torch.manual_seed(C * 100 + W)
    L, R = torch.randn(C, H, W, dtype=torch.float64), torch.randn(C, H, W, dtype=torch.float64)
    w = torch.randn(5, 2 * C, 3, 3, 3, dtype=torch.float64)
    vol = so.concat_volume(L[None], R[None], Dq)
    ref = F.conv3d(vol, w, padding=1)[0]
    torch.testing.assert_close(decomposed(L, R, w, Dq), ref, rtol=1e-11, atol=1e-11)

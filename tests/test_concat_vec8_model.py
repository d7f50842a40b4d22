"""CPU model of the index logic of ``concat_fwd_vec8_kernel`` (activezero_b200/csrc/concat_volume.cu, the 256-bit
load/store form of the concat volume forward, /root/reference/nets/psmnet/psmnet.py:151-165): a thread owns the octet
``x .. x+7`` of one feature row and writes it to the planes ``i0 .. i0+7`` (``i0`` a multiple of 8); the right half's
shift by ``i = i0 + r`` is the window ``[8-r, 16-r)`` of the two ALIGNED octets at ``x-i0-8`` and ``x-i0``, each either
wholly inside the row or wholly left of it (then zero).  The model replays exactly those loads and windows with numpy
and must reproduce the oracle's volume bit for bit -- documentation of the kernel's arithmetic that runs without a GPU."""
import numpy as np
import pytest
import torch

from oracle import stereo_oracle as so


def vec8_model(L, R, Dq):
    B, C, H, W = L.shape
    assert W % 8 == 0
    vol = np.full((B, 2 * C, Dq, H, W), np.nan, dtype=np.float32)  # NaN: every element must be written
    for b in range(B):
        for oc in range(2 * C):
            right, c = oc >= C, oc % C
            src = (R if right else L)[b, c]
            for i0 in range(0, Dq, 8):
                i1 = min(Dq, i0 + 8)
                for y in range(H):
                    for x in range(0, W, 8):
                        if not right:
                            v = src[y, x:x + 8]
                            for i in range(i0, i1):
                                o = v.copy()
                                for k in range(8):
                                    if x + k < i:
                                        o[k] = 0.0
                                vol[b, oc, i, y, x:x + 8] = o
                        else:
                            ab = np.zeros(16, dtype=np.float32)
                            if x - i0 >= 0:
                                ab[8:] = src[y, x - i0:x - i0 + 8]
                            if x - i0 - 8 >= 0:
                                ab[:8] = src[y, x - i0 - 8:x - i0]
                            for r in range(i1 - i0):
                                vol[b, oc, i0 + r, y, x:x + 8] = ab[8 - r:16 - r]
    return vol


@pytest.mark.parametrize("shape,dq", [((1, 2, 3, 16), 12), ((2, 1, 2, 8), 8), ((1, 1, 2, 24), 29), ((1, 2, 1, 32), 5)])
def test_vec8_windows_reproduce_the_volume(shape, dq):
    torch.manual_seed(3)
    L, R = torch.randn(shape), torch.randn(shape)
    ref = so.concat_volume(L, R, dq).numpy()
    out = vec8_model(L.numpy(), R.numpy(), dq)
    assert np.array_equal(out, ref)

"""CPU model of the travelling-accumulator schedule of `gwc_bwd_systolic_kernel` (csrc/gwc_volume.cu): the index logic of
the kernel, executed step by step in numpy for one image row and one channel, against the closed form
    gR[x] = sum_{i < Dq, x + i < W} g[i][x + i] * L[x + i].
Thread t owns the column quad 4t..4t+3.  At step m (planes 4m..4m+3) it holds the accumulator of target quad Q_{t-m}:
contributions with j >= r go to it before the hand-over ("high"), contributions with j < r after it ("low"); quads
whose walk leaves the row are written by the row's last thread; after M = ceil(Dq/4) steps thread t holds Q_{t-M}."""
import numpy as np
import pytest


def systolic_gR(g, L):
    Dq, W = g.shape
    W4, M = W // 4, (Dq + 3) // 4
    out = np.full(W, np.nan)
    V = np.zeros((W4, 4))
    for m in range(M):
        low = np.zeros((W4, 4))
        for t in range(W4):
            for r in range(4):
                i = 4 * m + r
                if i >= Dq:
                    continue
                for j in range(4):
                    p = g[i, 4 * t + j] * L[4 * t + j]
                    if j >= r:
                        V[t, j - r] += p          # "high": the quad this thread holds
                    else:
                        low[t, 4 + j - r] += p    # "low": the quad that arrives with the hand-over
        k = W4 - 1 - m                            # the walk of Q_k ends at the row's last thread
        if k >= 0:
            out[4 * k:4 * k + 4] = V[W4 - 1]
        recv = np.zeros_like(V)
        recv[1:] = V[:-1]                         # shfl_up by one thread; thread 0 of a row receives zero
        V = recv + low
    for t in range(W4):
        k = t - M
        if k >= 0:
            out[4 * k:4 * k + 4] = V[t]
    return out


@pytest.mark.parametrize("W,Dq", [(240, 48), (60, 48), (36, 13), (24, 30), (16, 8), (8, 8), (4, 9), (128, 47)])
def test_travelling_accumulators_reproduce_the_closed_form(W, Dq):
    rng = np.random.default_rng(W * 131 + Dq)
    g, L = rng.standard_normal((Dq, W)), rng.standard_normal(W)
    ref = np.array([sum(g[i, x + i] * L[x + i] for i in range(Dq) if x + i < W) for x in range(W)])
    out = systolic_gR(g, L)
    assert not np.isnan(out).any()      # every target quad is written exactly once
    np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12)

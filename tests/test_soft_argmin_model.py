"""CPU model of the one-pass soft-argmin that ``soft_argmin_fwd*_kernel`` implement (activezero_b200/csrc/soft_argmin.cu;
reference: ``F.softmax(cost, 1)`` + ``DisparityRegression``, /root/reference/nets/psmnet/psmnet.py:200-201 and
psmnet_submodule.py:80-89): online softmax in chunks of 8 planes, exponents formed in fp32 from the EXACT difference to
the running maximum, one rescale per chunk when the maximum moves, chunk sums as fp32 trees, running sums in fp64.  The
model replays that arithmetic with numpy -- including a 2-ulp perturbation of every exponential, the error bound of the
MUFU ``ex2.approx`` the kernel uses -- and is held to the 1e-4 px gate of SURVEY.md §8a row a4 against an fp64 softmax on
the distributions that stress it: flat, peaked, huge magnitudes, a maximum that moves in every chunk, leading -inf planes.
Documentation of the accuracy argument that runs without a GPU; the kernels themselves are checked in tests/test_gpu_*."""
import numpy as np
import pytest

F32 = np.float32
LOG2E = F32(1.4426950408889634)
CHUNK = 8


def ex2_approx(t, rng):
    """2^t in fp32 with up to 2 ulp of error (ex2.approx.ftz.f32); denormal results flushed to zero."""
    with np.errstate(under="ignore", over="ignore"):
        y = np.exp2(t.astype(np.float64)).astype(F32)
    y = (y.astype(np.float64) * (1.0 + rng.uniform(-2.0, 2.0, size=y.shape) * 2.0 ** -24)).astype(F32)
    y[np.abs(y) < np.finfo(F32).tiny] = 0
    return y


def online_soft_argmin(cost, rng):
    """cost [D, N] float32 -> [N] float32, the kernels' arithmetic."""
    D, N = cost.shape
    m = np.full(N, -np.inf, dtype=F32)
    s = np.zeros(N, dtype=np.float64)
    ws = np.zeros(N, dtype=np.float64)
    for d0 in range(0, D, CHUNK):
        x = np.full((CHUNK, N), -np.inf, dtype=F32)
        n = min(CHUNK, D - d0)
        x[:n] = cost[d0:d0 + n]
        cm = x.max(axis=0)
        move = cm > m
        with np.errstate(invalid="ignore"):
            sc = ex2_approx(((m - cm) * LOG2E).astype(F32), rng).astype(np.float64)  # exp2(-inf) = 0 on the first chunk
        sc = np.where(np.isnan(sc), 0.0, sc)
        sc = np.where(move, sc, 1.0)  # the kernel rescales only when the maximum moves
        s = s * sc
        ws = ws * sc
        m = np.where(move, cm, m)
        ref = np.where(np.isinf(m), F32(0), m).astype(F32)
        e = ex2_approx(((x - ref).astype(F32) * LOG2E).astype(F32), rng)
        k = np.arange(CHUNK, dtype=F32)[:, None]
        # fp32 trees of the chunk (pairwise), as in the kernel
        def tree(v):
            a = (v[0] + v[1]).astype(F32), (v[2] + v[3]).astype(F32), (v[4] + v[5]).astype(F32), (v[6] + v[7]).astype(F32)
            return ((a[0] + a[1]).astype(F32) + (a[2] + a[3]).astype(F32)).astype(F32)
        cs = tree(e)
        cw = tree((e * k).astype(F32))
        s = s + cs.astype(np.float64)
        ws = ws + cw.astype(np.float64) + float(d0) * cs.astype(np.float64)
    return (ws / s).astype(F32)


def fp64_soft_argmin(cost):
    c = cost.astype(np.float64)
    c = c - c.max(axis=0, keepdims=True)
    p = np.exp(c)
    p /= p.sum(axis=0, keepdims=True)
    return (p * np.arange(cost.shape[0], dtype=np.float64)[:, None]).sum(axis=0)


def cases(rng, D=192, N=4096):
    yield "N(0,1)", rng.standard_normal((D, N)).astype(F32)
    yield "N(0,10^2): peaked", (rng.standard_normal((D, N)) * 10).astype(F32)
    yield "N(0,1000^2): one-hot", (rng.standard_normal((D, N)) * 1000).astype(F32)
    yield "offset 1e6", (rng.standard_normal((D, N)) * 30 + 1e6).astype(F32)
    yield "flat zeros", np.zeros((D, N), dtype=F32)
    yield "flat with fp32 noise", (1.0 + rng.standard_normal((D, N)) * 1e-6).astype(F32)
    ramp = np.arange(D, dtype=F32)[:, None] * F32(0.75) + rng.standard_normal((D, N)).astype(F32) * F32(0.1)
    yield "rising ramp: the maximum moves in every chunk", ramp
    yield "falling ramp: the first chunk holds the maximum", ramp[::-1].copy()
    lead = rng.standard_normal((D, N)).astype(F32)
    lead[:20] = -np.inf
    yield "20 leading -inf planes", lead
    yield "D = 50 (ragged last chunk)", rng.standard_normal((50, N)).astype(F32) * 3


@pytest.mark.parametrize("seed", [0, 1])
def test_one_pass_soft_argmin_meets_the_gate(seed):
    rng = np.random.default_rng(seed)
    for name, cost in cases(rng):
        out = online_soft_argmin(cost, rng)
        ref = fp64_soft_argmin(cost)
        err = float(np.abs(out.astype(np.float64) - ref).max())
        assert np.isfinite(out).all(), name
        assert err <= 1e-4, f"{name}: {err:.3e} px"   # <= 1e-4 px
        assert err <= 3e-5, f"{name}: {err:.3e} px"   # what the fp64 running sums actually deliver

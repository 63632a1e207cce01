"""GPU parity of the two ORT custom ops (through the C ABI) against the oracle and the reference fixtures.

Tolerance (north_star): <= 1e-4 relative for Correlation / Warp fp32 tensors, measured norm-wise as
max|d| / max|ref| (BASELINE.md "Parity gates").  Observed values are ~1e-7 (fp32 rounding: FMA vs mul+add).
"""
import os

import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ops_golden.npz")
TOL = 1e-4


def rel(a, ref):
    return float(np.abs(a - ref).max() / max(float(np.abs(ref).max()), 1e-30))


def cu(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


# ---------------------------------------------------------------- Correlation
@pytest.mark.parametrize("shape", [
    (1, 1, 1, 1), (1, 3, 5, 7), (2, 12, 13, 21), (1, 8, 8, 32), (1, 9, 33, 65), (1, 64, 18, 30), (1, 196, 9, 15),
    (3, 16, 40, 44), (1, 7, 72, 120),
])
def test_correlation_vs_oracle(V, O, dev, shape):
    N, C, H, W = shape
    a, b = synth.features(N, C, H, W, 1), synth.features(N, C, H, W, 2)
    got = V.correlation(cu(a, dev), cu(b, dev)).cpu().numpy()
    ref = O.correlation(a, b)
    assert got.shape == (N, 9, 9, H, W)
    assert rel(got, ref) <= TOL


@pytest.mark.parametrize("C", [16, 12, 96])
def test_correlation_legacy_layout(V, O, dev, C):
    a, b = synth.features(2, C, 14, 20, 3), synth.features(2, C, 14, 20, 4)
    got = V.correlation(cu(a, dev), cu(b, dev), legacy=True).cpu().numpy()
    ref = O.correlation(a, b, legacy=True)
    assert got.shape == (2, 81, 14, 20)
    assert rel(got, ref) <= TOL
    # same bytes as the new layout up to the 1/C scale
    new = V.correlation(cu(a, dev), cu(b, dev)).cpu().numpy().reshape(2, 81, 14, 20)
    assert rel(got, new / np.float32(C)) <= 1e-6


@pytest.mark.parametrize("md", [0, 1, 2, 6])
def test_correlation_other_displacements(V, O, dev, md):
    a, b = synth.features(1, 10, 11, 17, 5), synth.features(1, 10, 11, 17, 6)
    got = V.correlation(cu(a, dev), cu(b, dev), max_displacement=md).cpu().numpy()
    ref = O.correlation(a, b, max_displacement=md)
    assert got.shape == ref.shape
    assert rel(got, ref) <= TOL


def test_correlation_golden(V, dev):
    g = np.load(GOLD)
    for name in ("corr_ragged", "corr_level", "corr_testpy"):
        got = V.correlation(cu(g[name + "_in1"], dev), cu(g[name + "_in2"], dev)).cpu().numpy()
        assert rel(got, g[name + "_out"]) <= TOL, name


def test_correlation_reference_test_shape(V, O, dev):
    """model-conversion/test.py:47-48: rand(4,64,128,128)*10, run twice on the same stream."""
    rng = np.random.default_rng(0)
    a = rng.random((4, 64, 128, 128), dtype=np.float32) * 10
    b = rng.random((4, 64, 128, 128), dtype=np.float32) * 10
    ta, tb = cu(a, dev), cu(b, dev)
    got1 = V.correlation(ta, tb)
    got2 = V.correlation(ta, tb)
    assert torch.equal(got1, got2)
    # oracle on one batch element (seconds on the host)
    ref = O.correlation(a[:1], b[:1])
    assert rel(got1[:1].cpu().numpy(), ref) <= TOL


def test_correlation_full_size_properties(V, dev):
    """dense-4K level 2 (32,544,960): properties that need no CPU reference."""
    C, H, W = 32, 544, 960
    g = torch.Generator(device=dev).manual_seed(3)
    a = torch.randn((1, C, H, W), device=dev, generator=g)
    b = torch.randn((1, C, H, W), device=dev, generator=g)
    out = V.correlation(a, b)
    # centre displacement == channel dot product
    dot = (a * b).sum(1)
    assert float((out[:, 4, 4] - dot).abs().max() / dot.abs().max()) <= 1e-5
    # displacement (ph,pw) == dot product with the shifted second input, zero outside
    for ph, pw in ((0, 0), (8, 8), (2, 7), (6, 1)):
        dy, dx = ph - 4, pw - 4
        sh = torch.zeros_like(b)
        ys, ye = max(0, -dy), min(H, H - dy)
        xs, xe = max(0, -dx), min(W, W - dx)
        sh[:, :, ys:ye, xs:xe] = b[:, :, ys + dy:ye + dy, xs + dx:xe + dx]
        ref = (a * sh).sum(1)
        assert float((out[:, ph, pw] - ref).abs().max() / ref.abs().max()) <= 1e-5
    # linearity in the first argument
    out2 = V.correlation(2.0 * a, b)
    assert torch.equal(out2, 2.0 * out)
    # symmetry: corr(a,b)[ph,pw](h,w) == corr(b,a)[8-ph,8-pw](h+dy,w+dx)
    outT = V.correlation(b, a)
    assert float((out[0, 6, 3, 0:H - 2, 1:W] - outT[0, 2, 5, 2:H, 0:W - 1]).abs().max()) <= 1e-3


# ---------------------------------------------------------------- Warp
@pytest.fixture(params=[0, 1, 2, 3, 1 | (2 << 4), 4, 4 | (3 << 4)],
                ids=["auto", "linear", "tiled", "linear-quad", "linear-2chunks", "tma-staged", "tma-staged-3chunks"])
def warp_mode(V, request):
    """every Warp test runs on both kernels (vsc_set_warp_mode)"""
    assert V.lib().vsc_set_warp_mode(request.param) == 0
    yield request.param
    V.lib().vsc_set_warp_mode(0)


@pytest.mark.parametrize("shape,sigma", [
    ((1, 1, 2, 2), 1.0), ((2, 5, 13, 21), 5.0), ((1, 8, 16, 12), 0.5), ((1, 64, 18, 30), 2.0), ((2, 3, 33, 64), 3.0),
    ((1, 4, 7, 100), 40.0), ((1, 96, 36, 60), 2.0), ((1, 3, 1, 1), 0.3), ((2, 2, 1, 9), 0.8), ((1, 2, 9, 1), 0.8),
    ((1, 5, 2, 40), 1.5), ((3, 9, 40, 2), 1.5),
])
def test_warp_vs_oracle(V, O, dev, shape, sigma, warp_mode):
    N, C, H, W = shape
    x = synth.features(N, C, H, W, 7)
    f = synth.op_flow(N, H, W, 8, sigma)
    got = V.warp(cu(x, dev), cu(f, dev)).cpu().numpy()
    ref = O.warp_nchw(x, f)
    assert rel(got, ref) <= TOL
    # the validity decision (mask > 0.999) is reproduced exactly: same zero pattern
    assert np.array_equal(got == 0, ref == 0)


def test_warp_golden(V, dev, warp_mode):
    g = np.load(GOLD)
    for name in ("warp_big", "warp_testpy", "warp_edge"):
        got = V.warp(cu(g[name + "_in"], dev), cu(g[name + "_flow"], dev)).cpu().numpy()
        ref = g[name + "_out"]
        assert rel(got, ref) <= TOL, name
        assert np.array_equal(got == 0, ref == 0), name


def test_warp_nonfinite(V, O, dev, warp_mode):
    """NaN / inf flow and non-finite input values outside the sampled taps must behave like the reference."""
    x = synth.features(1, 2, 6, 8, 9)
    f = synth.op_flow(1, 6, 8, 10, 1.0)
    f[0, 0, 2, 3] = np.nan
    f[0, 1, 4, 1] = np.inf
    f[0, 0, 0, 0] = -1e30
    got = V.warp(cu(x, dev), cu(f, dev)).cpu().numpy()
    ref = O.warp_nchw(x, f)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert rel(np.nan_to_num(got), np.nan_to_num(ref)) <= TOL


def test_warp_identity_and_shift(V, dev, warp_mode):
    """size-independent properties at the dense-4K level-2 shape: zero flow is the identity; an integer shift
    is a translation with zeros where the source leaves the image."""
    C, H, W = 32, 544, 960
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn((1, C, H, W), device=dev, generator=g)
    z = torch.zeros((1, 2, H, W), device=dev)
    assert torch.equal(V.warp(x, z), x)
    f = z.clone()
    f[:, 0] = 3.0
    f[:, 1] = -2.0
    out = V.warp(x, f)
    ref = torch.zeros_like(x)
    ref[:, :, 2:H, 0:W - 3] = x[:, :, 0:H - 2, 3:W]
    assert torch.equal(out, ref)


def test_warp_kernels_agree(V, dev):
    """the tiled kernel (shared 2x2 gather quad, border corners re-slotted) equals the one-pixel-per-thread
    kernel value for value, including every border case a large random flow produces"""
    g = torch.Generator(device=dev).manual_seed(17)
    # (the shapes with W % 4 == 0, W >= 72, H >= 12 run the TMA-staged kernel in mode 4: smooth flow -> staged tiles,
    # random flow -> its gather path, and (50, 148) / (33, 200) have partial tiles at the right and bottom edges)
    for (N, C, H, W) in ((2, 12, 67, 131), (1, 7, 5, 3), (1, 4, 2, 2), (1, 33, 40, 64), (2, 9, 41, 100), (1, 5, 70, 33),
                         (2, 13, 50, 148), (1, 8, 33, 200), (1, 32, 136, 240), (2, 6, 64, 128), (1, 12, 544, 960), (2, 7, 300, 480)):
        x = torch.randn((N, C, H, W), device=dev, generator=g)
        if (H, W) in ((41, 100), (70, 33), (50, 148), (136, 240), (544, 960)):
            # smooth flow: neighbouring lanes sample neighbouring taps (the shuffle kernel's fast path), with a few
            # NaN / huge displacements so that live and dead lanes alternate inside a warp
            f = torch.from_numpy(synth.op_flow_smooth(N, H, W, 3, amp=3.0, noise=0.05)).to(dev)
            f[:, 0, 5::11, 7::13] = float("nan")
            f[:, 1, 3::7, 2::9] = 1e9
        else:
            f = 6.0 * torch.randn((N, 2, H, W), device=dev, generator=g)
            f[:, :, ::3, ::2] = torch.round(f[:, :, ::3, ::2])  # integer displacements: alpha/beta exactly 0 at borders
        try:
            assert V.lib().vsc_set_warp_mode(1) == 0
            a = V.warp(x, f)
            assert V.lib().vsc_set_warp_mode(2) == 0
            b = V.warp(x, f)
            assert V.lib().vsc_set_warp_mode(3) == 0
            c = V.warp(x, f)
            assert V.lib().vsc_set_warp_mode(1 | (1 << 4)) == 0   # all channels in one chunk
            d = V.warp(x, f)
            assert V.lib().vsc_set_warp_mode(1 | (5 << 4)) == 0   # 5 channel chunks
            e = V.warp(x, f)
            assert V.lib().vsc_set_warp_mode(4) == 0              # TMA-staged tiles where applicable
            s4 = V.warp(x, f)
            assert V.lib().vsc_set_warp_mode(4 | (2 << 4)) == 0
            s5 = V.warp(x, f)
        finally:
            V.lib().vsc_set_warp_mode(0)
        assert torch.equal(a, b), (N, C, H, W)
        assert torch.equal(a, c), (N, C, H, W)
        assert torch.equal(a, d), (N, C, H, W)
        assert torch.equal(a, e), (N, C, H, W)
        # NaN != NaN: compare bit patterns
        assert torch.equal(a.view(torch.int32), s4.view(torch.int32)), (N, C, H, W)
        assert torch.equal(a.view(torch.int32), s5.view(torch.int32)), (N, C, H, W)
    assert V.lib().vsc_set_warp_mode(5) == -1


def test_ops_reject_bad_arguments(V, dev):
    a = torch.zeros((1, 2, 4, 4), device=dev)
    with pytest.raises(V.VscError):
        V.correlation(a, torch.zeros((1, 2, 4, 5), device=dev))
    with pytest.raises(V.VscError):
        V.warp(a, torch.zeros((1, 3, 4, 4), device=dev))
    with pytest.raises(V.VscError):
        V.warp(a.cpu(), torch.zeros((1, 2, 4, 4)))  # no CPU fallback
    assert V.lib().vsc_correlation_f32(None, None, None, 1, 1, 1, 1, 4, 0, None) == -1


@pytest.mark.parametrize("shape", [(1, 8, 8, 32), (2, 12, 16, 20), (1, 196, 36, 60), (1, 64, 72, 120), (1, 5, 9, 44),
                                   (1, 32, 136, 240), (2, 20, 23, 100), (1, 196, 34, 60)])
@pytest.mark.parametrize("legacy", [False, True])
def test_correlation_tma_and_plain_stagers_agree_bit_for_bit(V, dev, shape, legacy):
    """W % 4 == 0: the TMA-staged kernel (default) and the plain-load stager run the same contraction."""
    N, C, H, W = shape
    a, b = cu(synth.features(N, C, H, W, 21), dev), cu(synth.features(N, C, H, W, 22), dev)
    L = V.lib()
    try:
        assert L.vsc_set_correlation_mode(1) == 0
        plain = V.correlation(a, b, legacy=legacy)
        assert L.vsc_set_correlation_mode(2) == 0   # TMA, 32x8 tiles, 2 CTAs per SM
        tma32 = V.correlation(a, b, legacy=legacy)
        assert L.vsc_set_correlation_mode(3) == 0   # TMA, 64x8 tiles, software-pipelined, 1 CTA per SM
        tma64 = V.correlation(a, b, legacy=legacy)
        assert L.vsc_set_correlation_mode(5) == 0   # TMA, skewed 64x8 tiles, lane pairs share their in2 rows
        share = V.correlation(a, b, legacy=legacy)
        assert L.vsc_set_correlation_mode(6) == 0   # the same with 32x8 tiles, two CTAs of 6 warps per SM
        share32 = V.correlation(a, b, legacy=legacy)
        assert L.vsc_set_correlation_mode(4) == 0   # channel-split kernel (small maps)
        split = V.correlation(a, b, legacy=legacy)
        assert L.vsc_set_correlation_mode(7) == 0   # channel-split kernel with a 4-pixel register tile (W % 4 == 0)
        quads = V.correlation(a, b, legacy=legacy)
        assert L.vsc_set_correlation_mode(0) == 0
        auto = V.correlation(a, b, legacy=legacy)
        auto2 = V.correlation(a, b, legacy=legacy)  # twice on the same stream (reference test.py:70-71)
    finally:
        L.vsc_set_correlation_mode(0)
    torch.cuda.synchronize()
    assert torch.equal(tma32, plain)
    assert torch.equal(tma64, plain)
    assert torch.equal(share, plain)
    assert torch.equal(share32, plain)
    tol = 5e-6 * max(float(plain.abs().max()), 1e-30)   # 8 partial sums per value: equal within rounding
    assert float((split - plain).abs().max()) <= tol
    assert float((quads - plain).abs().max()) <= tol
    if H * W <= 12288:   # auto = channel-split kernel (its quad form when the rows are whole quads and aligned)
        assert torch.equal(auto, quads if W % 4 == 0 else split)
    else:
        assert torch.equal(auto, plain)
    assert torch.equal(auto, auto2)

"""CPU tests that pin the oracle (oracle/vsc_oracle.c) -- no GPU, no /root/reference needed at run time.

  * custom ops: against the committed fixtures produced by the reference's own CPU kernels
    (tests/golden/ops_golden.npz, generator make_ops_golden.py), against those kernels directly when
    oracle/_ref/libvsc_ref_cpu.so is present, and against the torch definitions the reference's test.py uses;
  * stabilization: against fixtures produced by the reference's own CUDA kernels on a B200
    (tests/golden/stab_golden.npz, generator make_stab_golden.py).
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import synth

GOLD_OPS = os.path.join(os.path.dirname(__file__), "golden", "ops_golden.npz")
GOLD_STAB = os.path.join(os.path.dirname(__file__), "golden", "stab_golden.npz")


def rel(a, ref):
    return float(np.abs(a - ref).max() / max(float(np.abs(ref).max()), 1e-30))


# ---------------------------------------------------------------- custom ops
def test_oracle_ops_match_reference_fixtures_bit_for_bit(O):
    g = np.load(GOLD_OPS)
    for name in ("corr_ragged", "corr_level", "corr_testpy"):
        assert np.array_equal(O.correlation(g[name + "_in1"], g[name + "_in2"]), g[name + "_out"]), name
    for name in ("warp_big", "warp_testpy", "warp_edge"):
        assert np.array_equal(O.warp_nchw(g[name + "_in"], g[name + "_flow"]), g[name + "_out"]), name


def test_oracle_ops_match_reference_cpu_library(O):
    if not O.ref_cpu_available():
        pytest.skip("oracle/_ref/libvsc_ref_cpu.so not built (needs /root/reference at build time)")
    for seed, (N, C, H, W) in enumerate([(1, 1, 1, 1), (2, 7, 9, 13), (1, 32, 20, 36), (1, 196, 9, 15)]):
        a, b = synth.features(N, C, H, W, seed), synth.features(N, C, H, W, seed + 50)
        assert np.array_equal(O.correlation(a, b), O.ref_cpu_correlation(a, b))
        for sigma in (0.3, 3.0, 30.0):
            f = synth.op_flow(N, H, W, seed + 70, sigma)
            assert np.array_equal(O.warp_nchw(a, f), O.ref_cpu_warp(a, f))


def test_reference_cpu_error_behaviour(O):
    """correlation.h:19-31,45-46: missing attributes and legacy-on-CPU throw std::runtime_error."""
    if not O.ref_cpu_available():
        pytest.skip("oracle/_ref/libvsc_ref_cpu.so not built")
    L = O.ref_cpu()
    assert L.vsc_ref_cpu_correlation_ctor_throws(1, 1) == 0
    assert L.vsc_ref_cpu_correlation_ctor_throws(0, 1) == 1
    assert L.vsc_ref_cpu_correlation_ctor_throws(1, 0) == 1
    a = synth.features(1, 2, 4, 4, 0)
    with pytest.raises(RuntimeError):
        O.ref_cpu_correlation(a, a, 4, legacy=1)


def test_correlation_definition_spatial_correlation_sample(O):
    """test.py:81: the op equals spatial_correlation_sample(patch_size=9) == pad + shifted channel dot products."""
    a, b = synth.features(2, 16, 24, 40, 1), synth.features(2, 16, 24, 40, 2)
    ta, tb = torch.from_numpy(a), torch.from_numpy(b)
    bp = F.pad(tb, (4, 4, 4, 4))
    ref = torch.empty(2, 9, 9, 24, 40)
    for ph in range(9):
        for pw in range(9):
            ref[:, ph, pw] = (ta * bp[:, :, ph:ph + 24, pw:pw + 40]).sum(1)
    assert rel(O.correlation(a, b), ref.numpy()) <= 1e-6
    # legacy layout: same values / C, channel = ph*9 + pw
    leg = O.correlation(a, b, legacy=True)
    assert rel(leg.reshape(2, 9, 9, 24, 40), ref.numpy() / 16) <= 1e-6


def test_warp_definition_masked_grid_sample(O):
    """test.py:99-136: the op equals grid_sample(align_corners=True) times (bilinear mask >= 0.999)."""
    N, C, H, W = 2, 6, 24, 40
    x = synth.features(N, C, H, W, 3)
    f = synth.op_flow(N, H, W, 4, 5.0)
    tx, tf = torch.from_numpy(x), torch.from_numpy(f)
    gy, gx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    vx = 2.0 * (gx[None] + tf[:, 0]) / max(W - 1, 1) - 1.0
    vy = 2.0 * (gy[None] + tf[:, 1]) / max(H - 1, 1) - 1.0
    grid = torch.stack((vx, vy), dim=3)
    out = F.grid_sample(tx, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    mask = F.grid_sample(torch.ones_like(tx), grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    ref = (out * (mask > 0.999)).numpy()
    got = O.warp_nchw(x, f)
    bad = np.abs(got - ref) > 1e-4 * np.abs(ref).max()
    assert bad.mean() < 1e-3  # only values sitting on the 0.999 threshold may differ


# ---------------------------------------------------------------- stabilization kernels: internal consistency
def test_u8_conversion_definitions(O):
    rgba = np.zeros((1, 256, 4), np.uint8)
    rgba[0, :, 0] = np.arange(256)
    rgba[0, :, 1] = np.arange(256)[::-1]
    f = O.rgba8_to_f32x3(rgba)
    # gpuimage.cu:49: float(double(v)/255.0); equal to the float division for every byte value
    assert np.array_equal(f[0, :, 0], (np.arange(256) / 255.0).astype(np.float32))
    assert np.array_equal(f[0, :, 0], np.arange(256, dtype=np.float32) / np.float32(255.0))
    back = O.f32x3_to_rgba8(f)
    assert (back[..., 3] == 1).all()
    # floor(|v|*255) is NOT a round trip: it may lose one level (reference behaviour)
    assert (np.abs(back[0, :, 0].astype(int) - np.arange(256)) <= 1).all()
    odd = np.array([[[-0.5, 1.5, 300.0 / 255.0]]], np.float32)
    got = O.f32x3_to_rgba8(odd)[0, 0, :3]
    assert list(got) == [127, (382) % 256, int(np.floor(np.float32(300.0 / 255.0) * np.float32(255.0))) % 256]


def test_bilinear_even_downsample_is_point_sampling(O):
    img = synth.f32_images(64, 48, 2, 1)[0]
    assert np.array_equal(O.bilinear(img, 32, 24), img[::2, ::2])
    up = O.bilinear(img[::2, ::2].copy(), 64, 48)
    assert np.array_equal(up[::2, ::2], img[::2, ::2])


def test_warp_hwc3_zero_flow_identity_inside(O):
    img = synth.f32_images(40, 30, 3, 1)[0]
    z = np.zeros((30, 40, 3), np.float32)
    out = O.warp_hwc3(img, z)
    # clamp to W-3/H-3 (flowconsistency.cu:95-96): identity only for x <= W-3, y <= H-3
    assert np.array_equal(out[:28, :38], img[:28, :38])
    assert np.array_equal(out[:28, 39], img[:28, 37])


def test_solver_fixed_point_and_orderings(O):
    W, H = 40, 30
    pr, tg, wt0 = synth.f32_images(W, H, 5, 3)
    # tgt == pr: the processed frame is already the minimiser -> nothing moves
    out = O.consist_out(pr, pr, wt0, 20, 0.15, 0.15, pr)
    assert np.abs(out - pr).max() <= 1e-6
    wt = (wt0 * 2.0 * (wt0 > 0.3)).astype(np.float32)
    j = O.consist_out(pr, tg, wt, 150, 0.15, 0.15, pr, mode=0)
    gs = O.consist_out(pr, tg, wt, 150, 0.15, 0.15, pr, mode=1)
    assert np.abs(j - gs).max() <= 1.0 / 255.0  # orderings agree within one grey level at 150 sweeps
    # strong uniform weight: converges to the target
    strong = np.full_like(pr, 2.0)
    conv = O.consist_out(tg, tg, strong, 150, 0.15, 0.15, pr)
    assert np.abs(conv - tg).max() < 2e-2


# ---------------------------------------------------------------- stabilization: against the reference's CUDA kernels
needs_gold = pytest.mark.skipif(not os.path.exists(GOLD_STAB), reason="stab_golden.npz not generated yet")


def near(a, b, tol, budget=0.0):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    return (d > tol).mean() <= budget, f"max {d.max():.3e}, frac>{tol:g}: {(d > tol).mean():.2e}"


@needs_gold
@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_kernels_match_reference_gpu_fixtures(O, tag):
    g = np.load(GOLD_STAB)
    o8, p8 = g[f"{tag}_orig8"], g[f"{tag}_proc8"]
    ff, fb = g[f"{tag}_flowFwd"], g[f"{tag}_flowBwd"]
    H, W = o8.shape[1:3]
    of = [O.rgba8_to_f32x3(x) for x in o8]
    pf = [O.rgba8_to_f32x3(x) for x in p8]
    assert np.array_equal(of[0], g[f"{tag}_to_float0"])
    # nvcc contracts a*b+c into FMAs in the reference binary: allow 1-2 ulp of values in [0,1]
    for got, key in ((O.warp_hwc3(of[0], fb), "warp_prevIn"), (O.warp_hwc3(pf[2], ff), "warp_nextPr"),
                     (O.warp_hwc3(of[0], np.ascontiguousarray(fb[..., :2])), "warp_2ch"),
                     (O.bilinear(pf[1], W // 2, H // 2), "bil_down"),
                     (O.bilinear(g[f"{tag}_bil_down"], W, H), "bil_up"),
                     (O.bilinear(ff[: H // 2, : W // 2].copy(), W, H), "bil_flow")):
        ok, msg = near(got, g[f"{tag}_{key}"], 5e-7 * max(1.0, float(np.abs(g[f'{tag}_{key}']).max())))
        assert ok, f"{key}: {msg}"
    pI, pP = O.warp_hwc3(of[0], fb), O.warp_hwc3(pf[0], fb)
    nI, nP = O.warp_hwc3(of[2], ff), O.warp_hwc3(pf[2], ff)
    lW = O.warp_hwc3(pf[2], fb)
    aI, aP = O.adap_comb(of[1], pf[1], pI, pP, nI, nP, lW, 6800.0)
    ok, msg = near(aI, g[f"{tag}_adapIn"], 5e-6, 2e-3)
    assert ok, f"adapIn: {msg}"
    ok, msg = near(aP, g[f"{tag}_adapPr"], 5e-6, 2e-3)
    assert ok, f"adapPr: {msg}"
    ok, msg = near(O.consist_wt(g[f"{tag}_adapIn"], of[1], 6800.0, 2.0), g[f"{tag}_consWt"], 5e-6, 2e-3)
    assert ok, f"consWt: {msg}"
    for it in (1, 10, 150):
        ref = g[f"{tag}_solve{it}"]
        j = O.consist_out(pf[1], g[f"{tag}_adapPr"], g[f"{tag}_consWt"], it, 0.15, 0.15, pf[1], mode=0)
        d = np.abs(j - ref).max()
        spread = float(g[f"{tag}_solve{it}_spread"])  # the reference's own run-to-run spread (in-place race)
        assert d <= (1.0 / 255.0 if it == 150 else 3.0 / 255.0 + 2.5 * spread), (it, d, spread)
    assert np.array_equal(O.f32x3_to_rgba8(g[f"{tag}_solve150"]), g[f"{tag}_to_char"])
    assert np.array_equal(O.f32x3_to_rgba8(g[f"{tag}_to_char_odd_in"]), g[f"{tag}_to_char_odd"])


@needs_gold
@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("pname", ["default", "slider"])
def test_oracle_sequence_matches_reference_gpu_fixtures(O, tag, pname):
    g = np.load(GOLD_STAB)
    o8, p8 = g[f"{tag}_orig8"], g[f"{tag}_proc8"]
    ff, fb = g[f"{tag}_flowFwd"], g[f"{tag}_flowBwd"]
    of = [O.rgba8_to_f32x3(x) for x in o8]
    pf = [O.rgba8_to_f32x3(x) for x in p8]
    params = None if pname == "default" else dict(numIter=40, gamma=4.0, alpha=3000.0)
    last = pf[2]
    for t in (1, 2, 3):
        co, rgba = O.do_one_step(of[t - 1], of[t], of[t + 1], pf[t - 1], pf[t], pf[t + 1], last, ff, fb, params)
        last = co
        d = np.abs(rgba.astype(np.int32) - g[f"{tag}_{pname}_step{t}_rgba"].astype(np.int32))
        assert d.max() <= 1, f"{tag}/{pname} step {t}: {d.max()}"

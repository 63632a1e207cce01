"""Seeded synthetic inputs shared by the tests, bench.py and the golden generators.

Frames: RGBA8, a smooth base (sum of sinusoids + checkerboard edges) translated by (2.5, -1.25) px per frame
plus per-pixel noise; the "processed" stream is the base with a per-frame gain/offset flicker and noise, so the
adaptive weights span 0..clamp and the solver has real work.  Flow: the analytic translation plus a small
sinusoidal field, forward and backward, HWC with 3 (model layout: u,v,0) or 2 (.flo layout) channels.
"""
from __future__ import annotations

import numpy as np

VX, VY = 2.5, -1.25


def _base(W, H, t, rng_noise=None):
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    xs, ys = x - VX * t, y - VY * t
    img = np.empty((H, W, 3), np.float32)
    for c, (fx, fy, ph) in enumerate(((0.031, 0.017, 0.0), (0.023, 0.029, 1.3), (0.011, 0.037, 2.1))):
        v = 0.5 + 0.2 * np.sin(fx * xs * 2 * np.pi / 4 + ph) + 0.15 * np.cos(fy * ys * 2 * np.pi / 4 + 0.5 * ph) \
            + 0.1 * np.sin(0.05 * (xs + ys) + c)
        chk = (((np.floor(xs / 24) + np.floor(ys / 24)) % 2) * 2 - 1) * 0.08
        img[..., c] = v + chk
    return img


def frames(W: int, H: int, T: int, seed: int = 1234, mismatch: float = 0.0):
    """-> (orig[T,H,W,4] u8, proc[T,H,W,4] u8).  mismatch: fraction of the image replaced by unrelated content
    in the processed stream of odd frames (zero-weight regions for the solver)."""
    rng = np.random.default_rng(seed)
    rng2 = np.random.default_rng(seed + 3087)
    orig = np.empty((T, H, W, 4), np.uint8)
    proc = np.empty((T, H, W, 4), np.uint8)
    for t in range(T):
        b = _base(W, H, t)
        o = b * 255.0 + rng.uniform(-2, 2, b.shape).astype(np.float32)
        g = rng2.uniform(0.9, 1.1)
        off = rng2.uniform(-8, 8)
        # a crude "stylisation": posterise + flicker
        p = (np.round(b * 12) / 12) * 255.0 * g + off + rng2.normal(0, 2.0, b.shape).astype(np.float32)
        if mismatch > 0 and (t % 2 == 1):
            hh = int(H * mismatch)
            p[:hh] = 255.0 - p[:hh]
        orig[t, ..., :3] = np.clip(o, 0, 255).astype(np.uint8)
        proc[t, ..., :3] = np.clip(p, 0, 255).astype(np.uint8)
        orig[t, ..., 3] = 255
        proc[t, ..., 3] = 255
    return orig, proc


def flows(W: int, H: int, channels: int = 3, seed: int = 99):
    """-> (flowFwd, flowBwd) float32 [H,W,channels]: cur->next and cur->prev displacement fields."""
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    wob_x = 0.5 * np.sin(0.02 * x + 0.013 * y + seed)
    wob_y = 0.5 * np.cos(0.017 * x - 0.021 * y + seed)
    fwd = np.zeros((H, W, channels), np.float32)
    bwd = np.zeros((H, W, channels), np.float32)
    fwd[..., 0] = VX + wob_x
    fwd[..., 1] = VY + wob_y
    bwd[..., 0] = -VX - wob_x
    bwd[..., 1] = -VY - wob_y
    return fwd, bwd


def features(N: int, C: int, H: int, W: int, seed: int):
    """PWC-Net-like feature maps: leaky-ReLU(N(0,1)), slope 0.1."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((N, C, H, W)).astype(np.float32)
    return np.where(a > 0, a, 0.1 * a).astype(np.float32)


def op_flow(N: int, H: int, W: int, seed: int, sigma: float = 2.0):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((N, 2, H, W)) * sigma).astype(np.float32)


def f32_images(W: int, H: int, seed: int, k: int):
    """k random float3 images in [0,1] (kernel-level tests)."""
    rng = np.random.default_rng(seed)
    return [rng.random((H, W, 3), dtype=np.float32) for _ in range(k)]


# PWC-Net level shapes (C, H, W) of the BASELINE configs
LIGHT_1080P_CORR = [(196, 9, 15), (128, 18, 30), (96, 36, 60), (64, 72, 120)]
LIGHT_1080P_WARP = [(128, 18, 30), (96, 36, 60), (64, 72, 120)]
DENSE_4K_CORR = [(196, 34, 60), (128, 68, 120), (96, 136, 240), (64, 272, 480), (32, 544, 960)]
DENSE_4K_WARP = [(128, 68, 120), (96, 136, 240), (64, 272, 480), (32, 544, 960)]


def op_flow_smooth(N: int, H: int, W: int, seed: int, amp: float = 2.0, noise: float = 0.25):
    """a spatially smooth flow field (what a PWC-Net decoder level actually feeds custom::Warp): a global
    translation + low-frequency sinusoids of amplitude `amp` px + N(0, noise) per pixel."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    f = np.empty((N, 2, H, W), np.float32)
    for n in range(N):
        tx, ty = rng.uniform(-amp, amp, 2)
        f[n, 0] = tx + amp * np.sin(2 * np.pi * (x / max(W, 1) + 0.5 * y / max(H, 1)) + n)
        f[n, 1] = ty + amp * np.cos(2 * np.pi * (y / max(H, 1) - 0.3 * x / max(W, 1)) + n)
    f += rng.normal(0, noise, f.shape).astype(np.float32)
    return f


def u8_distance(a, b):
    """per-byte distance of two 8-bit frames ON THE CIRCLE mod 256.

    The reference's float -> 8-bit conversion is floor(|v| * 255) truncated to a byte WITHOUT a clamp
    (gpuimage.cu:54-67): v = 1.0039 gives 256 -> 0.  Two fp32 images that agree to 1e-3 (the run-to-run spread of
    the reference's own racy in-place solver) can therefore differ by "255 grey levels" at a pixel whose value sits
    on a multiple of 256/255; on the circle that is the 1-level step it really is.  The north_star's <= 1/255 gate on
    8-bit frames is applied with this distance whenever the other side is the reference's GPU code."""
    import numpy as np

    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return np.minimum(d, 256 - d)

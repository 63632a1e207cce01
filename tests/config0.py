"""Loader for the BASELINE configs[0] fixture (tests/golden/config0_720p.npz, made by
tests/golden/make_config0_fixture.py from the reference's bundled videos/input.mp4) and a plain restatement of
the reference's stream loop for the tests: preloadProcessedFrames / doOneStep / outputFinalFrames
(videostabilizer.cpp:136-164,167-265 with k = 1, batchSize = 1).

Nothing here reads /root/reference: the committed fixture carries the decoded frames and the flows.
"""
from __future__ import annotations

import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "config0_720p.npz")
REFGPU = os.path.join(HERE, "golden", "config0_refgpu.npz")

_cache = {}


def _processed(orig_rgb: np.ndarray, t: int) -> np.ndarray:
    """a stand-in for the missing videos/processed.mp4: posterised copy + per-frame gain/offset flicker + noise,
    seeded per frame; u8 RGB -> u8 RGB"""
    rng = np.random.default_rng(4321 + 17 * t)
    g = rng.uniform(0.92, 1.08)
    off = rng.uniform(-7.0, 7.0)
    b = orig_rgb.astype(np.float32) / 255.0
    p = (np.round(b * 10.0) / 10.0) * 255.0 * g + off + rng.normal(0.0, 2.0, b.shape).astype(np.float32)
    return np.clip(p, 0, 255).astype(np.uint8)


def load():
    """-> dict(W, H, T, orig8 [T,H,W,4] u8, proc8 [T,H,W,4] u8, flows [(fwd, bwd)] * (T-2) float32 [H,W,3])

    flows[i] belongs to current frame i+1; they are stored at 1/flow_down resolution and up-sampled here with the
    oracle's get_bilinear restatement, values not rescaled -- the FLOWDOWNSCALE path of flowmodel.cpp:156-165."""
    if "c" in _cache:
        return _cache["c"]
    import cv2

    from oracle import oracle as O

    z = np.load(FIXTURE)
    T, W, H = int(z["T"]), int(z["W"]), int(z["H"])
    orig8 = np.empty((T, H, W, 4), np.uint8)
    proc8 = np.empty((T, H, W, 4), np.uint8)
    for t in range(T):
        bgr = cv2.imdecode(z["jpeg_%02d" % t], cv2.IMREAD_COLOR)
        assert bgr is not None and bgr.shape == (H, W, 3)
        rgb = bgr[..., ::-1]
        orig8[t, ..., :3] = rgb
        orig8[t, ..., 3] = 255
        proc8[t, ..., :3] = _processed(rgb, t)
        proc8[t, ..., 3] = 255
    flows = []
    for i in range(T - 2):
        pair = []
        for key in ("flow_fwd", "flow_bwd"):
            lo = z[key][i].astype(np.float32) / np.float32(z["flow_q"])
            lo3 = np.zeros(lo.shape[:2] + (3,), np.float32)   # model layout (u, v, 0), videostabilizer.cpp:50-51
            lo3[..., :2] = lo
            pair.append(O.bilinear(lo3, W, H))
        flows.append(tuple(pair))
    c = dict(W=W, H=H, T=T, orig8=orig8, proc8=proc8, flows=flows)
    _cache["c"] = c
    return c


def reference_stream_loop(T, proc8, step):
    """The reference's output sequence for a stream of T frames (k = 1, batchSize = 1).

    `step(t)` must stabilise current frame t (window t-1, t, t+1) and return its RGBA8 frame.  Returns a dict
    {frame index: RGBA8 frame} exactly as outputFrame() is called:
      * preloadProcessedFrames (:136-153): frames j <= k of the processed stream are written unchanged (through
        gpuToImage: alpha becomes 1) -- frame 1 is written here AND again by doOneStep(1), the later call wins;
      * doOneStep(t) for t = 1 .. T-2 (:167-265); after the last one loadFrame fails and
      * outputFinalFrames (:155-164) writes processed frame t+1 unchanged.
    """
    out = {}

    def passthrough(f):
        g = f.copy()
        g[..., 3] = 1      # float round trip u8/255 -> floor(v*255) is the identity on bytes; alpha = 1
        return g

    out[0] = passthrough(proc8[0])
    out[1] = passthrough(proc8[1])
    for t in range(1, T - 1):
        out[t] = step(t)
    out[T - 1] = passthrough(proc8[T - 1])
    return out

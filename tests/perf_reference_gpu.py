#!/usr/bin/env python
"""Same-box timing of the REFERENCE's own GPU kernels next to this library.

Test infrastructure, run by hand on a GPU box:

    python tests/perf_reference_gpu.py [--out gpurun_out/ref_gpu_same_box.json] [--sizes 1080p,4k]

The reference arm is oracle/_ref/libvsc_ref_gpu.so: the reference's flowconsistency.cu / gpuimage.cu /
correlation_cuda.cu / warp_cuda.cu compiled UNMODIFIED for sm_100a (oracle/Makefile) and driven in the call
sequence of VideoStabilizer::doOneStep (videostabilizer.cpp:167-265), with its own temporaries, cudaMallocs and
synchronisations, exactly as the reference runs them.  Both arms get device-resident fp32 frames and flow; both
produce the stabilized fp32 frame and the RGBA8 frame (the reference arm copies the RGBA8 frame to the host as its
gpuToImage does; ours converts on the device -- the D2H copy is measured separately by bench.py's e2e arm).
Times are host-clock between full device synchronisations (the reference synchronises inside every call, so CUDA
events would measure the same thing).  Nothing here is a bench.py number; the output is a profile note.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "video-stream-consistency_b200"), ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

SIZES = {"720p": (1280, 720), "1080p": (1920, 1080), "4k": (3840, 2160)}
LIGHT_CORR = [(196, 9, 15), (128, 18, 30), (96, 36, 60), (64, 72, 120)]
LIGHT_WARP = [(128, 18, 30), (96, 36, 60), (64, 72, 120)]
DENSE_CORR = [(196, 34, 60), (128, 68, 120), (96, 136, 240), (64, 272, 480), (32, 544, 960)]
DENSE_WARP = [(128, 68, 120), (96, 136, 240), (64, 272, 480), (32, 544, 960)]


def timed(fn, sync, warm, reps):
    for _ in range(warm):
        fn()
    sync()
    ts = []
    for _ in range(reps):
        sync()
        t0 = time.perf_counter()
        fn()
        sync()
        ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts), min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_gpu_same_box.json"))
    ap.add_argument("--sizes", default="1080p,4k")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()

    import torch

    import synth
    import vsc_b200 as V
    from oracle import oracle as O

    assert torch.cuda.is_available() and O.ref_gpu_available(), "needs a GPU and oracle/_ref/libvsc_ref_gpu.so"
    dev = torch.device("cuda", 0)
    sync = torch.cuda.synchronize
    L, R = V.lib(), O.ref_gpu()
    res = {"device": torch.cuda.get_device_name(0), "timing": "host clock, median of %d (min in brackets)" % args.reps,
           "stabilization": {}, "ops": []}

    def dp(t):
        return C.c_void_p(t.data_ptr())

    # ------------------------------------------------------------------ stabilization step (doOneStep)
    for name in args.sizes.split(","):
        W, H = SIZES[name]
        o8, p8 = synth.frames(W, H, 3, seed=5)
        of = [V.image_to_gpu(torch.from_numpy(x).to(dev)) for x in o8]
        pf = [V.image_to_gpu(torch.from_numpy(x).to(dev)) for x in p8]
        ff, fb = (torch.from_numpy(x).to(dev) for x in synth.flows(W, H, 3))
        hp = V.HyperParams()

        ref = O.RefGpuStepper(W, H, 3, hp.pyramidLevels)
        last_r = pf[2].clone()
        out_r = torch.zeros((H, W, 3), device=dev)
        rgba_r = np.zeros((H, W, 4), np.uint8)

        def ref_step():
            rc = R.vsc_ref_gpu_step(ref.h, dp(of[0]), dp(of[1]), dp(of[2]), dp(pf[0]), dp(pf[1]), dp(pf[2]), dp(last_r),
                                    dp(ff), dp(fb), C.c_float(hp.alpha), C.c_float(hp.beta), C.c_float(hp.gamma),
                                    int(hp.numIter), C.c_float(hp.stepSize), C.c_float(hp.momFac), dp(out_r),
                                    rgba_r.ctypes.data_as(C.c_void_p))
            assert rc == 0

        ws = torch.empty(int(L.vsc_frame_stabilize_workspace_bytes(W, H, hp.pyramidLevels)), device=dev,
                         dtype=torch.uint8)
        state = {"last": pf[2].clone(), "cons": torch.empty((H, W, 3), device=dev)}
        out8 = torch.empty((H, W, 4), device=dev, dtype=torch.uint8)

        def our_step():
            V.frame_stabilize(of[0], of[1], of[2], pf[0], pf[1], pf[2], state["last"], ff, fb, hp, out=state["cons"],
                              workspace=ws)
            V.check(L.vsc_f32x3_to_rgba8(dp(state["cons"]), dp(out8), W, H,
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            state["last"], state["cons"] = state["cons"], state["last"]

        r_med, r_min = timed(ref_step, sync, 2, args.reps)
        o_med, o_min = timed(our_step, sync, 3, args.reps)
        # parity of the two arms on this very input, one frame from identical state (8-bit levels)
        last_r.copy_(pf[2])
        state["last"].copy_(pf[2])
        ref_step()
        our_step()
        sync()
        d = np.abs(out8.cpu().numpy()[..., :3].astype(np.int32) - rgba_r[..., :3].astype(np.int32))
        res["stabilization"][name] = {
            "reference_gpu_ms_per_frame": r_med, "reference_gpu_ms_min": r_min, "ours_ms_per_frame": o_med,
            "ours_ms_min": o_min, "speedup": r_med / o_med, "max_abs_diff_8bit_levels": int(d.max()),
            "note": "reference = its kernels + its cudaMalloc/sync/D2D copies + D2H of the RGBA8 frame; "
                    "ours = vsc_frame_stabilize + vsc_f32x3_to_rgba8, device resident"}
        print(name, json.dumps(res["stabilization"][name]), flush=True)
        ref.close()
        del of, pf, ff, fb, ws, out8, last_r, out_r, state
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------ custom ops, per level shape
    g = torch.Generator(device="cpu").manual_seed(11)
    st = C.c_void_p(0)
    for kind, shapes in (("Correlation", LIGHT_CORR + DENSE_CORR), ("Warp", LIGHT_WARP + DENSE_WARP)):
        for (Cc, H, W) in shapes:
            a = torch.nn.functional.leaky_relu(torch.randn((1, Cc, H, W), generator=g), 0.1).to(dev)
            if kind == "Correlation":
                b = torch.nn.functional.leaky_relu(torch.randn((1, Cc, H, W), generator=g), 0.1).to(dev)
                out_r = torch.zeros((1, 9, 9, H, W), device=dev)
                out_o = torch.zeros_like(out_r)
                nbytes = 4 * H * W * (2 * Cc + 81)

                def ref_op():
                    assert R.vsc_ref_gpu_correlation(dp(a), dp(b), dp(out_r), C.c_size_t(out_r.numel() * 4),
                                                     C.c_int64(1), C.c_int64(Cc), C.c_int64(H), C.c_int64(W),
                                                     C.c_int64(4), C.c_int64(0), st) == 0

                def our_op():
                    V.check(L.vsc_correlation_f32(dp(a), dp(b), dp(out_o), 1, Cc, H, W, 4, 0, st))
            else:
                b = (2.0 * torch.randn((1, 2, H, W), generator=g)).to(dev)
                out_r = torch.zeros_like(a)
                out_o = torch.zeros_like(a)
                nbytes = 4 * H * W * (2 * Cc + 2)

                def ref_op():
                    assert R.vsc_ref_gpu_warp(dp(a), dp(b), dp(out_r), C.c_size_t(out_r.numel() * 4), C.c_int64(1),
                                              C.c_int64(Cc), C.c_int64(H), C.c_int64(W), st) == 0

                def our_op():
                    V.check(L.vsc_warp_nchw_f32(dp(a), dp(b), dp(out_o), 1, Cc, H, W, st))

            r_med, r_min = timed(ref_op, sync, 2, args.reps)
            o_med, o_min = timed(our_op, sync, 3, args.reps)
            diff = float((out_r - out_o).abs().max())
            scale = float(out_r.abs().max())
            row = {"op": kind, "C": Cc, "H": H, "W": W, "reference_gpu_us": r_med * 1e3, "ours_us": o_med * 1e3,
                   "ours_us_min": o_min * 1e3, "speedup": r_med / o_med, "algorithmic_MB": nbytes / 1e6,
                   "max_abs_diff": diff, "max_abs_ref": scale}
            res["ops"].append(row)
            print(json.dumps(row), flush=True)

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

"""A/B timing of the fused stage-A kernel selections (vsc_set_stage_a_mode) inside vsc_frame_stabilize, and of the
channel split of the linear custom::Warp kernel (vsc_set_warp_mode | k << 4).

Stage A is timed as vsc_frame_stabilize with numIter = 0 (stage A + momentum memset + the x2 up-scale; no sweeps):
the differences between modes are differences of the stage-A kernel alone; the absolute kernel times come from
the ncu launch list of the same script (`--ncu` runs every mode twice and nothing else).  All modes must give
bit-identical frames (checked with 6 sweeps).  CUDA events, L2 flushed before every timed call, median of 15.

    python profiles/time_stage_a.py [--ncu] > gpurun_out/time_stage_a.txt
"""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "video-stream-consistency_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import synth  # noqa: E402
import vsc_b200 as V  # noqa: E402

NCU = "--ncu" in sys.argv
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
L = V.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=15):
    ts = []
    for _ in range(3):
        fn()
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(300000)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


def mode_name(m):
    kind = m & 0xF
    rows = (m >> 12) & 0xFF or (1 << ((m >> 4) & 0xF) if (m >> 4) & 0xF else 8)
    s = {0: "default", 1: "row-per-CTA", 2: "walk", 3: "walk+flow-prefetch", 4: "walk+pipelined"}[kind]
    if kind >= 2:
        s += f" rows={rows}" + (" bs=128" if m & 0x100 else "") + ("", " 5cta", " 6cta")[(m >> 9) & 3]
    return s


FULL = "--full" in sys.argv   # the first sweep (profiles/r1_stage_a_walk_events.txt); default: the shortlist
MODES = [1]
if FULL:
    for kind in (2, 3, 4):
        for lg in (1, 2, 3, 4, 5):
            MODES.append(kind | (lg << 4))
    MODES += [3 | (2 << 4) | 0x100, 3 | (3 << 4) | 0x100, 4 | (2 << 4) | 0x100, 4 | (3 << 4) | 0x100,
              4 | (4 << 4) | 0x100]
else:   # (the second sweep, r1_stage_a_walk_events2.txt, also had register-capped builds; those were removed)
    for bs in (0, 0x100):
        for lg in (2, 3, 4):
            MODES.append(3 | (lg << 4) | bs)
if "--rows" in sys.argv:   # rows per CTA as a number, 128-thread CTAs
    MODES = [1] + [3 | 0x100 | (r << 12) for r in (3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16, 20, 24)]
if NCU:
    MODES = [1] + [3 | 0x100 | (r << 12) for r in (4, 5, 6, 7, 8, 9, 10, 12, 16)]

SIZES = ((1280, 720), (1920, 1080), (3840, 2160)) if "--rows" in sys.argv else ((1920, 1080), (3840, 2160))
for (W, H) in (() if "--warp" in sys.argv else SIZES):
    o8, p8 = synth.frames(W, H, 3)
    of = [V.image_to_gpu(torch.from_numpy(o8[t]).to(dev)) for t in range(3)]
    pf = [V.image_to_gpu(torch.from_numpy(p8[t]).to(dev)) for t in range(3)]
    ff, fb = (torch.from_numpy(a).to(dev) for a in synth.flows(W, H, 3))
    last = pf[0].clone()
    ws = torch.empty(int(L.vsc_frame_stabilize_workspace_bytes(W, H, 2)), device=dev, dtype=torch.uint8)
    out = torch.empty_like(of[0])

    def run(hp):
        V.frame_stabilize(of[0], of[1], of[2], pf[0], pf[1], pf[2], last, ff, fb, hp, out=out, workspace=ws)

    if NCU:
        hp0 = V.HyperParams(numIter=0)
        for m in MODES:
            assert L.vsc_set_stage_a_mode(m) == 0
            run(hp0)
            run(hp0)
        L.vsc_set_stage_a_mode(0)
        torch.cuda.synchronize()
        continue

    # bit-identity of every selection (6 sweeps so that the coefficients are consumed)
    hp6 = V.HyperParams(numIter=6)
    L.vsc_set_stage_a_mode(1)
    run(hp6)
    ref = out.clone()
    bad = []
    for m in MODES[1:]:
        assert L.vsc_set_stage_a_mode(m) == 0
        out.zero_()
        run(hp6)
        if not torch.equal(out, ref):
            bad.append(hex(m))
    print(f"{W}x{H}: bit-identical to mode 1: {'ALL' if not bad else 'NOT ' + ','.join(bad)}", flush=True)

    hp0 = V.HyperParams(numIter=0)
    alg = (108 + 24 + 9) * W * H
    base = None
    for m in MODES:
        assert L.vsc_set_stage_a_mode(m) == 0
        t = timed(lambda: run(hp0))
        if base is None:
            base = t
        print(f"  {W}x{H} mode {m:#05x} {mode_name(m):36s} {t:8.1f} us  ({t - base:+7.1f} vs mode 1)", flush=True)
    L.vsc_set_stage_a_mode(0)
    del of, pf, ff, fb, last, ws, out
    torch.cuda.empty_cache()

# (the walking / shuffle kernels, modes 4 and 5 of the sweeps recorded in r1_stage_a_walk_events2.txt and
# r1_warp_variants_*.txt, were removed after these measurements)
WARP_MODES = [1, 3, 2, 1 | (1 << 4), 1 | (2 << 4)]
g = torch.Generator(device=dev).manual_seed(0)
if "--rows" in sys.argv:
    sys.exit(0)
if NCU:
    for (C, H, W) in synth.DENSE_4K_WARP[1:]:
        a = torch.randn((1, C, H, W), device=dev, generator=g)
        fl = torch.from_numpy(synth.op_flow_smooth(1, H, W, 5)).to(dev)
        o = torch.empty_like(a)
        for m in WARP_MODES:
            assert L.vsc_set_warp_mode(m) == 0
            V.warp(a, fl, out=o)
        L.vsc_set_warp_mode(0)
    torch.cuda.synchronize()
else:
    print("custom::Warp (us, cold L2), modes " + ", ".join(hex(m) for m in WARP_MODES) + " | torch copy of the same "
          "tensor (read + write = the op's algorithmic bytes minus the flow)")
    for (C, H, W) in synth.LIGHT_1080P_WARP + synth.DENSE_4K_WARP:
        a = torch.randn((1, C, H, W), device=dev, generator=g)
        fl = torch.from_numpy(synth.op_flow_smooth(1, H, W, 5)).to(dev)
        o = torch.empty_like(a)
        row = []
        for m in WARP_MODES:
            assert L.vsc_set_warp_mode(m) == 0
            row.append(timed(lambda: V.warp(a, fl, out=o)))
        L.vsc_set_warp_mode(0)
        tcopy = timed(lambda: o.copy_(a))
        mb = 4 * H * W * (2 * C + 2) / 1e6
        print(f"  {C:4d}x{H:4d}x{W:4d}  " + "  ".join(f"{t:6.1f}" for t in row) + f"  | copy {tcopy:6.1f}   {mb:7.1f} MB  best "
              f"{mb / min(row) * 1e3:6.0f} GB/s", flush=True)

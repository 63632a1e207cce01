"""Event timing of custom::Warp per dense-4K level shape and kernel selection (vsc_set_warp_mode), L2 flushed before
every timed launch; a device-to-device copy of the same tensor beside it.  Flow = the bench's smooth field
(tests/synth.py: op_flow_smooth) and an i.i.d. sigma = 2 px field.  CUDA events, median of 15.  Not a bench.py number.

    python profiles/time_warp_r2.py > gpurun_out/time_warp_r2.txt
"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "video-stream-consistency_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import synth  # noqa: E402
import vsc_b200 as V  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
g = torch.Generator(device=dev).manual_seed(0)
L = V.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6547.0
SHAPES = [(32, 544, 960), (64, 272, 480), (96, 136, 240), (128, 68, 120), (64, 72, 120)]
MODES = [(1, "linear"), (2, "tiled"), (4, "staged"), (4 | (1 << 4), "staged/1"), (4 | (2 << 4), "staged/2"),
         (4 | (4 << 4), "staged/4")]


def timed(fn, reps=15):
    ts = []
    for _ in range(3):
        fn()
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(300000)   # keeps the GPU busy while the host enqueues: the interval is pure kernel time
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


print(f"custom::Warp, us per launch (fraction of the measured HBM peak {PEAK:.0f} GB/s on 4*H*W*(2C+2) bytes)")
for (C, H, W) in SHAPES:
    x = torch.randn((1, C, H, W), device=dev, generator=g)
    y = torch.empty_like(x)
    nbytes = 4 * H * W * (2 * C + 2)
    tc = timed(lambda: y.copy_(x))
    print(f"C{C} {H}x{W}: {nbytes / 1e6:.1f} MB; device copy of the tensor {tc:.1f} us ({4 * H * W * 2 * C / tc / 1e3 / PEAK:.2f})")
    flows = (("bench-smooth", torch.from_numpy(synth.op_flow_smooth(1, H, W, 3)).to(dev)),
             ("random-s2", 2.0 * torch.randn((1, 2, H, W), device=dev, generator=g)))
    # the bench's method for the default selection: 20 launches back to back over two tensor sets
    sets = [(torch.randn((1, C, H, W), device=dev, generator=g), flows[0][1], torch.empty((1, C, H, W), device=dev))
            for _ in range(2)]
    for s_ in sets:
        V.warp(s_[0], s_[1], out=s_[2])
    torch.cuda.synchronize()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    for i in range(20):
        s_ = sets[i & 1]
        V.warp(s_[0], s_[1], out=s_[2])
    eb.record()
    torch.cuda.synchronize()
    tb = ea.elapsed_time(eb) / 20 * 1e3
    print(f"   default selection, back to back on the bench's smooth flow: {tb:.1f} us ({nbytes / tb / 1e3 / PEAK:.3f})")
    for name, fl in flows:
        row = []
        ref = None
        for mode, label in MODES:
            V.check(L.vsc_set_warp_mode(mode))
            out = V.warp(x, fl)
            if ref is None:
                ref = out
            ok = torch.equal(out.view(torch.int32), ref.view(torch.int32))
            t = timed(lambda: V.warp(x, fl, out=y))
            row.append(f"{label}{'' if ok else '!MISMATCH'} {t:6.1f} ({nbytes / t / 1e3 / PEAK:.2f})")
        L.vsc_set_warp_mode(0)
        print(f"   {name:12s} " + "  ".join(row), flush=True)

"""A/B timing of the custom-op kernels per level shape and kernel selection (vsc_set_correlation_mode /
vsc_set_warp_mode), with an L2 flush before every timed launch (inputs DRAM-cold, as after a convolution that
wrote > L2 bytes) and without (inputs L2-resident).  CUDA events, median of 15.  Not a bench.py number.

    python profiles/time_ops.py > gpurun_out/time_ops.txt
"""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "video-stream-consistency_b200"), ROOT):
    sys.path.insert(0, p)

import torch  # noqa: E402

import vsc_b200 as V  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
g = torch.Generator(device=dev).manual_seed(0)
L = V.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

LIGHT_CORR = [(196, 9, 15), (128, 18, 30), (96, 36, 60), (64, 72, 120)]
DENSE_CORR = [(196, 34, 60), (128, 68, 120), (96, 136, 240), (64, 272, 480), (32, 544, 960)]
LIGHT_WARP = [(128, 18, 30), (96, 36, 60), (64, 72, 120)]
DENSE_WARP = [(128, 68, 120), (96, 136, 240), (64, 272, 480), (32, 544, 960)]


def timed(fn, cold, reps=15):
    ts = []
    for _ in range(3):
        fn()
    for _ in range(reps):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(300000)   # keeps the GPU busy while the host enqueues: the interval is pure kernel time
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


def smooth_flow(H, W, amp=2.0):
    y, x = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float32),
                          torch.arange(W, device=dev, dtype=torch.float32), indexing="ij")
    f = torch.stack([amp * torch.sin(0.05 * x + 0.03 * y) + 1.3, amp * torch.cos(0.04 * x - 0.02 * y) - 0.7])
    return f[None].contiguous()


print("Correlation (us): shape, then per mode cold/warm; algorithmic GB/s for the best cold time")
for (C, H, W) in LIGHT_CORR + DENSE_CORR:
    a = torch.randn((1, C, H, W), device=dev, generator=g)
    b = torch.randn((1, C, H, W), device=dev, generator=g)
    row = []
    for mode in (0, 1, 2, 3, 4):
        if mode in (2, 3) and W % 4:
            continue
        if mode == 4 and H * W > 140000:
            continue
        V.check(L.vsc_set_correlation_mode(mode))
        f = lambda: V.correlation(a, b)
        row.append((mode, timed(f, True), timed(f, False)))
    L.vsc_set_correlation_mode(0)
    best = min(r[1] for r in row)
    nbytes = 4 * H * W * (2 * C + 81)
    print(f"  C{C:<4d}{H:>4d}x{W:<4d} " + "  ".join(f"m{m}: {c:7.1f}/{w:7.1f}" for m, c, w in row)
          + f"   {nbytes / best / 1e3:7.0f} GB/s", flush=True)

print("Warp (us): shape, mode 1 (linear) and 2 (tiled) cold/warm, smooth and random (sigma 2 px) flow")
for (C, H, W) in LIGHT_WARP + DENSE_WARP:
    x = torch.randn((1, C, H, W), device=dev, generator=g)
    for name, fl in (("smooth", smooth_flow(H, W)), ("random", 2.0 * torch.randn((1, 2, H, W), device=dev, generator=g))):
        row = []
        for mode in (1, 2):
            V.check(L.vsc_set_warp_mode(mode))
            f = lambda: V.warp(x, fl)
            row.append((mode, timed(f, True), timed(f, False)))
        L.vsc_set_warp_mode(0)
        best = min(r[1] for r in row)
        nbytes = 4 * H * W * (2 * C + 2)
        print(f"  C{C:<4d}{H:>4d}x{W:<4d} {name:6s} " + "  ".join(f"m{m}: {c:7.1f}/{w:7.1f}" for m, c, w in row)
              + f"   {nbytes / best / 1e3:7.0f} GB/s", flush=True)

"""Event timing of custom::Correlation on the small (coarse PWC-Net) level shapes per kernel selection
(vsc_set_correlation_mode: 4 = channel-split rows kernel, 7 = its quad form, 0 = auto; earlier runs: 6 / 5 = shared-row tiles): back to back over two
tensor sets, and isolated with an L2 flush before every launch.  Not a bench.py number.

    python profiles/time_corr_small.py
"""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "video-stream-consistency_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import vsc_b200 as V  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
g = torch.Generator(device=dev).manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
L = V.lib()
for (C, H, W) in ((64, 72, 120), (96, 36, 60), (128, 18, 30), (196, 9, 15), (128, 68, 120), (196, 34, 60), (64, 136, 88), (96, 136, 240), (64, 136, 240), (64, 272, 480), (32, 544, 960)):
    sets = [(torch.randn((1, C, H, W), device=dev, generator=g), torch.randn((1, C, H, W), device=dev, generator=g),
             torch.empty((1, 9, 9, H, W), device=dev)) for _ in range(2)]
    row = []
    for mode in (0, 4, 7, 6):
        if mode in (5, 6, 7) and W % 4:
            continue
        V.check(L.vsc_set_correlation_mode(mode))
        for s in sets:
            V.correlation(s[0], s[1], out=s[2])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(40):
            s = sets[i & 1]
            V.correlation(s[0], s[1], out=s[2])
        b.record()
        torch.cuda.synchronize()
        t_b2b = a.elapsed_time(b) / 40 * 1e3
        ts = []
        for _ in range(11):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(300000)
            e0.record()
            V.correlation(sets[0][0], sets[0][1], out=sets[0][2])
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        row.append(f"m{mode}: {t_b2b:5.1f} b2b {statistics.median(ts):5.1f} iso")
    L.vsc_set_correlation_mode(0)
    print(f"C{C:<3d} {H:3d}x{W:<3d} " + " | ".join(row), flush=True)

"""ncu driver: every Correlation level shape up to 136x240 on each kernel selection (2 launches each), and the
Warp shapes on both kernels -- for `ncu --metrics gpu__time_duration.sum` launch lists (cold-cache per-launch times;
CUDA events are too coarse for 5-30 us kernels).

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python profiles/prof_ops_small.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "video-stream-consistency_b200"), ROOT):
    sys.path.insert(0, p)

import torch  # noqa: E402

import vsc_b200 as V  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
L = V.lib()
for (C, H, W) in [(196, 9, 15), (128, 18, 30), (96, 36, 60), (64, 72, 120), (196, 34, 60), (128, 68, 120),
                  (96, 136, 240)]:
    a = torch.randn((1, C, H, W), device=dev, generator=g)
    b = torch.randn((1, C, H, W), device=dev, generator=g)
    for mode in (4, 2, 3):
        V.check(L.vsc_set_correlation_mode(mode))
        for _ in range(2):
            V.correlation(a, b)
    L.vsc_set_correlation_mode(0)
torch.cuda.synchronize()

"""Event timing of custom::Correlation per dense-4K level shape for the library named by VSC_B200_LIB (A/B builds of the
64x8 TMA kernel), back to back over two tensor sets (the bench's method) and isolated with an L2 flush before every
launch.  Not a bench.py number.

    python profiles/time_corr_r2.py [label]
"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "video-stream-consistency_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import vsc_b200 as V  # noqa: E402

label = sys.argv[1] if len(sys.argv) > 1 else "default"
V.check(V.lib().vsc_set_correlation_mode(int(os.environ.get("VSC_CORR_MODE", "0"))))
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
g = torch.Generator(device=dev).manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6547.0
row = []
for (C, H, W) in ((32, 544, 960), (64, 272, 480), (96, 136, 240), (128, 68, 120)):
    sets = [(torch.randn((1, C, H, W), device=dev, generator=g), torch.randn((1, C, H, W), device=dev, generator=g),
             torch.empty((1, 9, 9, H, W), device=dev)) for _ in range(2)]
    for s in sets:
        V.correlation(s[0], s[1], out=s[2])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(20):
        s = sets[i & 1]
        V.correlation(s[0], s[1], out=s[2])
    b.record()
    torch.cuda.synchronize()
    t_b2b = a.elapsed_time(b) / 20 * 1e3
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
    t_iso = statistics.median(ts)
    nb = 4 * H * W * (2 * C + 81)
    row.append(f"C{C} {H}x{W}: {t_b2b:6.1f} us b2b ({nb / t_b2b / 1e3 / PEAK:.2f}) {t_iso:6.1f} isolated ({nb / t_iso / 1e3 / PEAK:.2f})")
print(f"{label:8s} " + " | ".join(row), flush=True)

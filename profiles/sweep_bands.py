"""Time one blocked solver pass (T = 8 sweeps) for every band geometry the launcher knows, at the image sizes
of the benchmark configs -- calibration data for the cost model in stab_solver_stream.cu.

    python profiles/sweep_bands.py > gpurun_out/sweep_bands.txt

Per size and geometry: (t(numIter=8*(n+m)) - t(numIter=8*n)) / m with CUDA events, which cancels the solver's
set-up kernel.  Not a bench.py number.
"""
import os
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
NAMES = {0: "auto", 1: "512", 2: "448", 3: "384", 4: "256"}
T = int(sys.argv[1]) if len(sys.argv) > 1 else 8


def time_solve(pr, tg, wt, out, iters, reps=5):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        V.get_consist_out(pr, tg, wt, iters, 0.15, 0.15, out)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for (W, H) in ((3840, 2160), (1920, 1080), (1280, 720), (960, 540), (640, 360)):
    pr = torch.rand((H, W, 3), device=dev, generator=g)
    tg = torch.rand((H, W, 3), device=dev, generator=g)
    wt = torch.rand((H, W, 3), device=dev, generator=g) * 2
    out = pr.clone()
    for tmain in (8, 10):
        row = []
        for k in range(0, 5):
            if tmain > 8 and k in (1, 2):
                continue
            V.check(L.vsc_set_solver_mode(2 | (k << 8) | (((tmain - 6) // 2) << 12)))
            time_solve(pr, tg, wt, out, tmain * 4, reps=2)
            a = time_solve(pr, tg, wt, out, tmain * 4)
            b = time_solve(pr, tg, wt, out, tmain * 20)
            row.append((NAMES[k], (b - a) / 16 * 1e3))
        L.vsc_set_solver_mode(0)
        print(f"{W}x{H} T={tmain:2d}: " + "  ".join(f"{n} {us:7.2f} us/pass {us / tmain:6.2f} us/sweep" for n, us in row),
              flush=True)

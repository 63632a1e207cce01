"""Round-2 A/B timing of the blocked solver pass for the library named by VSC_B200_LIB (default build, or a
profiles/build_variant.py build): main-pass depth T, band width, and the number of rows by which the first / last
row chunk is shortened (vsc_set_solver_mode bits 16.. / 23..).

    python profiles/sweep_solver_r2.py [label] >> gpurun_out/sweep_solver_r2.txt

Per configuration: (t(20T sweeps) - t(4T sweeps)) / 16 with CUDA events = time of one blocked pass; every configuration
is first checked bit for bit against the unblocked sweeps.  Not a bench.py number."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "video-stream-consistency_b200"), ROOT):
    sys.path.insert(0, p)

import torch  # noqa: E402

import vsc_b200 as V  # noqa: E402

label = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(os.environ.get("VSC_B200_LIB", "default"))
EDGES = [(0, 0), (8, 4), (12, 8), (16, 10), (24, 16)]
if len(sys.argv) > 2:
    EDGES = [tuple(int(v) for v in e.split(",")) for e in sys.argv[2].split(":")]
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
g = torch.Generator(device=dev).manual_seed(0)
L = V.lib()
NAMES = {0: "auto", 1: "512", 2: "448", 3: "384", 4: "256"}


def time_solve(pr, tg, wt, out, iters, reps=7):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        V.get_consist_out(pr, tg, wt, iters, 0.15, 0.15, out)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for (W, H) in ((1920, 1080), (960, 540), (3840, 2160), (1280, 720)):
    pr = torch.rand((H, W, 3), device=dev, generator=g)
    tg = torch.rand((H, W, 3), device=dev, generator=g)
    wt = torch.rand((H, W, 3), device=dev, generator=g) * 2
    out = pr.clone()
    refs = {}
    for tmain in (8, 10):
        V.check(L.vsc_set_solver_mode(1))
        refs[tmain] = V.get_consist_out(pr, tg, wt, 2 * tmain + 3, 0.15, 0.15, pr.clone())
        for (et, eb) in EDGES:
            row = []
            for unrolled in (0, 1):
                for k in range(0, 5):
                    if (tmain > 8 and k in (1, 2)) or (unrolled and (k != 0 or (et, eb) != EDGES[0])):
                        continue
                    mode = 2 | (0x8000 if unrolled else 0) | (k << 8) | (((tmain - 6) // 2) << 12) | ((et + 1) << 16) | ((eb + 1) << 22)
                    V.check(L.vsc_set_solver_mode(mode))
                    got = V.get_consist_out(pr, tg, wt, 2 * tmain + 3, 0.15, 0.15, pr.clone())
                    ok = torch.equal(got, refs[tmain])
                    time_solve(pr, tg, wt, out, tmain * 4, reps=2)
                    a = time_solve(pr, tg, wt, out, tmain * 4)
                    b = time_solve(pr, tg, wt, out, tmain * 20)
                    row.append((("UNROLLED-" if unrolled else "") + NAMES[k] + ("" if ok else "!MISMATCH"), (b - a) / 16 * 1e3))
            L.vsc_set_solver_mode(0)
            print(f"{label:8s} {W}x{H} T={tmain:2d} edge-{et:02d}/{eb:02d}: "
                  + "  ".join(f"{n} {us:7.2f} ({us / tmain:5.2f}/sw)" for n, us in row), flush=True)

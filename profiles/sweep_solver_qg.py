"""A/B timing of the blocked solver pass with the exchange ring in the scalar layout and in the quad-gather layout
(vsc_set_solver_mode bit 28 flips the build's default), per image size, pass depth and band width.

    python profiles/sweep_solver_qg.py >> gpurun_out/sweep_solver_qg.txt

Per configuration: (t(20T sweeps) - t(4T sweeps)) / 16 with CUDA events = time of one blocked pass; every configuration
is first checked bit for bit against the unblocked sweeps.  Not a bench.py number."""
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
# which vsc_set_solver_mode bit to flip: 28 = ring layout (scalar <-> quad gather), 29 = hand-off (named barriers <-> mbarriers)
FLIP_BIT = int(sys.argv[1]) if len(sys.argv) > 1 else 28
FLIPS = (0, 1)


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


for (W, H) in ((1920, 1080), (960, 540), (3840, 2160), (1280, 720), (640, 360)):
    pr = torch.rand((H, W, 3), device=dev, generator=g)
    tg = torch.rand((H, W, 3), device=dev, generator=g)
    wt = torch.rand((H, W, 3), device=dev, generator=g) * 2
    out = pr.clone()
    for tmain in (8, 10):
        V.check(L.vsc_set_solver_mode(1))
        ref = V.get_consist_out(pr, tg, wt, 2 * tmain + 3, 0.15, 0.15, pr.clone())
        row = []
        for k in (0, 1, 2, 3, 4):
            if tmain > 8 and k in (1, 2):
                continue
            for flip in FLIPS:
                # 0x4000: the 4-step-loop kernel for every pass (also the 10-sweep passes of 4K-class images)
                mode = 2 | 0x4000 | (k << 8) | (((tmain - 6) // 2) << 12) | (flip << FLIP_BIT)
                V.check(L.vsc_set_solver_mode(mode))
                got = V.get_consist_out(pr, tg, wt, 2 * tmain + 3, 0.15, 0.15, pr.clone())
                ok = torch.equal(got, ref)
                time_solve(pr, tg, wt, out, tmain * 4, reps=2)
                a = time_solve(pr, tg, wt, out, tmain * 4)
                b = time_solve(pr, tg, wt, out, tmain * 20)
                row.append((NAMES[k] + ("/flip" if flip else "") + ("" if ok else "!MISMATCH"), (b - a) / 16 * 1e3))
        L.vsc_set_solver_mode(0)
        print(f"{W}x{H} T={tmain:2d}: " + "  ".join(f"{n} {us:7.2f}" for n, us in row), flush=True)

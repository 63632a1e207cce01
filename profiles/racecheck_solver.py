"""racecheck of the blocked solver alone, one synchronisation variant per run:  python profiles/racecheck_solver.py MODE"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "video-stream-consistency_b200"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402

import vsc_b200 as V  # noqa: E402

mode = int(sys.argv[1], 0)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
W, H = 160, 96
pr = torch.rand((H, W, 3), device=dev, generator=g)
tg = torch.rand((H, W, 3), device=dev, generator=g)
wt = torch.rand((H, W, 3), device=dev, generator=g) * 2
V.check(V.lib().vsc_set_solver_mode(mode))
V.get_consist_out(pr, tg, wt, 16, 0.15, 0.15, pr.clone())
torch.cuda.synchronize()
print("done", hex(mode))

"""Small driver for ncu: launches every hot-path kernel a few times at the roofline shapes
(4K stabilization, dense-4K custom-op levels, and the 1080p solver) so that one
`ncu --set full` capture covers them without replaying a whole bench run.

    ncu --set full --clock-control none --import-source on -o gpurun_out/prof python profiles/prof_driver.py
"""
import os
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
which = set(sys.argv[1:]) or {"solver", "stage_a", "corr", "warp", "misc"}


def rnd(*shape):
    return torch.rand(shape, device=dev, generator=g)


for (W, H) in ((3840, 2160), (1920, 1080)):
    if "solver" in which:
        pr, tg, wt = rnd(H, W, 3), rnd(H, W, 3), rnd(H, W, 3) * 2
        out = pr.clone()
        V.get_consist_out(pr, tg, wt, 16, 0.15, 0.15, out)
        torch.cuda.synchronize()
    if "solver10" in which:   # the 10-sweep pass used on large images
        pr, tg, wt = rnd(H, W, 3), rnd(H, W, 3), rnd(H, W, 3) * 2
        out = pr.clone()
        V.check(V.lib().vsc_set_solver_mode(0x2002))
        V.get_consist_out(pr, tg, wt, 20, 0.15, 0.15, out)
        V.lib().vsc_set_solver_mode(0)
        torch.cuda.synchronize()
    if "stage_a" in which:
        ims = [rnd(H, W, 3) for _ in range(7)]
        ff, fb = (torch.from_numpy(x).to(dev) for x in synth.flows(W, H, 3))  # smooth, like real optical flow
        for _ in range(2):
            V.stage_a_fused(*ims, ff, fb, 6800.0, 6800.0, 2.0)
        torch.cuda.synchronize()
    if "stage_a_prep" in which:   # the stage-A kernel of the frame path (folds into the solver coefficients)
        o8, p8 = synth.frames(W, H, 3)
        of = [V.image_to_gpu(torch.from_numpy(x).to(dev)) for x in o8]
        pf = [V.image_to_gpu(torch.from_numpy(x).to(dev)) for x in p8]
        ff, fb = (torch.from_numpy(x).to(dev) for x in synth.flows(W, H, 3))
        for mode in (0, 1):   # default (row walk + flow prefetch), then one row per CTA
            V.check(V.lib().vsc_set_stage_a_mode(mode))
            for _ in range(2):
                V.frame_stabilize(of[0], of[1], of[2], pf[0], pf[1], pf[2], pf[0], ff, fb, V.HyperParams(numIter=0))
        V.lib().vsc_set_stage_a_mode(0)
        torch.cuda.synchronize()
        del of, pf, ff, fb
    if "misc" in which:
        img = rnd(H, W, 3)
        for _ in range(2):
            V.get_bilinear(img, W // 2, H // 2)
            u8 = V.gpu_to_image(img)
            V.image_to_gpu(u8)
        small = rnd(H // 2, W // 2, 3)
        for _ in range(2):
            V.get_bilinear(small, W, H)
        torch.cuda.synchronize()

if "corr" in which:
    for (C, h, w) in ((32, 544, 960), (64, 272, 480), (96, 136, 240), (64, 72, 120)):
        a, b = rnd(1, C, h, w) - 0.5, rnd(1, C, h, w) - 0.5
        for _ in range(2):
            V.correlation(a, b)
        torch.cuda.synchronize()
if "warp" in which:
    for (C, h, w) in ((32, 544, 960), (64, 272, 480), (96, 136, 240), (64, 72, 120)):
        x = rnd(1, C, h, w)
        f = torch.from_numpy(synth.op_flow_smooth(1, h, w, 3)).to(dev)
        for mode in (1, 2, 3, 4 | (1 << 4)):   # linear, tiled quad, linear quad, TMA-staged (all channels per CTA)
            V.check(V.lib().vsc_set_warp_mode(mode))
            for _ in range(2):
                V.warp(x, f)
        V.lib().vsc_set_warp_mode(0)
        torch.cuda.synchronize()
print("prof_driver done, launches:", V.launch_count())

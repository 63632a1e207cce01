"""Custom ops of one frame: the two flow directions as two batch-1 calls per op (what two session runs do) against
one batch-2 call per op (both directions stacked in the batch dimension, one session run).  Whole op sequence of a
frame back to back on one stream, CUDA events (the host enqueues ahead behind a sleep kernel, so the interval is GPU
time, not launch overhead), 8 input sets cycled (no L2 reuse between frames), median of 5 x 20.

    python profiles/time_ops_batched.py > gpurun_out/time_ops_batched.txt
"""
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
NSETS = 8


def make(N, corr_shapes, warp_shapes):
    sets = []
    for s in range(NSETS):
        corr = [(torch.randn((N, C, h, w), device=dev, generator=g), torch.randn((N, C, h, w), device=dev, generator=g),
                 torch.empty((N, 9, 9, h, w), device=dev)) for (C, h, w) in corr_shapes]
        warp = [(torch.randn((N, C, h, w), device=dev, generator=g),
                 torch.from_numpy(synth.op_flow_smooth(N, h, w, 7 * s + i)).to(dev),
                 torch.empty((N, C, h, w), device=dev)) for i, (C, h, w) in enumerate(warp_shapes)]
        sets.append((corr, warp))
    return sets


def run(sets, i, calls):
    corr, warp = sets[i % NSETS]
    for _ in range(calls):
        for a, b, o in corr:
            V.correlation(a, b, out=o)
        for x, f, o in warp:
            V.warp(x, f, out=o)


def timed(sets, calls):
    for i in range(4):
        run(sets, i, calls)
    res = []
    for _ in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(10_000_000)   # ~5 ms: the host enqueues all 20 frames meanwhile, the interval is pure GPU time
        e0.record()
        for i in range(20):
            run(sets, i, calls)
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) * 1e3 / 20)
    return statistics.median(res)


for name, cs, ws in (("light-1080p/2", synth.LIGHT_1080P_CORR, synth.LIGHT_1080P_WARP),
                     ("dense-4K", synth.DENSE_4K_CORR, synth.DENSE_4K_WARP)):
    s1 = make(1, cs, ws)
    t1 = timed(s1, 2)
    del s1
    s2 = make(2, cs, ws)
    t2 = timed(s2, 1)
    del s2
    torch.cuda.empty_cache()
    print(f"{name:14s} custom ops of one frame: 2 x batch-1 {t1:7.1f} us   1 x batch-2 {t2:7.1f} us   ({t1 - t2:+.1f} us)")

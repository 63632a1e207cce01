"""Aggregate throughput of S independent video streams sharing ONE GPU (one vsc_stabilizer each, calls interleaved
from one host thread): do concurrent streams fill the SM time a single stream leaves idle (level-1 passes, launch
gaps, conversions)?  Stabilization only (precomputed device flows), 1080p, host frames in pinned memory.

    python profiles/multi_stream.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "video-stream-consistency_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import bench  # noqa: E402
import synth  # noqa: E402
import vsc_b200 as V  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
W, H = 1920, 1080
ho, hp = bench.make_host_frames(W, H, pin=True)
flf, flb = synth.flows(W, H, 3)
d_flf, d_flb = torch.from_numpy(flf).to(dev), torch.from_numpy(flb).to(dev)
N = bench.NFRAMES
for S in (1, 2, 3, 4):
    sts = [V.Stabilizer(W, H, 3) for _ in range(S)]
    outs = [[V.pinned_empty((H, W, 4)) for _ in range(2)] for _ in range(S)]
    for st in sts:
        for t in range(3):
            st.push_frame(ho[t], hp[t])

    def run(K):
        for t in range(K):
            for i, st in enumerate(sts):
                st.step(d_flf, d_flb, outs[i][t & 1])
                st.push_frame(ho[(t + 2) % N], hp[(t + 2) % N])
        for st in sts:
            st.sync()
        torch.cuda.synchronize()

    run(5)
    K = 40
    t0 = time.perf_counter()
    run(K)
    dt = time.perf_counter() - t0
    print(f"{S} stream(s) on one GPU: {S * K / dt:7.1f} frames/s aggregate ({K / dt:6.1f} per stream)", flush=True)
    for st in sts:
        st.close()

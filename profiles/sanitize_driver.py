"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every kernel of the library at
small sizes, including the blocked solver in all its variants and the pipeline object with host frames.

    compute-sanitizer --tool memcheck python profiles/sanitize_driver.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "video-stream-consistency_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import synth  # noqa: E402
import vsc_b200 as V  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
g = torch.Generator(device=dev).manual_seed(0)
L = V.lib()

# custom ops: tiled TMA kernels (64- and 32-wide), plain stager, rows kernel, generic; both warp kernels
for (N, C, H, W) in ((1, 12, 24, 128), (2, 9, 17, 40), (1, 20, 9, 15), (1, 5, 7, 33)):
    a = torch.randn((N, C, H, W), device=dev, generator=g)
    b = torch.randn((N, C, H, W), device=dev, generator=g)
    for mode in (0, 1, 2, 3, 4, 5, 6):   # 5 / 6: round-2 shared-row kernels (64- / 32-wide skewed tiles)
        if mode in (2, 3, 5, 6) and W % 4:
            continue
        V.check(L.vsc_set_correlation_mode(mode))
        V.correlation(a, b)
        V.correlation(a, b, legacy=True)
    L.vsc_set_correlation_mode(0)
    V.correlation(a, b, max_displacement=2)
    f = 3.0 * torch.randn((N, 2, H, W), device=dev, generator=g)
    for mode in (1, 2, 3):
        V.check(L.vsc_set_warp_mode(mode))
        V.warp(a, f)
    L.vsc_set_warp_mode(0)
# TMA-staged Warp (round 2): staged tiles (smooth flow), its gather path (scattered flow), partial tiles, channel chunks
for (N, C, H, W) in ((1, 12, 40, 148), (2, 9, 33, 200)):
    a = torch.randn((N, C, H, W), device=dev, generator=g)
    for f in (torch.from_numpy(synth.op_flow_smooth(N, H, W, 3)).to(dev), 6.0 * torch.randn((N, 2, H, W), device=dev, generator=g)):
        for mode in (4, 4 | (2 << 4)):
            V.check(L.vsc_set_warp_mode(mode))
            V.warp(a, f)
    L.vsc_set_warp_mode(0)

# solver: unblocked, blocked (pair barriers, CTA barrier, private staging, no PDL, 8 and 10 sweeps), odd widths
for (W, H) in ((160, 96), (45, 37), (400, 64)):
    pr = torch.rand((H, W, 3), device=dev, generator=g)
    tg = torch.rand((H, W, 3), device=dev, generator=g)
    wt = torch.rand((H, W, 3), device=dev, generator=g) * 2
    # (round 2: the default is the 4-step loop; 0x8000 = the fully unrolled form, 0x4000 forces the loop, edge fields)
    for mode in (1, 2, 0x12, 0x22, 0x82, 0x1002, 0x2002, 0x2022, 0x8002, 0x8012, 0xA002, 0x6002, 2 | (17 << 16) | (11 << 22),
                 0x0802, 0x2802,   # balanced plan: odd pass depths
                 2 | (1 << 28), 0x2002 | (1 << 28), 2 | (1 << 30), 0x1302 | (1 << 30)):   # scalar ring, permuted column blocks
        V.check(L.vsc_set_solver_mode(mode))
        for iters in (1, 9, 21):
            V.get_consist_out(pr, tg, wt, iters, 0.15, 0.15, pr.clone())
    L.vsc_set_solver_mode(0)

# the pipeline object: host frames, device flow, low-res flow, host flow, flow-network input, 1-3 pyramid levels
for (W, H, fc) in ((128, 96, 3), (46, 38, 2)):
    o8, p8 = synth.frames(W, H, 6, seed=3)
    ff, fb = synth.flows(W, H, fc)
    st = V.Stabilizer(W, H, fc)
    for t in range(3):
        st.push_frame(o8[t], p8[t])
    out = np.zeros((H, W, 4), np.uint8)
    st.flow_input(1, W // 2, H // 2)
    st.step(torch.from_numpy(ff).to(dev), torch.from_numpy(fb).to(dev), out)
    st.push_frame(o8[3], p8[3])
    st.hyper_params.pyramidLevels = 3
    st.step_host_flow(ff, fb, out)
    st.push_frame(o8[4], p8[4])
    st.hyper_params.pyramidLevels = 1
    lf, lb = synth.flows(W // 2, H // 2, fc)
    st.step(torch.from_numpy(lf).to(dev), torch.from_numpy(lb).to(dev), out)   # low-res flow: up-scaled inside
    st.sync()
    st.close()
# fused stage A: every kernel selection (one row per CTA, row walk without / with flow prefetch, pipelined; chunk
# heights from 2 rows to taller than the image; 128- and 256-thread CTAs), frame path and public entry point
for (W, H, fc) in ((96, 40, 3), (322, 6, 2), (8, 2, 3)):
    o8, p8 = synth.frames(W, H, 3, seed=5)
    of = [V.image_to_gpu(torch.from_numpy(x).to(dev)) for x in o8]
    pf = [V.image_to_gpu(torch.from_numpy(x).to(dev)) for x in p8]
    ff, fb = (torch.from_numpy(x + np.random.default_rng(1).normal(0, 3, x.shape).astype(np.float32)).to(dev)
              for x in synth.flows(W, H, fc))
    for m in (0, 1, 2 | (1 << 4), 3 | (1 << 4), 3 | (3 << 4) | 0x100, 3 | (6 << 4), 4 | (1 << 4), 4 | (2 << 4) | 0x100,
              4 | (6 << 4)):
        V.check(L.vsc_set_stage_a_mode(m))
        V.frame_stabilize(of[0], of[1], of[2], pf[0], pf[1], pf[2], pf[0], ff, fb, V.HyperParams(numIter=3))
        V.stage_a_fused(of[0], of[1], of[2], pf[0], pf[1], pf[2], pf[0], ff, fb, 6800.0, 6800.0, 2.0, want_adap_in=True)
    L.vsc_set_stage_a_mode(0)
torch.cuda.synchronize()
print("sanitize_driver done,", V.launch_count(), "launches")

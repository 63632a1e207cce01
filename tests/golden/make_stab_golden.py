"""Generates stab_golden.npz by running the REFERENCE'S OWN CUDA stabilization kernels
(flowconsistency.cu, gpuimage.cu/.cpp compiled unmodified for sm_100a into oracle/_ref/libvsc_ref_gpu.so)
on a B200, kernel by kernel and through the doOneStep call sequence (oracle/refdrv/ref_gpu.cu).

The reference has no CPU implementation and no tests for this path, so these fixtures are what pins
the oracle (tests/test_oracle.py) and, through it and directly, the product kernels (tests/test_stab_gpu.py).

Run on the GPU box (the .so was built in the container, where /root/reference exists):
    gpurun -- 'python tests/golden/make_stab_golden.py gpurun_out/stab_golden.npz'
then copy gpurun_out/stab_golden.npz to tests/golden/.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = {"a": (64, 48), "b": (45, 37)}  # (W, H): one 4-aligned, one odd in both dimensions
T = 6


def main(path):
    import torch

    assert torch.cuda.is_available()
    out = {"gpu_name": np.array(torch.cuda.get_device_name(0))}
    for tag, (W, H) in CASES.items():
        orig8, proc8 = synth.frames(W, H, T, seed=1234 + W, mismatch=0.3 if tag == "b" else 0.0)
        ff, fb = synth.flows(W, H, 3)
        out[f"{tag}_orig8"], out[f"{tag}_proc8"] = orig8, proc8
        out[f"{tag}_flowFwd"], out[f"{tag}_flowBwd"] = ff, fb
        of = [O.ref_gpu_to_float(f) for f in orig8]
        pf = [O.ref_gpu_to_float(f) for f in proc8]
        out[f"{tag}_to_float0"] = of[0]
        # kernel by kernel on window (0,1,2), lastStab = processed frame 2
        pI = O.ref_gpu_warp_result(of[0], fb)
        pP = O.ref_gpu_warp_result(pf[0], fb)
        nI = O.ref_gpu_warp_result(of[2], ff)
        nP = O.ref_gpu_warp_result(pf[2], ff)
        lW = O.ref_gpu_warp_result(pf[2], fb)
        out[f"{tag}_warp_prevIn"], out[f"{tag}_warp_nextPr"] = pI, nP
        out[f"{tag}_warp_2ch"] = O.ref_gpu_warp_result(of[0], np.ascontiguousarray(fb[..., :2]))
        aI, aP = O.ref_gpu_adap_comb(of[1], pf[1], pI, pP, nI, nP, lW, 6800.0)
        out[f"{tag}_adapIn"], out[f"{tag}_adapPr"] = aI, aP
        wt = O.ref_gpu_consist_wt(aI, of[1], 6800.0, 2.0)
        out[f"{tag}_consWt"] = wt
        out[f"{tag}_bil_down"] = O.ref_gpu_bilinear(pf[1], W // 2, H // 2)
        out[f"{tag}_bil_up"] = O.ref_gpu_bilinear(out[f"{tag}_bil_down"], W, H)
        out[f"{tag}_bil_flow"] = O.ref_gpu_bilinear(ff[: H // 2, : W // 2].copy(), W, H)
        for it in (1, 10, 150):
            runs = [O.ref_gpu_consist_out(pf[1], aP, wt, it, 0.15, 0.15, pf[1]) for _ in range(3)]
            out[f"{tag}_solve{it}"] = runs[0]
            out[f"{tag}_solve{it}_spread"] = np.array(max(np.abs(runs[0] - r).max() for r in runs[1:]), np.float32)
        out[f"{tag}_to_char"] = O.ref_gpu_to_char(out[f"{tag}_solve150"])
        # wrap-around / no-clamp behaviour of the 8-bit conversion
        odd = np.linspace(-1.5, 2.5, W * H * 3, dtype=np.float32).reshape(H, W, 3)
        out[f"{tag}_to_char_odd_in"], out[f"{tag}_to_char_odd"] = odd, O.ref_gpu_to_char(odd)

        # the recurrence: preload 0..2, L <- P_2, then steps t = 1..3 with window (t-1, t, t+1)
        for pname, params in {"default": None, "slider": dict(numIter=40, gamma=4.0, alpha=3000.0)}.items():
            st = O.RefGpuStepper(W, H, 3, 2)
            d_of = [torch.from_numpy(x).cuda() for x in of]
            d_pf = [torch.from_numpy(x).cuda() for x in pf]
            d_ff, d_fb = torch.from_numpy(ff).cuda(), torch.from_numpy(fb).cuda()
            last = d_pf[2].clone()
            for t in (1, 2, 3):
                co, rgba = st.step(d_of[t - 1], d_of[t], d_of[t + 1], d_pf[t - 1], d_pf[t], d_pf[t + 1], last, d_ff,
                                   d_fb, params)
                out[f"{tag}_{pname}_step{t}_out"] = co
                out[f"{tag}_{pname}_step{t}_rgba"] = rgba
            st.close()
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k in sorted(out):
        if k.endswith("_spread"):
            print(k, float(out[k]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "stab_golden.npz"))

"""Generates tests/golden/ops_golden.npz by running the REFERENCE'S OWN CPU custom-op kernels
(correlation.cc / warp.cc, compiled unmodified into oracle/_ref/libvsc_ref_cpu.so by oracle/Makefile)
through CorrelationKernel::Compute / WarpKernel::Compute on seeded inputs.

Run in the build container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_ops_golden.py
The fixture pins the oracle (tests/test_oracle.py) and the CUDA kernels (tests/test_ops_gpu.py) to the
reference's outputs; the reference itself ships no golden vectors.
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


def main():
    O.build(("ref_cpu",))
    out = {}
    # Correlation: ragged shape (nothing divides the tile sizes), batch 2; and a PWC-Net-like level
    for name, (N, C, H, W, seed) in {"corr_ragged": (2, 12, 13, 21, 11), "corr_level": (1, 64, 18, 30, 12)}.items():
        a = synth.features(N, C, H, W, seed)
        b = synth.features(N, C, H, W, seed + 100)
        out[name + "_in1"], out[name + "_in2"] = a, b
        out[name + "_out"] = O.ref_cpu_correlation(a, b, 4, 0)
    # reference test.py:47-48 style input (uniform * 10), reduced size
    rng = np.random.default_rng(5)
    a = (rng.random((2, 16, 24, 24), dtype=np.float32) * 10)
    b = (rng.random((2, 16, 24, 24), dtype=np.float32) * 10)
    out["corr_testpy_in1"], out["corr_testpy_in2"] = a, b
    out["corr_testpy_out"] = O.ref_cpu_correlation(a, b, 4, 0)
    # Warp: large flow (many samples leave the image), and test.py:145-146 style flow in [0,1)
    x = synth.features(2, 5, 13, 21, 21)
    f = synth.op_flow(2, 13, 21, 22, sigma=5.0)
    out["warp_big_in"], out["warp_big_flow"], out["warp_big_out"] = x, f, O.ref_cpu_warp(x, f)
    x = (rng.random((2, 8, 16, 12), dtype=np.float32) * 10)
    f = rng.random((2, 2, 16, 12), dtype=np.float32)
    out["warp_testpy_in"], out["warp_testpy_flow"], out["warp_testpy_out"] = x, f, O.ref_cpu_warp(x, f)
    # integer and near-threshold flows: alpha/beta == 0 at the image border exercises mask > 0.999
    x = synth.features(1, 3, 9, 11, 31)
    f = np.zeros((1, 2, 9, 11), np.float32)
    f[0, 0] = np.tile(np.array([0, 1, -1, 0.0005, -0.0005, 2, 0.9995, -0.9995, 10, -10, 0.5], np.float32), (9, 1))
    f[0, 1] = np.tile(np.array([0, -1, 1, 0.0004, -0.0004, 0.9996, -2, 8, -8], np.float32)[:, None], (1, 11))
    out["warp_edge_in"], out["warp_edge_flow"], out["warp_edge_out"] = x, f, O.ref_cpu_warp(x, f)
    path = os.path.join(HERE, "ops_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

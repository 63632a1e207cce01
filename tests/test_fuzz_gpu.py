"""Seeded random-shape sweeps over the kernels whose launch geometry depends on the shape (band widths, row chunks,
channel slices, tile counts): every configuration must reproduce its plain counterpart."""
import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu


def cu(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_fuzz_blocked_solver_vs_unblocked(V, dev):
    """40 random (W, H, numIter, variant) draws, including widths that are not multiples of 4 (per-thread staging),
    images narrower than a band, a single row chunk, and both main pass depths: bit-identical to plain sweeps."""
    rng = np.random.default_rng(2024)
    g = torch.Generator(device=dev).manual_seed(7)
    L = V.lib()
    variants = (2, 0x12, 0x22, 0x82, 0x1002, 0x2002, 0x2022, 0x102, 0x402, 0x2402)
    try:
        for _ in range(40):
            W = int(rng.integers(8, 700))
            H = int(rng.integers(4, 260))
            if rng.random() < 0.5:
                W = (W + 3) // 4 * 4
            iters = int(rng.integers(1, 45))
            mode = int(variants[int(rng.integers(len(variants)))])
            pr = torch.rand((H, W, 3), device=dev, generator=g)
            tg = torch.rand((H, W, 3), device=dev, generator=g)
            wt = torch.rand((H, W, 3), device=dev, generator=g) * 2.0
            wt = wt * (wt > 0.5)
            assert L.vsc_set_solver_mode(1) == 0
            ref = V.get_consist_out(pr, tg, wt, iters, 0.15, 0.15, pr.clone())
            assert L.vsc_set_solver_mode(mode) == 0
            got = V.get_consist_out(pr, tg, wt, iters, 0.15, 0.15, pr.clone())
            assert torch.equal(got, ref), (W, H, iters, hex(mode), float((got - ref).abs().max()))
    finally:
        L.vsc_set_solver_mode(0)


def test_fuzz_correlation_kernels_vs_oracle(V, O, dev):
    """random small / medium maps through every kernel selection against the oracle (1e-4 relative)"""
    rng = np.random.default_rng(77)
    L = V.lib()
    try:
        for _ in range(16):
            N = int(rng.integers(1, 3))
            C = int(rng.integers(1, 70))
            H = int(rng.integers(1, 40))
            W = int(rng.integers(1, 70))
            if rng.random() < 0.6:
                W = (W + 3) // 4 * 4
            legacy = bool(rng.integers(2))
            a, b = synth.features(N, C, H, W, 3), synth.features(N, C, H, W, 4)
            ref = O.correlation(a, b, legacy=legacy)
            scale = max(float(np.abs(ref).max()), 1e-30)
            for mode in (0, 1, 2, 3, 4, 5, 6, 7):
                if mode in (2, 3, 5, 6, 7) and W % 4:
                    continue
                assert L.vsc_set_correlation_mode(mode) == 0
                got = V.correlation(cu(a, dev), cu(b, dev), legacy=legacy).cpu().numpy().reshape(ref.shape)
                assert float(np.abs(got - ref).max()) <= 1e-4 * scale, (N, C, H, W, legacy, mode)
    finally:
        L.vsc_set_correlation_mode(0)


def test_fuzz_warp_kernels_vs_oracle(V, O, dev):
    rng = np.random.default_rng(78)
    L = V.lib()
    try:
        for _ in range(16):
            N = int(rng.integers(1, 3))
            C = int(rng.integers(1, 40))
            H = int(rng.integers(1, 50))
            W = int(rng.integers(1, 90))
            if _ % 4 == 3:   # shapes the TMA-staged kernel takes (mode 4): W % 4 == 0, W >= 72, H >= 12, C >= 4
                C, H, W = C + 4, H + 12, 72 + 4 * int(rng.integers(0, 12))
            x = synth.features(N, C, H, W, 5)
            f = synth.op_flow(N, H, W, 6, float(rng.choice([0.3, 2.0, 30.0])))
            ref = O.warp_nchw(x, f)
            scale = max(float(np.abs(ref).max()), 1e-30)
            for mode in (1, 2, 3, 4):
                assert L.vsc_set_warp_mode(mode) == 0
                got = V.warp(cu(x, dev), cu(f, dev)).cpu().numpy()
                assert float(np.abs(got - ref).max()) <= 1e-4 * scale, (N, C, H, W, mode)
                assert np.array_equal(got == 0, ref == 0), (N, C, H, W, mode)
    finally:
        L.vsc_set_warp_mode(0)

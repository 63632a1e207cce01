"""CPU tests of the drop-in boundary: libvsc_b200.so loads without a GPU, exports exactly the symbols
include/vsc/vsc.h declares, validates its arguments before touching CUDA, and the product never imports the
oracle.  No compute is launched here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vsc", "vsc.h")
PKG = os.path.join(ROOT, "video-stream-consistency_b200")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"VSC_API\s+[^;(]*?\b(vsc_\w+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("vsc_correlation_f32", "vsc_warp_nchw_f32", "vsc_warp_hwc3", "vsc_adap_comb", "vsc_consist_wt",
                 "vsc_bilinear", "vsc_consist_solve", "vsc_stage_a_fused", "vsc_rgba8_to_f32x3",
                 "vsc_f32x3_to_rgba8", "vsc_frame_solve", "vsc_stabilizer_create", "vsc_stabilizer_step"):
        assert must in syms


def test_library_exports_every_declared_symbol(V):
    lib = C.CDLL(V.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in vsc.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", V.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (vsc_\w+)", out)))
    assert exported == declared_symbols(), "exported vsc_* symbols differ from the header"


def test_python_prototypes_cover_the_header():
    from vsc_b200._lib import PROTOTYPES

    assert sorted(PROTOTYPES) == declared_symbols()


def test_library_is_built_for_sm_100a(V):
    out = subprocess.run(["cuobjdump", "-lelf", V.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:400]


def test_argument_validation_without_gpu(V):
    L = V.lib()
    assert L.vsc_version() == 100
    assert L.vsc_error_string(0) == b"ok"
    assert b"invalid" in L.vsc_error_string(-1)
    # null pointers / bad sizes are rejected before any CUDA call
    assert L.vsc_correlation_f32(None, None, None, 1, 1, 1, 1, 4, 0, None) == -1
    assert L.vsc_warp_nchw_f32(None, None, None, 1, 1, 1, 1, None) == -1
    assert L.vsc_warp_hwc3(None, None, None, 8, 8, 3, None) == -1
    assert L.vsc_bilinear(None, 1, 1, 3, None, 1, 1, 3, None) == -1
    assert L.vsc_stage_a_fused(*([None] * 9), 3, 1.0, 1.0, 1.0, None, None, None, 8, 8, None) == -1
    one = C.c_void_p(16)
    assert L.vsc_correlation_f32(one, one, one, 0, 1, 1, 1, 4, 0, None) == -1
    assert L.vsc_warp_hwc3(one, one, one, 8, 8, 4, None) == -1          # flow channels must be 2 or 3
    assert L.vsc_bilinear(one, 4, 4, 2, one, 4, 4, 3, None) == -1      # Co > Ci
    assert L.vsc_correlation_f32(C.c_void_p(2), one, one, 1, 1, 1, 1, 4, 0, None) == -4  # misaligned
    assert L.vsc_consist_solve(one, one, one, 5, 0.1, 0.1, one, 8, 8, None, 0, None) == -2  # no workspace
    assert L.vsc_consist_solve(one, one, one, 0, 0.1, 0.1, one, 8, 8, None, 0, None) == 0   # 0 sweeps: no-op
    h = C.c_void_p(0)
    assert L.vsc_stabilizer_create(C.byref(h), 8, 8, 5) == -1
    assert L.vsc_stabilizer_push_frame(None, None, None) == -1


def test_kernel_selection_words_are_validated(V):
    """vsc_set_*_mode are plain host-side switches: valid words are accepted and reset, others rejected (no GPU)."""
    L = V.lib()
    for ok in (0, 1, 2 | (3 << 4), 3 | (4 << 4) | 0x100, 4 | (2 << 4), 3 | (7 << 12) | 0x100, 3 | (255 << 12)):
        assert L.vsc_set_stage_a_mode(ok) == 0, hex(ok)
    for bad in (-1, 5, 0x203, 0x403, 0x803, 1 << 20):
        assert L.vsc_set_stage_a_mode(bad) == -1, hex(bad)
    assert L.vsc_set_stage_a_mode(0) == 0
    for ok in (0, 1, 2, 3, 4, 4 | (2 << 4), 1 | (1 << 4), 1 | (255 << 4)):
        assert L.vsc_set_warp_mode(ok) == 0, hex(ok)
    for bad in (-1, 5, 6, 0x1001):
        assert L.vsc_set_warp_mode(bad) == -1, hex(bad)
    assert L.vsc_set_warp_mode(0) == 0
    for ok in (0, 1, 2, 3, 4, 5, 6, 7):   # 5 / 6: shared-row tiles (64 / 32 wide), 7: quad form of the channel-split kernel
        assert L.vsc_set_correlation_mode(ok) == 0, ok
    for bad in (-1, 8, 0x100):
        assert L.vsc_set_correlation_mode(bad) == -1, bad
    assert L.vsc_set_correlation_mode(0) == 0
    # solver: low nibble 0..2, flags, band (k << 8, k <= 4), depth (j << 12, j <= 2), edge fields, bit 28 (ring layout),
    # bit 30 (column blocks); bit 29 is not assigned
    for ok in (0, 1, 2, 2 | 0x10, 2 | 0x0800, 2 | 0x2400, 2 | 0x8000, 2 | (17 << 16) | (11 << 22), 2 | (1 << 28), 2 | (1 << 30)):
        assert L.vsc_set_solver_mode(ok) == 0, hex(ok)
    for bad in (-1, 3, 2 | 0x0500, 2 | 0x3000, 2 | 0xC000, 2 | (1 << 29)):
        assert L.vsc_set_solver_mode(bad) == -1, hex(bad)
    assert L.vsc_set_solver_mode(0) == 0


def test_workspace_sizes(V):
    L = V.lib()
    n = 1920 * 1080 * 3 * 4
    assert L.vsc_consist_solve_workspace_bytes(1920, 1080) >= 4 * n
    two = L.vsc_frame_solve_workspace_bytes(1920, 1080, 2)
    assert two >= 4 * n + 8 * (960 * 540 * 3 * 4)
    assert L.vsc_frame_solve_workspace_bytes(1920, 1080, 1) < two < L.vsc_frame_solve_workspace_bytes(1920, 1080, 3)
    assert L.vsc_frame_solve_workspace_bytes(1920, 1080, 0) == 0
    assert L.vsc_frame_solve_workspace_bytes(1920, 1080, 9) == 0


def test_hyper_param_defaults(V):
    p = V.HyperParams()  # videostabilizer.cpp:104-112
    assert (p.alpha, p.beta, p.gamma, p.pyramidLevels, p.numIter) == (6800.0, 6800.0, 2.0, 2, 150)
    assert abs(p.stepSize - 0.15) < 1e-7 and abs(p.momFac - 0.15) < 1e-7
    assert C.sizeof(V.HyperParams) == 28


def test_no_cpu_fallback(V):
    import torch

    with pytest.raises(V.VscError):
        V.correlation(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4))
    if not torch.cuda.is_available():
        with pytest.raises(V.VscError):
            V.Stabilizer(16, 16, 3)


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may use oracle/."""
    bad = []
    for dp, _, files in os.walk(PKG):
        if os.path.basename(dp) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", ".txt", ".cmake")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"\boracle\b", txt) and "never" not in txt.lower():
                    bad.append(os.path.join(dp, f))
                if re.search(r"vsc_oracle|libvsc_ref|/root/reference", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
    deps = subprocess.run(["ldd", os.path.join(PKG, "lib", "libvsc_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in deps and "vsc_ref" not in deps

"""TEST INFRASTRUCTURE -- numpy front-end of the oracle.  NOT product code.

Loads ``oracle/_build/libvsc_oracle.so`` (the plain-C restatement, ``vsc_oracle.c``) and, when
present, ``oracle/_ref/libvsc_ref_cpu.so`` (the reference's own CPU custom-op kernels compiled
unmodified, ``oracle/refdrv/ref_cpu_ops.cpp``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module; nothing under
``video-stream-consistency_b200/`` does.

Nothing here reads /root/reference at run time (it does not exist on the GPU box); the
libraries are built beforehand by ``oracle/Makefile`` (``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "_build", "libvsc_oracle.so")
REF_CPU_SO = os.path.join(_HERE, "_ref", "libvsc_ref_cpu.so")
REF_GPU_SO = os.path.join(_HERE, "_ref", "libvsc_ref_gpu.so")

_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)


def build(targets=("oracle", "ref")) -> None:
    """Run oracle/Makefile (building the checker is not using it)."""
    subprocess.run(["make", "-s", "-C", _HERE, *targets], check=True)


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f32p)


def _up(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u8p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(("oracle",))
        _lib = C.CDLL(ORACLE_SO)
        _lib.vsc_oracle_consist_out.restype = C.c_int
        _lib.vsc_oracle_do_one_step.restype = C.c_int
        _lib.vsc_oracle_num_threads.restype = C.c_int
    return _lib


def num_threads() -> int:
    return int(lib().vsc_oracle_num_threads())


def use_all_cores() -> int:
    """Size the OpenMP team to the cores this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().vsc_oracle_set_num_threads(int(n))
    return num_threads()


# ---------------------------------------------------------------- custom ops (NCHW fp32)
def correlation(in1, in2, max_displacement: int = 4, legacy: bool = False) -> np.ndarray:
    """custom::Correlation.  legacy=False -> [N,P,P,H,W]; legacy=True -> [N,P*P,H,W] divided by C."""
    in1, in2 = _f32(in1), _f32(in2)
    N, Cc, H, W = in1.shape
    assert in2.shape == in1.shape
    P = 2 * max_displacement + 1
    if legacy:
        out = np.empty((N, P * P, H, W), np.float32)
        lib().vsc_oracle_correlation_legacy(_fp(in1), _fp(in2), _fp(out), N, Cc, H, W, max_displacement)
    else:
        out = np.empty((N, P, P, H, W), np.float32)
        lib().vsc_oracle_correlation(_fp(in1), _fp(in2), _fp(out), N, Cc, H, W, max_displacement)
    return out


def warp_nchw(inp, flow) -> np.ndarray:
    """custom::Warp: masked bilinear backward warp, input [N,C,H,W], flow [N,2,H,W]."""
    inp, flow = _f32(inp), _f32(flow)
    N, Cc, H, W = inp.shape
    assert flow.shape == (N, 2, H, W)
    out = np.empty_like(inp)
    lib().vsc_oracle_warp_nchw(_fp(inp), _fp(flow), _fp(out), N, Cc, H, W)
    return out


# ---------------------------------------------------------------- stabilization (HWC fp32)
def warp_hwc3(img, flow) -> np.ndarray:
    img, flow = _f32(img), _f32(flow)
    H, W, c = img.shape
    assert c == 3 and flow.shape[:2] == (H, W) and flow.shape[2] in (2, 3)
    out = np.empty_like(img)
    lib().vsc_oracle_warp_hwc3(_fp(img), _fp(flow), _fp(out), W, H, flow.shape[2])
    return out


def adap_comb(crntIn, crntPr, prevWarpIn, prevWarpPr, nextWarpIn, nextWarpPr, lastStabWarp, alpha):
    a = [_f32(x) for x in (crntIn, crntPr, prevWarpIn, prevWarpPr, nextWarpIn, nextWarpPr, lastStabWarp)]
    adapIn = np.empty_like(a[0])
    adapPr = np.empty_like(a[0])
    lib().vsc_oracle_adap_comb(_fp(a[0]), _fp(a[1]), _fp(a[2]), _fp(a[3]), _fp(a[4]), _fp(a[5]), _fp(adapIn),
                               _fp(adapPr), _fp(a[6]), C.c_float(alpha), C.c_size_t(a[0].size))
    return adapIn, adapPr


def consist_wt(adapCmbIn, crntIn, beta, gamma) -> np.ndarray:
    a, b = _f32(adapCmbIn), _f32(crntIn)
    out = np.empty_like(a)
    lib().vsc_oracle_consist_wt(_fp(a), _fp(b), _fp(out), C.c_float(beta), C.c_float(gamma), C.c_size_t(a.size))
    return out


def bilinear(img, Wo: int, Ho: int, Co: int | None = None) -> np.ndarray:
    img = _f32(img)
    Hi, Wi, Ci = img.shape
    Co = Ci if Co is None else Co
    out = np.empty((Ho, Wo, Co), np.float32)
    lib().vsc_oracle_bilinear(_fp(img), Wi, Hi, Ci, _fp(out), Wo, Ho, Co)
    return out


def consist_out(crntPr, prevStabWarp, consWt, numIter, stepSize, momFac, init, mode: int = 0) -> np.ndarray:
    """get_consist_out; `init` is the caller-initialised consisOut.  mode 0 Jacobi, 1 in-place raster."""
    pr, tg, wt = _f32(crntPr), _f32(prevStabWarp), _f32(consWt)
    out = np.array(init, dtype=np.float32, order="C", copy=True)
    H, W, _ = pr.shape
    rc = lib().vsc_oracle_consist_out(_fp(pr), _fp(tg), _fp(wt), int(numIter), C.c_float(stepSize), C.c_float(momFac),
                                      _fp(out), W, H, int(mode))
    assert rc == 0
    return out


def rgba8_to_f32x3(rgba) -> np.ndarray:
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    H, W, c = rgba.shape
    assert c == 4
    out = np.empty((H, W, 3), np.float32)
    lib().vsc_oracle_rgba8_to_f32x3(_up(rgba), _fp(out), W, H)
    return out


def f32x3_to_rgba8(img) -> np.ndarray:
    img = _f32(img)
    H, W, c = img.shape
    assert c == 3
    out = np.empty((H, W, 4), np.uint8)
    lib().vsc_oracle_f32x3_to_rgba8(_fp(img), _up(out), W, H)
    return out


DEFAULT_PARAMS = dict(alpha=6800.0, beta=6800.0, gamma=2.0, pyramidLevels=2, numIter=150, stepSize=0.15, momFac=0.15)


def do_one_step(origPrev, origCur, origNext, procPrev, procCur, procNext, lastStab, flowFwd, flowBwd, params=None,
                mode: int = 0):
    """One frame of VideoStabilizer::doOneStep.  Returns (consisOut f32 HWC, rgba8 HWC4); `lastStab` is not
    modified (the new recurrence state is the returned consisOut)."""
    p = dict(DEFAULT_PARAMS)
    if params:
        p.update(params)
    a = [_f32(x) for x in (origPrev, origCur, origNext, procPrev, procCur, procNext)]
    last = np.array(lastStab, dtype=np.float32, order="C", copy=True)
    ff, fb = _f32(flowFwd), _f32(flowBwd)
    H, W, _ = a[0].shape
    assert ff.shape == fb.shape and ff.shape[:2] == (H, W)
    out = np.empty((H, W, 3), np.float32)
    rgba = np.empty((H, W, 4), np.uint8)
    rc = lib().vsc_oracle_do_one_step(_fp(a[0]), _fp(a[1]), _fp(a[2]), _fp(a[3]), _fp(a[4]), _fp(a[5]), _fp(last),
                                      _fp(ff), _fp(fb), W, H, ff.shape[2], C.c_float(p["alpha"]), C.c_float(p["beta"]),
                                      C.c_float(p["gamma"]), int(p["pyramidLevels"]), int(p["numIter"]),
                                      C.c_float(p["stepSize"]), C.c_float(p["momFac"]), int(mode), _fp(out), _up(rgba))
    assert rc == 0
    return out, rgba


# ---------------------------------------------------------------- the reference's own CPU ops (oracle/_ref)
_ref_cpu = None


def ref_cpu_available() -> bool:
    return os.path.exists(REF_CPU_SO)


def ref_cpu():
    global _ref_cpu
    if _ref_cpu is None:
        _ref_cpu = C.CDLL(REF_CPU_SO)
    return _ref_cpu


def ref_cpu_correlation(in1, in2, max_displacement: int = 4, legacy: int = 0) -> np.ndarray:
    """CorrelationKernel::Compute (CPU provider) of the reference, on host arrays."""
    in1, in2 = _f32(in1), _f32(in2)
    N, Cc, H, W = in1.shape
    P = 2 * max_displacement + 1
    out = np.empty((N, P, P, H, W), np.float32)
    dims = (C.c_int64 * 8)()
    rank = C.c_int(0)
    rc = ref_cpu().vsc_ref_cpu_correlation(_fp(in1), _fp(in2), _fp(out), C.c_size_t(out.nbytes), C.c_int64(N),
                                           C.c_int64(Cc), C.c_int64(H), C.c_int64(W), C.c_int64(max_displacement),
                                           C.c_int64(legacy), dims, C.byref(rank))
    if rc != 0:
        raise RuntimeError("reference CorrelationKernel::Compute threw")
    assert list(dims[: rank.value]) == [N, P, P, H, W]
    return out


def ref_cpu_warp(inp, flow) -> np.ndarray:
    """WarpKernel::Compute (CPU provider) of the reference, on host arrays."""
    inp, flow = _f32(inp), _f32(flow)
    N, Cc, H, W = inp.shape
    out = np.empty_like(inp)
    rc = ref_cpu().vsc_ref_cpu_warp(_fp(inp), _fp(flow), _fp(out), C.c_size_t(out.nbytes), C.c_int64(N), C.c_int64(Cc),
                                    C.c_int64(H), C.c_int64(W))
    if rc != 0:
        raise RuntimeError("reference WarpKernel::Compute threw")
    return out


# ---------------------------------------------------------------- the reference's own CUDA code (oracle/_ref)
# Device memory comes from torch; inputs/outputs are numpy arrays.  GPU box only.
_ref_gpu = None


def ref_gpu_available() -> bool:
    return os.path.exists(REF_GPU_SO)


def ref_gpu():
    global _ref_gpu
    if _ref_gpu is None:
        _ref_gpu = C.CDLL(REF_GPU_SO)
        _ref_gpu.vsc_ref_gpu_step_create.restype = C.c_void_p
    return _ref_gpu


def _dev(a, dtype=None):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(a if dtype is None else a.astype(dtype))).cuda()
    return t


def _dp(t):
    return C.c_void_p(t.data_ptr())


def _sync():
    import torch

    torch.cuda.synchronize()


def ref_gpu_correlation(in1, in2, max_displacement=4, legacy=0) -> np.ndarray:
    import torch

    a, b = _dev(_f32(in1)), _dev(_f32(in2))
    N, Cc, H, W = a.shape
    P = 2 * max_displacement + 1
    out = torch.zeros((N, P, P, H, W), device="cuda", dtype=torch.float32)
    _sync()
    rc = ref_gpu().vsc_ref_gpu_correlation(_dp(a), _dp(b), _dp(out), C.c_size_t(out.numel() * 4), C.c_int64(N),
                                           C.c_int64(Cc), C.c_int64(H), C.c_int64(W), C.c_int64(max_displacement),
                                           C.c_int64(legacy), C.c_void_p(0))
    assert rc == 0
    _sync()
    return out.cpu().numpy()


def ref_gpu_warp(inp, flow) -> np.ndarray:
    import torch

    a, f = _dev(_f32(inp)), _dev(_f32(flow))
    N, Cc, H, W = a.shape
    out = torch.zeros_like(a)
    _sync()
    rc = ref_gpu().vsc_ref_gpu_warp(_dp(a), _dp(f), _dp(out), C.c_size_t(out.numel() * 4), C.c_int64(N),
                                    C.c_int64(Cc), C.c_int64(H), C.c_int64(W), C.c_void_p(0))
    assert rc == 0
    _sync()
    return out.cpu().numpy()


def ref_gpu_warp_result(img, flow) -> np.ndarray:
    import torch

    a, f = _dev(_f32(img)), _dev(_f32(flow))
    H, W, _ = a.shape
    out = torch.zeros_like(a)
    _sync()
    assert ref_gpu().vsc_ref_gpu_warp_result(_dp(a), _dp(f), _dp(out), W, H, int(f.shape[2])) == 0
    return out.cpu().numpy()


def ref_gpu_adap_comb(crntIn, crntPr, prevWarpIn, prevWarpPr, nextWarpIn, nextWarpPr, lastStabWarp, alpha):
    import torch

    t = [_dev(_f32(x)) for x in (crntIn, crntPr, prevWarpIn, prevWarpPr, nextWarpIn, nextWarpPr, lastStabWarp)]
    H, W, _ = t[0].shape
    ai, ap = torch.zeros_like(t[0]), torch.zeros_like(t[0])
    _sync()
    assert ref_gpu().vsc_ref_gpu_adap_comb(_dp(t[0]), _dp(t[1]), _dp(t[2]), _dp(t[3]), _dp(t[4]), _dp(t[5]), _dp(ai),
                                           _dp(ap), _dp(t[6]), C.c_float(alpha), W, H) == 0
    return ai.cpu().numpy(), ap.cpu().numpy()


def ref_gpu_consist_wt(adapCmbIn, crntIn, beta, gamma) -> np.ndarray:
    import torch

    a, b = _dev(_f32(adapCmbIn)), _dev(_f32(crntIn))
    H, W, _ = a.shape
    out = torch.zeros_like(a)
    _sync()
    assert ref_gpu().vsc_ref_gpu_consist_wt(_dp(a), _dp(b), _dp(out), C.c_float(beta), C.c_float(gamma), W, H) == 0
    return out.cpu().numpy()


def ref_gpu_bilinear(img, Wo, Ho, Co=None) -> np.ndarray:
    import torch

    a = _dev(_f32(img))
    Hi, Wi, Ci = a.shape
    Co = Ci if Co is None else Co
    out = torch.zeros((Ho, Wo, Co), device="cuda", dtype=torch.float32)
    _sync()
    assert ref_gpu().vsc_ref_gpu_bilinear(_dp(a), Wi, Hi, Ci, _dp(out), Wo, Ho, Co) == 0
    return out.cpu().numpy()


def ref_gpu_consist_out(crntPr, prevStabWarp, consWt, numIter, stepSize, momFac, init) -> np.ndarray:
    a, b, c = _dev(_f32(crntPr)), _dev(_f32(prevStabWarp)), _dev(_f32(consWt))
    out = _dev(_f32(init)).clone()
    H, W, _ = a.shape
    _sync()
    assert ref_gpu().vsc_ref_gpu_consist_out(_dp(a), _dp(b), _dp(c), int(numIter), C.c_float(stepSize),
                                             C.c_float(momFac), _dp(out), W, H) == 0
    return out.cpu().numpy()


def ref_gpu_to_float(rgba) -> np.ndarray:
    import torch

    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    H, W, _ = rgba.shape
    out = torch.zeros((H, W, 3), device="cuda", dtype=torch.float32)
    _sync()
    assert ref_gpu().vsc_ref_gpu_to_float(_up(rgba), _dp(out), W, H) == 0
    return out.cpu().numpy()


def ref_gpu_to_char(img) -> np.ndarray:
    a = _dev(_f32(img))
    H, W, _ = a.shape
    out = np.zeros((H, W, 4), np.uint8)
    _sync()
    assert ref_gpu().vsc_ref_gpu_to_char(_dp(a), _up(out), W, H) == 0
    return out


class RefGpuStepper:
    """doOneStep of the reference (its own kernels, its own temporaries), on device tensors."""

    def __init__(self, W, H, flowC=3, levels=2):
        self.W, self.H, self.flowC = W, H, flowC
        self.h = C.c_void_p(ref_gpu().vsc_ref_gpu_step_create(W, H, flowC, levels))
        assert self.h

    def step(self, oP, oC, oN, pP, pC, pN, last, ff, fb, params=None):
        """all arguments torch CUDA float tensors; `last` is updated in place.  -> (consisOut np, rgba np)"""
        import torch

        p = dict(DEFAULT_PARAMS)
        if params:
            p.update(params)
        out = torch.zeros((self.H, self.W, 3), device="cuda", dtype=torch.float32)
        rgba = np.zeros((self.H, self.W, 4), np.uint8)
        _sync()
        rc = ref_gpu().vsc_ref_gpu_step(self.h, _dp(oP), _dp(oC), _dp(oN), _dp(pP), _dp(pC), _dp(pN), _dp(last),
                                        _dp(ff), _dp(fb), C.c_float(p["alpha"]), C.c_float(p["beta"]),
                                        C.c_float(p["gamma"]), int(p["numIter"]), C.c_float(p["stepSize"]),
                                        C.c_float(p["momFac"]), _dp(out), _up(rgba))
        assert rc == 0
        return out.cpu().numpy(), rgba

    def close(self):
        if self.h:
            ref_gpu().vsc_ref_gpu_step_destroy(self.h)
            self.h = None

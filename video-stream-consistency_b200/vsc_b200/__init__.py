"""vsc_b200 -- Python host side of the B200-native hot path (ctypes over the C ABI, include/vsc/vsc.h).

The reference's host is C++ (the drop-in shims live in ../host/); this package is the thin binding
the tests and bench.py use.  PyTorch is used for device memory and streams only: every function
takes CUDA tensors, passes raw device pointers to libvsc_b200.so and enqueues on torch's current
stream.  There is NO CPU fallback: if the library is missing, or a tensor is not on a CUDA device,
the call raises.

Names mirror the reference:
  * custom ops (src/ort_custom_ops): ``correlation`` (custom::Correlation), ``warp`` (custom::Warp)
  * flowconsistency.cuh: ``get_warp_result``, ``get_adap_comb``, ``get_consist_wt``, ``get_bilinear``,
    ``get_consist_out``
  * gpuimage: ``image_to_gpu`` (imageToGPU / copyFromQImage), ``gpu_to_image`` (copyToQImage)
  * videostabilizer: ``HyperParams``, ``Stabilizer`` (preload / doOneStep recurrence)
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from ._lib import LIB_PATH, VscError, check, lib

__all__ = [
    "LIB_PATH", "VscError", "lib", "launch_count", "correlation", "warp", "get_warp_result", "get_adap_comb",
    "get_consist_wt", "get_bilinear", "get_consist_out", "image_to_gpu", "gpu_to_image", "stage_a_fused",
    "frame_solve", "frame_stabilize", "HyperParams", "Stabilizer", "pinned_empty",
]


def launch_count() -> int:
    return int(lib().vsc_launch_count())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t: torch.Tensor, name: str) -> C.c_void_p:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise VscError(f"{name}: expected a CUDA tensor (no CPU fallback exists)")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise VscError(f"{name}: expected contiguous float32")
    return C.c_void_p(t.data_ptr())


def _u8(t: torch.Tensor, name: str) -> C.c_void_p:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise VscError(f"{name}: expected a CUDA tensor (no CPU fallback exists)")
    if t.dtype != torch.uint8 or not t.is_contiguous():
        raise VscError(f"{name}: expected contiguous uint8")
    return C.c_void_p(t.data_ptr())


# ------------------------------------------------------------------ custom ops
def correlation(in1: torch.Tensor, in2: torch.Tensor, max_displacement: int = 4, legacy: bool = False,
                out: torch.Tensor | None = None) -> torch.Tensor:
    """custom::Correlation.  [N,C,H,W] x2 -> [N,P,P,H,W] (legacy: [N,P*P,H,W], divided by C)."""
    N, Cc, H, W = in1.shape
    if tuple(in2.shape) != (N, Cc, H, W):
        raise VscError("correlation: input shapes differ")
    P = 2 * max_displacement + 1
    shape = (N, P * P, H, W) if legacy else (N, P, P, H, W)
    if out is None:
        out = torch.empty(shape, device=in1.device, dtype=torch.float32)
    elif tuple(out.shape) != shape:
        raise VscError("correlation: bad output shape")
    check(lib().vsc_correlation_f32(_f32(in1, "in1"), _f32(in2, "in2"), _f32(out, "out"), N, Cc, H, W,
                                    int(max_displacement), int(bool(legacy)), _stream()))
    return out


def warp(inp: torch.Tensor, flow: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """custom::Warp.  input [N,C,H,W], flow [N,2,H,W] -> [N,C,H,W]."""
    N, Cc, H, W = inp.shape
    if tuple(flow.shape) != (N, 2, H, W):
        raise VscError("warp: flow must be [N,2,H,W]")
    if out is None:
        out = torch.empty_like(inp)
    check(lib().vsc_warp_nchw_f32(_f32(inp, "input"), _f32(flow, "flow"), _f32(out, "out"), N, Cc, H, W, _stream()))
    return out


# ------------------------------------------------------------------ flowconsistency.cuh
def _hw3(t: torch.Tensor, name: str):
    if t.dim() != 3 or t.shape[2] != 3:
        raise VscError(f"{name}: expected an HWC image with 3 channels")
    return int(t.shape[0]), int(t.shape[1])


def get_warp_result(inp: torch.Tensor, flow: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    H, W = _hw3(inp, "input")
    if flow.dim() != 3 or tuple(flow.shape[:2]) != (H, W) or flow.shape[2] not in (2, 3):
        raise VscError("get_warp_result: flow must be HWC with 2 or 3 channels at the image size")
    if out is None:
        out = torch.empty_like(inp)
    check(lib().vsc_warp_hwc3(_f32(inp, "input"), _f32(flow, "flow"), _f32(out, "out"), W, H, int(flow.shape[2]),
                              _stream()))
    return out


def get_adap_comb(crntIn, crntPr, prevWarpIn, prevWarpPr, nextWarpIn, nextWarpPr, lastStabWarp, alpha: float,
                  want_adap_in: bool = True):
    H, W = _hw3(crntIn, "crntIn")
    adapIn = torch.empty_like(crntIn) if want_adap_in else None
    adapPr = torch.empty_like(crntIn)
    check(lib().vsc_adap_comb(_f32(crntIn, "crntIn"), _f32(crntPr, "crntPr"), _f32(prevWarpIn, "prevWarpIn"),
                              _f32(prevWarpPr, "prevWarpPr"), _f32(nextWarpIn, "nextWarpIn"),
                              _f32(nextWarpPr, "nextWarpPr"),
                              _f32(adapIn, "adapCmbIn") if adapIn is not None else C.c_void_p(0),
                              _f32(adapPr, "adapCmbPr"), _f32(lastStabWarp, "lastStabWarp"), C.c_float(alpha), W, H,
                              _stream()))
    return adapIn, adapPr


def get_consist_wt(adapCmbIn, crntIn, beta: float, gamma: float) -> torch.Tensor:
    H, W = _hw3(crntIn, "crntIn")
    out = torch.empty_like(crntIn)
    check(lib().vsc_consist_wt(_f32(adapCmbIn, "adapCmbIn"), _f32(crntIn, "crntIn"), _f32(out, "consWt"),
                               C.c_float(beta), C.c_float(gamma), W, H, _stream()))
    return out


def get_bilinear(inp: torch.Tensor, Wo: int, Ho: int, Co: int | None = None) -> torch.Tensor:
    Hi, Wi, Ci = (int(v) for v in inp.shape)
    Co = Ci if Co is None else int(Co)
    out = torch.empty((Ho, Wo, Co), device=inp.device, dtype=torch.float32)
    check(lib().vsc_bilinear(_f32(inp, "input"), Wi, Hi, Ci, _f32(out, "out"), int(Wo), int(Ho), Co, _stream()))
    return out


def get_consist_out(crntPr, prevStabWarp, consWt, numIter: int, stepSize: float, momFac: float,
                    consisOut: torch.Tensor) -> torch.Tensor:
    """In place on consisOut (caller-initialised), like the reference."""
    H, W = _hw3(crntPr, "crntPr")
    nbytes = int(lib().vsc_consist_solve_workspace_bytes(W, H))
    ws = torch.empty(nbytes, device=crntPr.device, dtype=torch.uint8)
    check(lib().vsc_consist_solve(_f32(crntPr, "crntPr"), _f32(prevStabWarp, "prevStabWarp"), _f32(consWt, "consWt"),
                                  int(numIter), C.c_float(stepSize), C.c_float(momFac), _f32(consisOut, "consisOut"),
                                  W, H, C.c_void_p(ws.data_ptr()), C.c_size_t(nbytes), _stream()))
    return consisOut


def image_to_gpu(rgba: torch.Tensor) -> torch.Tensor:
    """RGBA8888 device bytes [H,W,4] -> float3 [H,W,3] (copyFromQImage's kernel)."""
    H, W, c = (int(v) for v in rgba.shape)
    if c != 4:
        raise VscError("image_to_gpu: expected [H,W,4] uint8")
    out = torch.empty((H, W, 3), device=rgba.device, dtype=torch.float32)
    check(lib().vsc_rgba8_to_f32x3(_u8(rgba, "rgba"), _f32(out, "out"), W, H, _stream()))
    return out


def gpu_to_image(img: torch.Tensor) -> torch.Tensor:
    """float3 [H,W,3] -> RGBA8888 device bytes [H,W,4] (copyToQImage's kernel)."""
    H, W = _hw3(img, "image")
    out = torch.empty((H, W, 4), device=img.device, dtype=torch.uint8)
    check(lib().vsc_f32x3_to_rgba8(_f32(img, "image"), _u8(out, "out"), W, H, _stream()))
    return out


def rgba8_scale_nearest(rgba: torch.Tensor, dstW: int, dstH: int) -> torch.Tensor:
    """RGBA8888 device bytes [H,W,4] -> [dstH,dstW,4], nearest neighbour at pixel centres (vsc_rgba8_scale_nearest)."""
    H, W, c = (int(v) for v in rgba.shape)
    if c != 4:
        raise VscError("rgba8_scale_nearest: expected [H,W,4] uint8")
    out = torch.empty((int(dstH), int(dstW), 4), device=rgba.device, dtype=torch.uint8)
    check(lib().vsc_rgba8_scale_nearest(_u8(rgba, "rgba"), W, H, _u8(out, "out"), int(dstW), int(dstH), _stream()))
    return out


# ------------------------------------------------------------------ fused per-frame pieces
class HyperParams(C.Structure):
    """hyperParams (videostabilizer.h:38-46), same field order; defaults of initHyperParams."""
    _fields_ = [("alpha", C.c_float), ("beta", C.c_float), ("gamma", C.c_float), ("pyramidLevels", C.c_int),
                ("numIter", C.c_int), ("stepSize", C.c_float), ("momFac", C.c_float)]

    def __init__(self, **kw):
        super().__init__()
        lib().vsc_hyper_params_default(C.byref(self))
        for k, v in kw.items():
            if not hasattr(self, k):
                raise VscError(f"unknown hyper-parameter {k}")
            setattr(self, k, v)


def stage_a_fused(origPrev, origCur, origNext, procPrev, procCur, procNext, lastStab, flowFwd, flowBwd,
                  alpha: float, beta: float, gamma: float, want_adap_in: bool = False):
    H, W = _hw3(origCur, "origCur")
    fc = int(flowFwd.shape[2])
    adapIn = torch.empty_like(origCur) if want_adap_in else None
    adapPr = torch.empty_like(origCur)
    consWt = torch.empty_like(origCur)
    check(lib().vsc_stage_a_fused(_f32(origPrev, "origPrev"), _f32(origCur, "origCur"), _f32(origNext, "origNext"),
                                  _f32(procPrev, "procPrev"), _f32(procCur, "procCur"), _f32(procNext, "procNext"),
                                  _f32(lastStab, "lastStab"), _f32(flowFwd, "flowFwd"), _f32(flowBwd, "flowBwd"), fc,
                                  C.c_float(alpha), C.c_float(beta), C.c_float(gamma),
                                  _f32(adapIn, "adapCmbIn") if adapIn is not None else C.c_void_p(0),
                                  _f32(adapPr, "adapCmbPr"), _f32(consWt, "consWt"), W, H, _stream()))
    return adapIn, adapPr, consWt


def frame_solve(procCur, adapCmbPr, consWt, params: HyperParams, workspace: torch.Tensor | None = None,
                out: torch.Tensor | None = None) -> torch.Tensor:
    H, W = _hw3(procCur, "procCur")
    nbytes = int(lib().vsc_frame_solve_workspace_bytes(W, H, params.pyramidLevels))
    if workspace is None:
        workspace = torch.empty(nbytes, device=procCur.device, dtype=torch.uint8)
    if out is None:
        out = torch.empty_like(procCur)
    check(lib().vsc_frame_solve(_f32(procCur, "procCur"), _f32(adapCmbPr, "adapCmbPr"), _f32(consWt, "consWt"),
                                C.byref(params), _f32(out, "consisOut"), W, H, C.c_void_p(workspace.data_ptr()),
                                C.c_size_t(workspace.numel()), _stream()))
    return out


def frame_stabilize(origPrev, origCur, origNext, procPrev, procCur, procNext, lastStab, flowFwd, flowBwd,
                    params: HyperParams, out: torch.Tensor | None = None,
                    workspace: torch.Tensor | None = None) -> torch.Tensor:
    """Stage A + pyramid + solve of one frame (doOneStep between the flow and the 8-bit conversion)."""
    H, W = _hw3(origCur, "origCur")
    nbytes = int(lib().vsc_frame_stabilize_workspace_bytes(W, H, params.pyramidLevels))
    ws = workspace if workspace is not None else torch.empty(nbytes, device=origCur.device, dtype=torch.uint8)
    nbytes = ws.numel()
    if out is None:
        out = torch.empty_like(origCur)
    check(lib().vsc_frame_stabilize(_f32(origPrev, "origPrev"), _f32(origCur, "origCur"), _f32(origNext, "origNext"),
                                    _f32(procPrev, "procPrev"), _f32(procCur, "procCur"), _f32(procNext, "procNext"),
                                    _f32(lastStab, "lastStab"), _f32(flowFwd, "flowFwd"), _f32(flowBwd, "flowBwd"),
                                    int(flowFwd.shape[2]), C.byref(params), _f32(out, "consisOut"), W, H,
                                    C.c_void_p(ws.data_ptr()), C.c_size_t(nbytes), _stream()))
    return out


def pinned_empty(shape, dtype=torch.uint8) -> torch.Tensor:
    return torch.empty(shape, dtype=dtype, pin_memory=True)


class Stabilizer:
    """One video stream on the current CUDA device: VideoStabilizer's preload + doOneStep recurrence
    (videostabilizer.cpp:136-153,167-265) on top of vsc_stabilizer_*."""

    def __init__(self, width: int, height: int, flow_channels: int = 3, batch_size: int = 1):
        self._h = C.c_void_p(0)
        if not torch.cuda.is_available():
            raise VscError("Stabilizer needs a CUDA device (no CPU fallback exists)")
        self.W, self.H, self.flow_channels = int(width), int(height), int(flow_channels)
        self.batch_size = int(batch_size)   # `-b`: the window holds 2 + batch_size frames (videostabilizer.cpp:136-153)
        self._h = C.c_void_p(0)
        check(lib().vsc_stabilizer_create_batched(C.byref(self._h), self.W, self.H, self.flow_channels, self.batch_size))
        self._inflight = []     # tensors / host buffers the asynchronous steps still use (released by sync())
        self._xstream = None

    # The pipeline object runs on its own non-blocking compute stream.  Device tensors handed to it were produced
    # on torch's current stream, and tensors it fills are consumed there: order the two with events, and keep
    # every buffer of an enqueued step referenced until the next sync() so that torch's caching allocator (or the
    # garbage collector, for host arrays) cannot recycle memory the stream still uses.
    def _compute_ext(self):
        if self._xstream is None:
            self._xstream = torch.cuda.ExternalStream(self.compute_stream)
        return self._xstream

    def _hold(self, *bufs):
        self._inflight.append(bufs)
        if len(self._inflight) > 64:        # a caller that never syncs: calls that far back are long done
            del self._inflight[:32]

    def _after_torch(self, *bufs):
        self._compute_ext().wait_stream(torch.cuda.current_stream())
        self._hold(*bufs)

    def _before_torch(self):
        torch.cuda.current_stream().wait_stream(self._compute_ext())

    def close(self):
        if getattr(self, "_h", None) and lib is not None:   # `lib` is None while the interpreter shuts down
            lib().vsc_stabilizer_destroy(self._h)
            self._h = C.c_void_p(0)

    __del__ = close

    @property
    def hyper_params(self) -> HyperParams:
        p = lib().vsc_stabilizer_hyper_params(self._h)
        return C.cast(p, C.POINTER(HyperParams)).contents

    @staticmethod
    def _host_u8(t, name):
        if isinstance(t, torch.Tensor):
            if t.is_cuda or t.dtype != torch.uint8 or not t.is_contiguous():
                raise VscError(f"{name}: expected a contiguous uint8 HOST tensor")
            return C.c_void_p(t.data_ptr())
        import numpy as np  # numpy array
        if not (isinstance(t, np.ndarray) and t.dtype == np.uint8 and t.flags["C_CONTIGUOUS"]):
            raise VscError(f"{name}: expected a contiguous uint8 host array")
        return C.c_void_p(t.ctypes.data)

    def push_frame(self, orig_rgba_host, proc_rgba_host):
        """loadFrame: host RGBA8888 [H,W,4] x2 (pinned tensors are copied without staging)."""
        self._hold(orig_rgba_host, proc_rgba_host)   # pinned frames are read asynchronously
        check(lib().vsc_stabilizer_push_frame(self._h, self._host_u8(orig_rgba_host, "orig"),
                                              self._host_u8(proc_rgba_host, "proc")))

    def step(self, flowFwd: torch.Tensor, flowBwd: torch.Tensor, out_rgba_host=None):
        """doOneStep.  Flows: device HWC at frame resolution, or lower (FLOWDOWNSCALE) -> upsampled inside."""
        fh, fw = int(flowFwd.shape[0]), int(flowFwd.shape[1])
        outp = self._host_u8(out_rgba_host, "out") if out_rgba_host is not None else C.c_void_p(0)
        self._after_torch(flowFwd, flowBwd, out_rgba_host)
        if (fw, fh) == (self.W, self.H):
            check(lib().vsc_stabilizer_step(self._h, _f32(flowFwd, "flowFwd"), _f32(flowBwd, "flowBwd"), outp))
        else:
            check(lib().vsc_stabilizer_step_lowres_flow(self._h, _f32(flowFwd, "flowFwd"), _f32(flowBwd, "flowBwd"),
                                                        fw, fh, outp))

    def step_host_flow(self, flowFwd_host, flowBwd_host, out_rgba_host=None):
        """doOneStep with precomputed flows in HOST memory (.flo mode): float32 [h,w,flow_channels] arrays."""
        import numpy as np

        def fptr(t, name):
            if isinstance(t, torch.Tensor):
                if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                    raise VscError(f"{name}: expected a contiguous float32 HOST tensor")
                return C.c_void_p(t.data_ptr()), int(t.shape[1]), int(t.shape[0]), int(t.shape[2])
            if not (isinstance(t, np.ndarray) and t.dtype == np.float32 and t.flags["C_CONTIGUOUS"]):
                raise VscError(f"{name}: expected a contiguous float32 host array")
            return C.c_void_p(t.ctypes.data), int(t.shape[1]), int(t.shape[0]), int(t.shape[2])

        pf, fw, fh, fc = fptr(flowFwd_host, "flowFwd")
        pb, bw, bh, bc = fptr(flowBwd_host, "flowBwd")
        if (fw, fh, fc) != (bw, bh, bc) or fc != self.flow_channels:
            raise VscError("step_host_flow: flows must have the same shape and the stabilizer's channel count")
        self._hold(flowFwd_host, flowBwd_host, out_rgba_host)
        outp = self._host_u8(out_rgba_host, "out") if out_rgba_host is not None else C.c_void_p(0)
        check(lib().vsc_stabilizer_step_host_flow(self._h, pf, pb, fw, fh, outp))

    def step_flow_files(self, flow_dir: str, current_frame: int, out_rgba_host=None):
        """doOneStep of the file mode (-f <flowdir>): reads <flow_dir>/frame_%06d.flo (current_frame + 1) and
        frame_%06d_bwd.flo (current_frame) like FileStabilizer::retrieveOpticalFlow, then steps."""
        outp = self._host_u8(out_rgba_host, "out") if out_rgba_host is not None else C.c_void_p(0)
        self._hold(out_rgba_host)
        check(lib().vsc_stabilizer_step_flow_files(self._h, os.fsencode(flow_dir), int(current_frame), outp),
              os.fspath(flow_dir))

    def flow_input(self, window_index: int, netW: int, netH: int, out: torch.Tensor | None = None) -> torch.Tensor:
        """RGBA8 frame `window_index` (0 prev, 1 cur, 2 next) of the original stream at the flow network's input
        size, produced on the device (FlowModel::run's input path); ordered on the pipeline's compute stream."""
        if out is None:
            out = torch.empty((int(netH), int(netW), 4), device=torch.device("cuda", torch.cuda.current_device()),
                              dtype=torch.uint8)
        self._after_torch(out)
        check(lib().vsc_stabilizer_flow_input(self._h, int(window_index), _u8(out, "out"), int(netW), int(netH)))
        self._before_torch()
        return out

    def prefetch_flow_files(self, flow_dir: str, current_frame: int):
        """start reading the .flo pair of `current_frame` in the background (see vsc_stabilizer_prefetch_flow_files)"""
        check(lib().vsc_stabilizer_prefetch_flow_files(self._h, os.fsencode(flow_dir), int(current_frame)))

    def sync(self):
        check(lib().vsc_stabilizer_sync(self._h))
        self._inflight.clear()

    def reset(self):
        check(lib().vsc_stabilizer_reset(self._h))

    @property
    def compute_stream(self) -> int:
        return int(lib().vsc_stabilizer_compute_stream(self._h) or 0)

    def last_output(self) -> torch.Tensor:
        """fp32 result of the last step (a copy), [H,W,3] on the device."""
        out = torch.empty((self.H, self.W, 3), device="cuda", dtype=torch.float32)
        self._after_torch(out)
        check(lib().vsc_stabilizer_copy_last_output(self._h, C.c_void_p(out.data_ptr())))
        self._before_torch()
        self.sync()      # (also delivers pending 8-bit results to pageable host buffers)
        return out


# ------------------------------------------------------------------ .flo ingestion (host side)
def flo_frame_path(flow_dir: str, frame: int, backward: bool = False) -> str:
    """<flow_dir>/frame_%06d.flo or frame_%06d_bwd.flo (stabilizefiles.cpp:141-144)."""
    buf = C.create_string_buffer(4096)
    check(lib().vsc_flo_frame_path(os.fsencode(flow_dir), int(frame), 1 if backward else 0, buf, len(buf)))
    return os.fsdecode(buf.value)


def flo_read(path: str, pinned: bool = False) -> torch.Tensor:
    """ReadFlowFile (flowIO.cpp:31-78): -> float32 HOST tensor [H,W,2]; raises VscError with the reference's
    message on the reference's error conditions."""
    w, h = C.c_int(0), C.c_int(0)
    bpath = os.fsencode(path)
    check(lib().vsc_flo_read_header(bpath, C.byref(w), C.byref(h)), os.fspath(path))
    out = pinned_empty((h.value, w.value, 2), torch.float32) if pinned else torch.empty((h.value, w.value, 2))
    check(lib().vsc_flo_read(bpath, C.c_void_p(out.data_ptr()), C.c_size_t(out.numel()), C.byref(w), C.byref(h)),
          os.fspath(path))
    return out

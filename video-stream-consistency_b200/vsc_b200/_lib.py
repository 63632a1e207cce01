"""Loads libvsc_b200.so and declares the prototypes of include/vsc/vsc.h for ctypes.

Fails loudly when the library is missing (build it with __graft_entry__.build() or
``python video-stream-consistency_b200/build.py``): there is no fallback implementation."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VSC_B200_LIB: another build of the SAME library (profiles/build_variant.py compiles csrc/ with extra -D flags for
# A/B timing of kernel variants); it must export every symbol below like the default build.
LIB_PATH = os.environ.get("VSC_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libvsc_b200.so")


class VscError(RuntimeError):
    pass


_p, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol include/vsc/vsc.h declares
PROTOTYPES = {
    "vsc_version": (_i, []),
    "vsc_error_string": (C.c_char_p, [_i]),
    "vsc_launch_count": (C.c_uint64, []),
    "vsc_correlation_f32": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "vsc_set_correlation_mode": (_i, [_i]),
    "vsc_warp_nchw_f32": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "vsc_warp_hwc3": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "vsc_adap_comb": (_i, [_p] * 9 + [_f, _i, _i, _p]),
    "vsc_consist_wt": (_i, [_p, _p, _p, _f, _f, _i, _i, _p]),
    "vsc_bilinear": (_i, [_p, _i, _i, _i, _p, _i, _i, _i, _p]),
    "vsc_consist_solve_workspace_bytes": (_sz, [_i, _i]),
    "vsc_consist_solve": (_i, [_p, _p, _p, _i, _f, _f, _p, _i, _i, _p, _sz, _p]),
    "vsc_set_solver_mode": (_i, [_i]),
    "vsc_set_warp_mode": (_i, [_i]),
    "vsc_set_stage_a_mode": (_i, [_i]),
    "vsc_rgba8_to_f32x3": (_i, [_p, _p, _i, _i, _p]),
    "vsc_f32x3_to_rgba8": (_i, [_p, _p, _i, _i, _p]),
    "vsc_rgba8_scale_nearest": (_i, [_p, _i, _i, _p, _i, _i, _p]),
    "vsc_stage_a_fused": (_i, [_p] * 9 + [_i, _f, _f, _f, _p, _p, _p, _i, _i, _p]),
    "vsc_hyper_params_default": (None, [_p]),
    "vsc_frame_solve_workspace_bytes": (_sz, [_i, _i, _i]),
    "vsc_frame_solve": (_i, [_p, _p, _p, _p, _p, _i, _i, _p, _sz, _p]),
    "vsc_frame_stabilize_workspace_bytes": (_sz, [_i, _i, _i]),
    "vsc_frame_stabilize": (_i, [_p] * 9 + [_i, _p, _p, _i, _i, _p, _sz, _p]),
    "vsc_stabilizer_create": (_i, [_p, _i, _i, _i]),
    "vsc_stabilizer_create_batched": (_i, [_p, _i, _i, _i, _i]),
    "vsc_stabilizer_batch_size": (_i, [_p]),
    "vsc_stabilizer_window_count": (_i, [_p]),
    "vsc_stabilizer_destroy": (None, [_p]),
    "vsc_stabilizer_hyper_params": (_p, [_p]),
    "vsc_stabilizer_push_frame": (_i, [_p, _p, _p]),
    "vsc_stabilizer_step": (_i, [_p, _p, _p, _p]),
    "vsc_stabilizer_step_lowres_flow": (_i, [_p, _p, _p, _i, _i, _p]),
    "vsc_stabilizer_step_host_flow": (_i, [_p, _p, _p, _i, _i, _p]),
    "vsc_stabilizer_step_flow_files": (_i, [_p, C.c_char_p, _i, _p]),
    "vsc_stabilizer_prefetch_flow_files": (_i, [_p, C.c_char_p, _i]),
    "vsc_stabilizer_flow_input": (_i, [_p, _i, _p, _i, _i]),
    "vsc_stabilizer_sync": (_i, [_p]),
    "vsc_stabilizer_wait_uploads": (_i, [_p]),
    "vsc_stabilizer_last_output_dev": (_p, [_p]),
    "vsc_stabilizer_copy_last_output": (_i, [_p, _p]),
    "vsc_stabilizer_compute_stream": (_p, [_p]),
    "vsc_stabilizer_reset": (_i, [_p]),
    "vsc_flo_read_header": (_i, [C.c_char_p, _p, _p]),
    "vsc_flo_read": (_i, [C.c_char_p, _p, _sz, _p, _p]),
    "vsc_flo_frame_path": (_i, [C.c_char_p, _i, _i, C.c_char_p, _sz]),
    "vsc_host_alloc": (_i, [_p, _sz]),
    "vsc_host_free": (_i, [_p]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VscError(
                f"{LIB_PATH} not found: build the CUDA library first (__graft_entry__.build()); "
                "there is no CPU or PyTorch fallback for this path")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(h, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().vsc_error_string(int(rc))
        raise VscError(f"vsc error {rc}: {msg.decode() if msg else '?'}" + (f" {what}" if what else ""))

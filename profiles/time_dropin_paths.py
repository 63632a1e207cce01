"""Host-clock time per stabilized frame of the three ways the reference application can reach the library, 1080p and
4K, flows at frame resolution already on the device, every path ending with the 8-bit frame in host memory:

  shim        the UNCHANGED videostabilizer.cpp call sequence through the flowconsistency.cuh / GPUImage drop-in
              (host/stabilization/vsc_flowconsistency.cpp): 7 + ~30 synchronous calls and 8 device copies per frame;
  drop-in     class VideoStabilizer with the product's member functions (host/stabilization/vsc_videostabilizer.cpp):
              doOneStep = one fused vsc_frame_stabilize call + conversion + one synchronisation.  Measured inside
              tests/cxx/videostab_driver.cpp as doOneStep minus FlowModel::run, loadFrame and outputFrame (the application's);
  pipeline    the vsc_stabilizer object (pinned asynchronous uploads / downloads overlapping the solve), the path
              bench.py's e2e number measures.

    python profiles/time_dropin_paths.py > gpurun_out/time_dropin_paths.txt
"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "video-stream-consistency_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import synth  # noqa: E402
import vsc_b200 as V  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
B = os.path.join(ROOT, "tests", "cxx", "_build")
shim = C.CDLL(os.path.join(B, "libvsc_stab_shim_test.so"))
shim.vsc_shim_create.restype = C.c_void_p
shim.vsc_shim_create.argtypes = [C.c_int] * 4
shim.vsc_shim_push.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
shim.vsc_shim_time_steps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
shim.vsc_shim_destroy.argtypes = [C.c_void_p]
vs = C.CDLL(os.path.join(B, "libvsc_videostab_test.so"))
vs.vsc_vs_test_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_int, C.c_float]

for (W, H, T) in ((1920, 1080, 24), (3840, 2160, 12)):
    o8, p8 = synth.frames(W, H, T, seed=5)
    ff, fb = synth.flows(W, H, 3)
    # --- shim path
    h = shim.vsc_shim_create(W, H, 3, 2)
    for t in range(3):
        assert shim.vsc_shim_push(h, o8[t].ctypes.data_as(C.c_void_p), p8[t].ctypes.data_as(C.c_void_p)) == 0
    ms = C.c_double(0)
    assert shim.vsc_shim_time_steps(h, ff.ctypes.data_as(C.c_void_p), fb.ctypes.data_as(C.c_void_p), 2, 150, C.byref(ms)) == 0
    assert shim.vsc_shim_time_steps(h, ff.ctypes.data_as(C.c_void_p), fb.ctypes.data_as(C.c_void_p), 8, 150, C.byref(ms)) == 0
    shim.vsc_shim_destroy(h)
    t_shim = ms.value
    # --- drop-in path (the stand-in flow network runs at frame resolution; its time is subtracted)
    os.environ.pop("FLOWDOWNSCALE", None)
    orig, proc = np.ascontiguousarray(np.stack(o8)), np.ascontiguousarray(np.stack(p8))
    outs, have = np.zeros((T, H, W, 4), np.uint8), np.zeros(T, np.int32)
    tm = (C.c_double * 4)()
    vs.vsc_vs_test_timing(tm)
    steps = vs.vsc_vs_test_run(W, H, T, 1, orig.ctypes.data_as(C.c_void_p), proc.ctypes.data_as(C.c_void_p),
                               outs.ctypes.data_as(C.c_void_p), have.ctypes.data_as(C.c_void_p), 0, 0.0)
    vs.vsc_vs_test_timing(tm)
    t_drop = (tm[0] - tm[1] - tm[2] - tm[3]) / steps
    t_load = tm[2] / steps
    # --- pipeline object, pinned frames, device flows
    st = V.Stabilizer(W, H, 3)
    dff, dfb = torch.from_numpy(ff).to(dev), torch.from_numpy(fb).to(dev)
    pin = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b in zip(o8, p8)]
    out = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
    for t in range(3):
        st.push_frame(*pin[t])
    for rep in range(2):
        n = 0
        st.sync()
        t0 = time.perf_counter()
        for t in range(1, T - 2):
            st.step(dff, dfb, out)
            st.push_frame(*pin[(t + 2) % T])
            n += 1
        st.sync()
        t_pipe = (time.perf_counter() - t0) * 1e3 / n
    st.close()
    print(f"{W}x{H}: shim (unchanged videostabilizer.cpp) {t_shim:7.2f} ms/frame = {1e3 / t_shim:6.1f} fps | "
          f"drop-in VideoStabilizer::doOneStep {t_drop:6.2f} ms = {1e3 / t_drop:6.1f} fps (its loadFrame: {t_load:.2f} ms) | "
          f"vsc_stabilizer pipeline {t_pipe:6.2f} ms = {1e3 / t_pipe:6.1f} fps", flush=True)

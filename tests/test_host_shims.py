"""The C++ drop-in shims (video-stream-consistency_b200/host/) exercised the way their hosts would:

  * ORT plug-in: RegisterCustomOps -> domain "custom" -> op lookup by name + execution provider -> CreateKernel
    from node attributes -> KernelCompute on ORT's stream, with the stand-in ORT API (standins/ort) playing the
    session; registration is compared with the REFERENCE's own RegisterCustomOps (custom_ops.cpp compiled
    unmodified into oracle/_ref) run against the same stand-in;
  * stabilization: GPUImage + the six flowconsistency.cuh functions (compiled against the reference's unmodified
    headers) driven through the doOneStep call sequence from RGBA host frames.
"""
import ctypes as C
import os

import numpy as np
import pytest

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORT_SO = os.path.join(ROOT, "tests", "cxx", "_build", "libvsc_ort_shim_test.so")
STAB_SO = os.path.join(ROOT, "tests", "cxx", "_build", "libvsc_stab_shim_test.so")
F32P = C.POINTER(C.c_float)


def _registry(lib, fn):
    buf = C.create_string_buffer(4096)
    n = getattr(lib, fn)(buf, C.c_size_t(4096))
    rows = [tuple(r.split("|")) for r in buf.value.decode().split(";") if r]
    assert n == len(rows)
    return rows


@pytest.fixture(scope="module")
def ort():
    if not os.path.exists(ORT_SO):
        pytest.fail(f"{ORT_SO} missing: run __graft_entry__.build()")
    lib = C.CDLL(ORT_SO)
    lib.vsc_ort_test_last_error.restype = C.c_char_p
    return lib


def test_registration_matches_the_reference(ort, O):
    ours = _registry(ort, "vsc_ort_test_registry")
    # (domain, op, provider, n_inputs, n_outputs, input type, output type); 1 == ONNX FLOAT
    assert ("custom", "Correlation", "CUDAExecutionProvider", "2", "1", "1", "1") in ours
    assert ("custom", "Warp", "CUDAExecutionProvider", "2", "1", "1", "1") in ours
    assert len(ours) == 2  # no CPU-provider kernels: this library has no CPU path
    if O.ref_cpu_available():
        ref = _registry(O.ref_cpu(), "vsc_ref_cpu_registry")
        assert len(ref) == 4
        assert set(ours) <= set(ref)
        assert {r for r in ref if r[2] == "CUDAExecutionProvider"} == set(ours)


def test_missing_attributes_throw_like_the_reference(ort):
    """correlation.h:19-31: the kernel constructor throws std::runtime_error when an attribute is absent."""
    z = C.c_void_p(0)
    dims = (C.c_int64 * 8)()
    rank = C.c_int(0)
    args = (z, z, z, C.c_size_t(0), C.c_int64(1), C.c_int64(1), C.c_int64(1), C.c_int64(1), C.c_int64(4), C.c_int64(0))
    assert ort.vsc_ort_test_correlation(*args, 0, 1, z, dims, C.byref(rank)) == 1
    assert b"legacy" in ort.vsc_ort_test_last_error()
    assert ort.vsc_ort_test_correlation(*args, 1, 0, z, dims, C.byref(rank)) == 1
    assert b"max_displacement" in ort.vsc_ort_test_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize("legacy", [0, 1])
def test_ort_correlation_through_registration(ort, O, dev, legacy):
    import torch

    N, Cc, H, W = 2, 24, 18, 30
    a, b = synth.features(N, Cc, H, W, 1), synth.features(N, Cc, H, W, 2)
    ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    out = torch.zeros((N, 81, H, W), device=dev)
    dims = (C.c_int64 * 8)()
    rank = C.c_int(0)
    s = torch.cuda.Stream()
    torch.cuda.synchronize()
    rc = ort.vsc_ort_test_correlation(C.c_void_p(ta.data_ptr()), C.c_void_p(tb.data_ptr()), C.c_void_p(out.data_ptr()),
                                      C.c_size_t(out.numel() * 4), C.c_int64(N), C.c_int64(Cc), C.c_int64(H),
                                      C.c_int64(W), C.c_int64(4), C.c_int64(legacy), 1, 1, C.c_void_p(s.cuda_stream),
                                      dims, C.byref(rank))
    assert rc == 0, ort.vsc_ort_test_last_error()
    s.synchronize()
    # output shape the op asks ORT for (correlation_cuda.cc:69-76)
    assert list(dims[: rank.value]) == ([N, 81, H, W] if legacy else [N, 9, 9, H, W])
    ref = O.correlation(a, b, legacy=bool(legacy)).reshape(N, 81, H, W)
    got = out.cpu().numpy()
    assert np.abs(got - ref).max() / np.abs(ref).max() <= 1e-4


@pytest.mark.gpu
def test_ort_warp_through_registration(ort, O, dev):
    import torch

    N, Cc, H, W = 2, 16, 18, 30
    x, f = synth.features(N, Cc, H, W, 3), synth.op_flow(N, H, W, 4, 3.0)
    tx, tf = torch.from_numpy(x).to(dev), torch.from_numpy(f).to(dev)
    out = torch.zeros_like(tx)
    torch.cuda.synchronize()
    rc = ort.vsc_ort_test_warp(C.c_void_p(tx.data_ptr()), C.c_void_p(tf.data_ptr()), C.c_void_p(out.data_ptr()),
                               C.c_size_t(out.numel() * 4), C.c_int64(N), C.c_int64(Cc), C.c_int64(H), C.c_int64(W),
                               C.c_int64(2), C.c_void_p(0))
    assert rc == 0, ort.vsc_ort_test_last_error()
    torch.cuda.synchronize()
    ref = O.warp_nchw(x, f)
    assert np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max() <= 1e-4
    # a 3-channel flow is rejected with an exception, not a crash
    rc = ort.vsc_ort_test_warp(C.c_void_p(tx.data_ptr()), C.c_void_p(tf.data_ptr()), C.c_void_p(out.data_ptr()),
                               C.c_size_t(out.numel() * 4), C.c_int64(N), C.c_int64(Cc), C.c_int64(H), C.c_int64(W),
                               C.c_int64(3), C.c_void_p(0))
    assert rc == 1 and b"flow" in ort.vsc_ort_test_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize("W,H", [(64, 48), (45, 37), (1920, 1080), (3840, 2160)])
def test_stabilization_shim_sequence(O, dev, W, H):
    """GPUImage + flowconsistency.cuh drop-in: preload + 2 doOneStep calls from RGBA host frames vs the oracle.
    At 1080p / 4K the D2D copyFrom calls take tens of microseconds: a get_* kernel that were not ordered behind
    them (copies on the legacy stream, kernels on a non-blocking stream) would read half-copied images."""
    if not os.path.exists(STAB_SO):
        pytest.skip("stabilization shim test library not built (needs the reference headers at build time)")
    lib = C.CDLL(STAB_SO)
    lib.vsc_shim_create.restype = C.c_void_p
    o8, p8 = synth.frames(W, H, 4, seed=91)
    ff, fb = synth.flows(W, H, 3)
    h = C.c_void_p(lib.vsc_shim_create(W, H, 3, 2))
    assert h
    for t in range(3):
        assert lib.vsc_shim_push(h, o8[t].ctypes.data_as(C.c_void_p), p8[t].ctypes.data_as(C.c_void_p)) == 0
    of = [O.rgba8_to_f32x3(x) for x in o8]
    pf = [O.rgba8_to_f32x3(x) for x in p8]
    last = pf[2]
    O.use_all_cores()
    for t in (1, 2) if W < 3840 else (1,):
        rgba = np.zeros((H, W, 4), np.uint8)
        cons = np.zeros((H, W, 3), np.float32)
        rc = lib.vsc_shim_step(h, ff.ctypes.data_as(F32P), fb.ctypes.data_as(F32P), C.c_float(6800.0),
                               C.c_float(6800.0), C.c_float(2.0), 150, C.c_float(0.15), C.c_float(0.15),
                               rgba.ctypes.data_as(C.c_void_p), cons.ctypes.data_as(F32P))
        assert rc == 0
        ref_f, ref8 = O.do_one_step(of[t - 1], of[t], of[t + 1], pf[t - 1], pf[t], pf[t + 1], last, ff, fb)
        last = ref_f
        assert np.abs(cons - ref_f).max() <= 3e-5
        assert np.abs(rgba.astype(np.int32) - ref8.astype(np.int32)).max() <= 1
        if t == 1 and W < 3840:
            assert lib.vsc_shim_push(h, o8[3].ctypes.data_as(C.c_void_p), p8[3].ctypes.data_as(C.c_void_p)) == 0
    lib.vsc_shim_destroy(h)


# ---------------------------------------------------------------- flow session (ORT IoBinding on device buffers)
FS_SO = os.path.join(ROOT, "tests", "cxx", "_build", "libvsc_flow_session_test.so")


@pytest.fixture(scope="module")
def fs():
    if not os.path.exists(FS_SO):
        pytest.fail(f"{FS_SO} missing: run __graft_entry__.build()")
    lib = C.CDLL(FS_SO)
    lib.vsc_fs_test_last_error.restype = C.c_char_p
    lib.vsc_fs_test_create.restype = C.c_void_p
    lib.vsc_fs_test_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
    lib.vsc_fs_test_destroy.argtypes = [C.c_void_p]
    lib.vsc_fs_test_stabilizer.restype = C.c_void_p
    lib.vsc_fs_test_stabilizer.argtypes = [C.c_void_p]
    lib.vsc_fs_test_step.argtypes = [C.c_void_p, C.c_void_p]
    lib.vsc_fs_test_flow.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return lib


def _fs_counters(fs):
    c = (C.c_long * 6)()
    fs.vsc_fs_test_counters(c)
    return dict(zip(("runs", "provider_syncs", "sessions", "bound", "graph_calls", "saw_domain"), c))


def test_flow_session_library_exports_the_session_entry_points(fs):
    """CPU: the flow-session test library (product source + stand-in ORT) loads and exports its driver symbols;
    the product class was compiled against the same Ort:: names the real headers declare."""
    for name in ("vsc_fs_test_create", "vsc_fs_test_step", "vsc_fs_test_flow", "vsc_fs_test_counters",
                 "RegisterCustomOps"):
        assert hasattr(fs, name)


@pytest.mark.gpu
@pytest.mark.parametrize("batched", [0, 1], ids=["two-runs", "batch-2"])
@pytest.mark.parametrize("W,H,scale", [(96, 64, 1), (128, 72, 2), (90, 50, 2)])
def test_flow_session_runs_on_device_buffers(fs, V, dev, W, H, scale, batched):
    """FlowModel::run + doOneStep through VscFlowSession: frames pushed from host memory once, the network inputs
    written on the device (nearest-neighbour scale for FLOWDOWNSCALE), a persistent IoBinding, enqueue-only runs
    on the stabilizer's stream, flows consumed in place.  The stand-in graph's flow and the stabilized frames must
    equal the same computation done step by step through the Python binding, bit for bit.  batch-2: both directions
    of a frame as ONE run on [2,H,W,4] inputs (frame1 = [cur, next], frame2 = [next, cur])."""
    import torch

    netW, netH = W // scale, H // scale
    T = 6
    o8, p8 = synth.frames(W, H, T, seed=77)
    before = _fs_counters(fs)
    rig = fs.vsc_fs_test_create(W, H, netW, netH, None, batched)
    assert rig, fs.vsc_fs_test_last_error()
    st = C.c_void_p(fs.vsc_fs_test_stabilizer(rig))
    L = V.lib()
    ref = V.Stabilizer(W, H, 3)
    outs, refs = [], []
    try:
        for t in range(3):
            assert L.vsc_stabilizer_push_frame(st, o8[t].ctypes.data_as(C.c_void_p), p8[t].ctypes.data_as(C.c_void_p)) == 0
            ref.push_frame(o8[t], p8[t])

        def graph(first, second):   # what the stand-in model computes, via the Python binding
            a = V.image_to_gpu(V.rgba8_scale_nearest(torch.from_numpy(first).to(dev), netW, netH))
            b = V.image_to_gpu(V.rgba8_scale_nearest(torch.from_numpy(second).to(dev), netW, netH))
            return V.get_warp_result(a, b)

        # one direction on its own, copied back: inputs, binding and output slot are the right ones
        got = np.empty((netH, netW, 3), np.float32)
        if batched:   # single directions are not available on a batched session
            assert fs.vsc_fs_test_flow(rig, 1, 2, 0, got.ctypes.data_as(C.c_void_p)) == 1
            assert b"one batch" in fs.vsc_fs_test_last_error()
        else:
            assert fs.vsc_fs_test_flow(rig, 1, 2, 0, got.ctypes.data_as(C.c_void_p)) == 0, fs.vsc_fs_test_last_error()
            assert np.array_equal(got, graph(o8[1], o8[2]).cpu().numpy())
            assert fs.vsc_fs_test_flow(rig, 2, 1, 1, got.ctypes.data_as(C.c_void_p)) == 0, fs.vsc_fs_test_last_error()
            assert np.array_equal(got, graph(o8[2], o8[1]).cpu().numpy())

        for t in range(1, T - 1):
            out = np.zeros((H, W, 4), np.uint8)
            assert fs.vsc_fs_test_step(rig, out.ctypes.data_as(C.c_void_p)) == 0, fs.vsc_fs_test_last_error()
            assert L.vsc_stabilizer_sync(st) == 0
            outs.append(out)
            r = np.zeros((H, W, 4), np.uint8)
            ref.step(graph(o8[t], o8[t + 1]), graph(o8[t + 1], o8[t]), r)
            ref.sync()
            refs.append(r)
            if t + 2 < T:
                assert L.vsc_stabilizer_push_frame(st, o8[t + 2].ctypes.data_as(C.c_void_p),
                                                   p8[t + 2].ctypes.data_as(C.c_void_p)) == 0
                ref.push_frame(o8[t + 2], p8[t + 2])
    finally:
        fs.vsc_fs_test_destroy(rig)
        ref.close()
    for a, b in zip(outs, refs):
        assert np.array_equal(a, b)
    assert any(np.any(a[..., :3] != p8[i + 1][..., :3]) for i, a in enumerate(outs))   # the step did something
    c = _fs_counters(fs)
    steps = T - 2
    assert c["sessions"] - before["sessions"] == 1
    if batched:
        assert c["bound"] - before["bound"] == 3                    # one persistent binding of [2,H,W,*] tensors
        assert c["runs"] - before["runs"] == steps                  # ONE run per frame
    else:
        assert c["bound"] - before["bound"] == 6                    # 2 persistent bindings x 3 tensors, built once
        assert c["runs"] - before["runs"] == 2 + 2 * steps          # two directions per frame
    assert c["provider_syncs"] - before["provider_syncs"] == 0      # every Run was enqueue-only
    assert c["saw_domain"] == 1                                     # RegisterCustomOps reached the session options


@pytest.mark.gpu
def test_flow_session_unknown_model_throws(fs, V, dev):
    """InferenceModelVariant::createSession rethrows ORT's load failure (:175-183); so does the flow session."""
    assert not fs.vsc_fs_test_create(64, 48, 64, 48, b"models/does-not-exist.onnx", 0)
    assert b"does-not-exist" in fs.vsc_fs_test_last_error()

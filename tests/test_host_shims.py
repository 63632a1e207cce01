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
    lib.vsc_fs_test_create_b.restype = C.c_void_p
    lib.vsc_fs_test_create_b.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int]
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
@pytest.mark.parametrize("batch", [2, 3])
def test_flow_session_flow_batches(fs, V, dev, batch):
    """`-b batchSize` through VscFlowSession: [batchSize,H,W,4] session tensors filled on the device from window
    frames 1..batchSize / 2..batchSize+1, ONE run per direction every batchSize frames, the frames in between index
    into the two flow batches (videostabilizer.cpp:176-179,269-273; flowmodel.cpp:137-165).  Stabilized frames equal a
    batch-1 Python-binding run that starts its recurrence from the same frame, bit for bit."""
    import torch

    W, H, netW, netH, T = 96, 64, 48, 32, 10
    o8, p8 = synth.frames(W, H, T, seed=81)
    before = _fs_counters(fs)
    rig = fs.vsc_fs_test_create_b(W, H, netW, netH, None, 0, batch)
    assert rig, fs.vsc_fs_test_last_error()
    st = C.c_void_p(fs.vsc_fs_test_stabilizer(rig))
    L = V.lib()
    ref = V.Stabilizer(W, H, 3, batch_size=batch)
    outs, refs = [], []

    def graph(first, second):   # what the stand-in model computes, via the Python binding
        a = V.image_to_gpu(V.rgba8_scale_nearest(torch.from_numpy(first).to(dev), netW, netH))
        b = V.image_to_gpu(V.rgba8_scale_nearest(torch.from_numpy(second).to(dev), netW, netH))
        return V.get_warp_result(a, b)

    nsteps = 2 * batch   # two whole flow batches
    try:
        for t in range(2 + batch):
            assert L.vsc_stabilizer_push_frame(st, o8[t].ctypes.data_as(C.c_void_p), p8[t].ctypes.data_as(C.c_void_p)) == 0
            ref.push_frame(o8[t], p8[t])
        for t in range(1, 1 + nsteps):
            out = np.zeros((H, W, 4), np.uint8)
            assert fs.vsc_fs_test_step(rig, out.ctypes.data_as(C.c_void_p)) == 0, fs.vsc_fs_test_last_error()
            assert L.vsc_stabilizer_sync(st) == 0
            outs.append(out)
            r = np.zeros((H, W, 4), np.uint8)
            ref.step(graph(o8[t], o8[t + 1]), graph(o8[t + 1], o8[t]), r)
            ref.sync()
            refs.append(r)
            nxt = t + 1 + batch
            if nxt < T:
                assert L.vsc_stabilizer_push_frame(st, o8[nxt].ctypes.data_as(C.c_void_p),
                                                   p8[nxt].ctypes.data_as(C.c_void_p)) == 0
                ref.push_frame(o8[nxt], p8[nxt])
    finally:
        fs.vsc_fs_test_destroy(rig)
        ref.close()
    for a, b in zip(outs, refs):
        assert np.array_equal(a, b)
    c = _fs_counters(fs)
    assert c["runs"] - before["runs"] == 2 * 2                      # two directions per flow batch, two batches
    assert c["provider_syncs"] - before["provider_syncs"] == 0


@pytest.mark.gpu
def test_flow_session_unknown_model_throws(fs, V, dev):
    """InferenceModelVariant::createSession rethrows ORT's load failure (:175-183); so does the flow session."""
    assert not fs.vsc_fs_test_create(64, 48, 64, 48, b"models/does-not-exist.onnx", 0)
    assert b"does-not-exist" in fs.vsc_fs_test_last_error()


# ---------------------------------------------------------------- VideoStabilizer drop-in (host/stabilization/vsc_videostabilizer.cpp)
@pytest.fixture(scope="module")
def vs():
    path = os.path.join(os.path.dirname(__file__), "cxx", "_build", "libvsc_videostab_test.so")
    if not os.path.exists(path):
        pytest.skip("libvsc_videostab_test.so not built (needs the reference checkout at build time)")
    lib = C.CDLL(path)
    lib.vsc_vs_test_last_error.restype = C.c_char_p
    lib.vsc_vs_test_flow_runs.restype = C.c_long
    lib.vsc_vs_test_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_float]
    return lib


def test_videostabilizer_dropin_library_exports(vs):
    """CPU: the drop-in TU compiled against the reference's unmodified videostabilizer.h (+ flowmodel.h, imagehelpers.h,
    the inference headers) and linked with the reference's own imagehelpers.cpp; the library loads and exports its
    driver."""
    assert hasattr(vs, "vsc_vs_test_run") and hasattr(vs, "vsc_vs_test_flow_runs")


@pytest.mark.gpu
@pytest.mark.parametrize("batch,scale", [(1, 1), (1, 2), (3, 2), (2, 1)])
def test_videostabilizer_dropin_sequence(vs, V, O, dev, batch, scale, monkeypatch):
    """The reference's class VideoStabilizer with the product's member functions: preloadProcessedFrames, doOneStep
    until the stream ends, outputFinalFrames -- driven like StreamStabilizer::stabilizeAll, with `-b batch` flow batches
    and FLOWDOWNSCALE.  Every emitted frame equals the same clip run through the pipeline object of the Python binding
    (same kernels underneath) bit for bit; batch 1 at full flow resolution is also held against the oracle."""
    import torch

    W, H, T = 96, 64, 9
    netW, netH = round(W / scale), round(H / scale)
    o8, p8 = synth.frames(W, H, T, seed=83, mismatch=0.2)
    if scale > 1:
        monkeypatch.setenv("FLOWDOWNSCALE", str(scale))
    else:
        monkeypatch.delenv("FLOWDOWNSCALE", raising=False)
    orig = np.ascontiguousarray(np.stack(o8))
    proc = np.ascontiguousarray(np.stack(p8))
    outs = np.zeros((T, H, W, 4), np.uint8)
    have = np.zeros(T, np.int32)
    runs0 = vs.vsc_vs_test_flow_runs()
    steps = vs.vsc_vs_test_run(W, H, T, batch, orig.ctypes.data_as(C.c_void_p), proc.ctypes.data_as(C.c_void_p),
                               outs.ctypes.data_as(C.c_void_p), have.ctypes.data_as(C.c_void_p), 0, 0.0)
    assert steps > 0, vs.vsc_vs_test_last_error()

    def graph(first, second):   # the driver's stand-in network, via the Python binding
        a = V.image_to_gpu(V.rgba8_scale_nearest(torch.from_numpy(first).to(dev), netW, netH))
        b = V.image_to_gpu(V.rgba8_scale_nearest(torch.from_numpy(second).to(dev), netW, netH))
        return V.get_warp_result(a, b)

    # the same control flow on the pipeline object (videostabilizer.cpp:136-164,167-279)
    ref = V.Stabilizer(W, H, 3, batch_size=batch)
    exp, exp_have = np.zeros_like(outs), np.zeros(T, np.int32)
    win = list(range(2 + batch))
    nxt = 2 + batch
    for t in win:
        ref.push_frame(o8[t], p8[t])
    for t in (0, 1):   # j <= k
        exp[t], exp_have[t] = p8[t], exp_have[t] + 1
    exp[:2, ..., 3] = 1   # gpuToImage writes alpha = 1 (gpuimage.cu:66)
    i, nsteps, flow_batches = 1, 0, 0
    ff = fb = None
    ora = []
    while True:
        if (i - 1) % batch == 0:
            ff = [graph(o8[win[1 + b]], o8[win[2 + b]]) for b in range(batch)]
            fb = [graph(o8[win[2 + b]], o8[win[1 + b]]) for b in range(batch)]
            flow_batches += 1
        out = np.zeros((H, W, 4), np.uint8)
        bi = (i - 1) % batch
        ora.append((win[0], win[1], win[2], ff[bi], fb[bi]))
        ref.step(ff[bi], fb[bi], out)
        ref.sync()
        exp[i], exp_have[i] = out, exp_have[i] + 1
        nsteps += 1
        win.pop(0)
        if nxt < T:
            ref.push_frame(o8[nxt], p8[nxt])
            win.append(nxt)
            nxt += 1
        else:
            exp[i + 1], exp_have[i + 1] = p8[win[1]], exp_have[i + 1] + 1   # outputFinalFrames
            exp[i + 1, ..., 3] = 1
            break
        i += 1
    ref.close()
    assert steps == nsteps
    assert np.array_equal(have, exp_have)
    assert vs.vsc_vs_test_flow_runs() - runs0 == 2 * flow_batches   # two directions per flow batch
    for t in range(T):
        if exp_have[t]:
            assert np.array_equal(outs[t], exp[t]), t
    if batch == 1 and scale == 1:   # and against the oracle (first three steps)
        of = [O.rgba8_to_f32x3(x) for x in o8]
        pf = [O.rgba8_to_f32x3(x) for x in p8]
        last = pf[2]
        for step_i, (a, b, c, f1, f2) in enumerate(ora[:3]):
            last, rgba = O.do_one_step(of[a], of[b], of[c], pf[a], pf[b], pf[c], last, f1.cpu().numpy(), f2.cpu().numpy())
            d = np.abs(outs[step_i + 1].astype(np.int32) - rgba.astype(np.int32))
            assert d.max() <= 1, (step_i, d.max())

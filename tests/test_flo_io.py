"""Precomputed-flow ingestion: the product's .flo reader (C ABI vsc_flo_read, the
ReadFlowFile drop-in shim, the Python binding) against the REFERENCE's own ReadFlowFile (flowIO.cpp compiled
unmodified into oracle/_ref) -- same values, same accepted / rejected files, same exception text.  Host-side
code: runs without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAB_SO = os.path.join(ROOT, "tests", "cxx", "_build", "libvsc_stab_shim_test.so")


def write_flo(path, flow2, tag=b"PIEH", dims=None, extra=b"", drop=0):
    h, w, _ = flow2.shape
    w, h = dims if dims else (w, h)
    payload = np.ascontiguousarray(flow2, np.float32).tobytes()
    with open(path, "wb") as f:
        f.write(tag)
        f.write(np.array([w, h], np.int32).tobytes())
        f.write(payload[:len(payload) - drop] + extra)


def ref_read(O, path, cap):
    L = O.ref_cpu()
    buf = np.zeros(cap, np.float32)
    w, h = C.c_int(0), C.c_int(0)
    msg = C.create_string_buffer(1024)
    rc = L.vsc_ref_read_flo_msg(os.fsencode(path), buf.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(w),
                                C.byref(h), msg, C.c_size_t(1024))
    return rc, buf, w.value, h.value, msg.value.decode()


def shim_read(path, cap):
    if not os.path.exists(STAB_SO):
        pytest.fail(f"{STAB_SO} missing: run __graft_entry__.build()")
    L = C.CDLL(STAB_SO)
    buf = np.zeros(cap, np.float32)
    w, h = C.c_int(0), C.c_int(0)
    msg = C.create_string_buffer(1024)
    rc = L.vsc_shim_read_flo(os.fsencode(path), buf.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(w),
                             C.byref(h), msg, C.c_size_t(1024))
    return rc, buf, w.value, h.value, msg.value.decode()


@pytest.mark.parametrize("W,H", [(1, 1), (7, 3), (64, 48), (333, 21)])
def test_flo_read_matches_the_reference_reader(V, O, tmp_path, W, H):
    ff, _ = synth.flows(W, H, 2)
    ff[0, 0] = (1e10, -1e10)          # "unknown flow" marker values pass through unchanged
    path = str(tmp_path / "frame_000001.flo")
    write_flo(path, ff)
    got = V.flo_read(path).numpy()
    assert got.shape == (H, W, 2) and np.array_equal(got, ff)
    if O.ref_cpu_available():
        rc, buf, w, h, _ = ref_read(O, path, W * H * 2)
        assert rc == 0 and (w, h) == (W, H)
        assert np.array_equal(buf.reshape(H, W, 2), got)
    rc, buf, w, h, _ = shim_read(path, W * H * 2)
    assert rc == 0 and (w, h) == (W, H) and np.array_equal(buf.reshape(H, W, 2), ff)
    # header-only query, and a too-small destination is refused before anything is written
    L = V.lib()
    cw, ch = C.c_int(0), C.c_int(0)
    assert L.vsc_flo_read_header(os.fsencode(path), C.byref(cw), C.byref(ch)) == 0 and (cw.value, ch.value) == (W, H)
    small = np.full(max(W * H * 2 - 1, 1), 7.0, np.float32)
    if W * H * 2 > 1:
        assert L.vsc_flo_read(os.fsencode(path), small.ctypes.data_as(C.c_void_p), C.c_size_t(small.size),
                              C.byref(cw), C.byref(ch)) == -2
        assert np.all(small == 7.0)


CASES = {
    "missing": None,
    "short_header": dict(raw=b"PIEH\x04\x00"),
    "wrong_tag": dict(tag=b"HEIP"),
    "width_zero": dict(dims=(0, 3)),
    "width_huge": dict(dims=(100000, 3)),
    "height_zero": dict(dims=(5, 0)),
    "height_negative": dict(dims=(5, -2)),
    "too_short": dict(drop=4),
    "too_long": dict(extra=b"\x00"),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_flo_errors_match_the_reference(V, O, tmp_path, case):
    """every rejection of the reference (flowIO.cpp:34-74) is a rejection here, with the same exception text"""
    spec = CASES[case]
    path = str(tmp_path / f"{case}.flo")
    ff, _ = synth.flows(5, 3, 2)
    if spec is not None:
        if "raw" in spec:
            with open(path, "wb") as f:
                f.write(spec["raw"])
        else:
            write_flo(path, ff, **spec)
    with pytest.raises(V.VscError) as ei:
        V.flo_read(path)
    rc, _, _, _, shim_msg = shim_read(path, 1 << 16)
    assert rc == 1 and shim_msg.startswith("ReadFlowFile: ") and shim_msg.endswith(path)
    assert shim_msg in str(ei.value)
    if O.ref_cpu_available():
        rrc, _, _, _, ref_msg = ref_read(O, path, 1 << 16)
        assert rrc == 1
        assert shim_msg == ref_msg


def test_flo_frame_paths(V):
    """FileStabilizer::retrieveOpticalFlow naming (stabilizefiles.cpp:99-101,141-144)"""
    assert V.flo_frame_path("flows", 7) == "flows/frame_000007.flo"
    assert V.flo_frame_path("flows/", 7, backward=True) == "flows/frame_000007_bwd.flo"
    assert V.flo_frame_path("/a/b", 1234567) == "/a/b/frame_1234567.flo"
    assert V.flo_frame_path("", 0) == "frame_000000.flo"
    buf = C.create_string_buffer(8)
    assert V.lib().vsc_flo_frame_path(b"a-long-directory", 1, 0, buf, 8) == -1
    assert V.lib().vsc_flo_frame_path(None, 1, 0, buf, 8) == -1
    for code, text in ((-5, "could not open"), (-10, "too short"), (-11, "too long"), (-12, "does not match")):
        assert text in V.lib().vsc_error_string(code).decode()


def test_flo_large_files_take_the_parallel_path(V, O, tmp_path):
    """payloads >= 8 MB are read by several pread threads and size-checked with fstat: same values and the same
    too-short / too-long verdicts as the reference's row loop + EOF probe"""
    W, H = 1024, 1100
    rng = np.random.default_rng(3)
    ff = rng.standard_normal((H, W, 2)).astype(np.float32)
    good, short, long_ = (str(tmp_path / n) for n in ("good.flo", "short.flo", "long.flo"))
    write_flo(good, ff)
    write_flo(short, ff, drop=4096 * 3 + 1)
    write_flo(long_, ff, extra=b"\x01\x02")
    assert np.array_equal(V.flo_read(good).numpy(), ff)
    for path, text in ((short, "too short"), (long_, "too long")):
        with pytest.raises(V.VscError, match=text):
            V.flo_read(path)
        if O.ref_cpu_available():
            rc, _, _, _, msg = ref_read(O, path, W * H * 2)
            assert rc == 1 and text in msg

"""TEST INFRASTRUCTURE: builds the host-shim test libraries (tests/cxx/_build/*.so).

The C++ drop-in shims (video-stream-consistency_b200/host/) need onnxruntime / Qt headers, which this image does
not have.  They are compile-checked -- and made runnable for tests/test_host_shims.py -- against the stand-in
headers in standins/, together with the small session-like drivers in this directory.  The stabilization shim is
compiled against the reference's own unmodified headers (gpuimage.h, flowconsistency.cuh), found on the include
path, so it is built only where the reference checkout exists (the build container); the resulting .so files
travel to the GPU box.
"""
import importlib.util
import os
import subprocess
import sys

HERE_T = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE_T))
_spec = importlib.util.spec_from_file_location("vsc_b200_build", os.path.join(ROOT, "video-stream-consistency_b200", "build.py"))
_b = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_b)
HERE, LIBDIR, LIB, NVCC, HOSTCXX = _b.HERE, _b.LIBDIR, _b.LIB, _b.NVCC, _b.HOSTCXX
REF_STAB = os.environ.get("VSC_REFERENCE_STAB", "/root/reference/src/stabilization")
TEST_SO_DIR = os.path.join(ROOT, "tests", "cxx", "_build")
CUDA_INC = os.path.join(os.path.dirname(os.path.dirname(NVCC)), "include")
CUDA_LIB = os.path.join(os.path.dirname(os.path.dirname(NVCC)), "lib64")


def build_host_shims(force: bool = False):
    os.makedirs(TEST_SO_DIR, exist_ok=True)
    built = []
    common = [HOSTCXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I", os.path.join(ROOT, "include"),
              "-Wl,-Bsymbolic", f"-Wl,-rpath,{LIBDIR}", "-L", LIBDIR]
    jobs = [(os.path.join(TEST_SO_DIR, "libvsc_ort_shim_test.so"),
             [os.path.join(HERE, "host", "ort_custom_ops", "vsc_custom_ops.cpp"),
              os.path.join(ROOT, "tests", "cxx", "ort_driver.cpp")],
             ["-I", os.path.join(ROOT, "standins", "ort")], ["-lvsc_b200"]),
            # the flow session (ORT IoBinding on device buffers) + the custom-op library it registers + its driver
            (os.path.join(TEST_SO_DIR, "libvsc_flow_session_test.so"),
             [os.path.join(HERE, "host", "inference", "vsc_flow_session.cpp"),
              os.path.join(HERE, "host", "ort_custom_ops", "vsc_custom_ops.cpp"),
              os.path.join(ROOT, "tests", "cxx", "flow_session_driver.cpp")],
             ["-I", os.path.join(ROOT, "standins", "ort"), "-I", CUDA_INC],
             ["-lvsc_b200", "-L", CUDA_LIB, f"-Wl,-rpath,{CUDA_LIB}", "-lcudart"])]
    if os.path.exists(os.path.join(REF_STAB, "flowconsistency.cuh")):
        jobs.append((os.path.join(TEST_SO_DIR, "libvsc_stab_shim_test.so"),
                     [os.path.join(HERE, "host", "stabilization", "vsc_flowconsistency.cpp"),
                      os.path.join(HERE, "host", "stabilization", "vsc_flowio.cpp"),
                      os.path.join(ROOT, "tests", "cxx", "stab_shim_driver.cpp")],
                     ["-I", REF_STAB, "-I", os.path.join(ROOT, "standins", "qt"), "-I", CUDA_INC],
                     ["-lvsc_b200", "-L", CUDA_LIB, f"-Wl,-rpath,{CUDA_LIB}", "-lcudart"]))
        # the VideoStabilizer drop-in: the product's vsc_videostabilizer.cpp against the reference's unmodified
        # videostabilizer.h (+ flowmodel.h, imagehelpers.h, the inference headers), the reference's own
        # imagehelpers.cpp compiled as it is, the GPUImage / flowconsistency shim, and a driver that plays the
        # StreamStabilizer subclass and the flow network
        jobs.append((os.path.join(TEST_SO_DIR, "libvsc_videostab_test.so"),
                     [os.path.join(HERE, "host", "stabilization", "vsc_videostabilizer.cpp"),
                      os.path.join(HERE, "host", "stabilization", "vsc_flowconsistency.cpp"),
                      os.path.join(REF_STAB, "imagehelpers.cpp"),
                      os.path.join(ROOT, "tests", "cxx", "videostab_driver.cpp")],
                     ["-I", REF_STAB, "-I", os.path.dirname(REF_STAB), "-I", os.path.join(ROOT, "standins", "qt"),
                      "-I", CUDA_INC],
                     ["-lvsc_b200", "-L", CUDA_LIB, f"-Wl,-rpath,{CUDA_LIB}", "-lcudart"]))
    for out, srcs, inc, libs in jobs:
        deps = srcs + [LIB] + [os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(ROOT, "standins")) for f in fs] \
            + [os.path.join(HERE, "host", "inference", "vsc_flow_session.h")]
        if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
            built.append(out)
            continue
        r = subprocess.run([*common, *inc, "-o", out, *srcs, *libs], capture_output=True, text=True)
        if r.stderr.strip():
            print(r.stderr, file=sys.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"host shim build failed: {out}")
        built.append(out)
    return built


if __name__ == "__main__":
    print(build_host_shims(force="--force" in sys.argv))

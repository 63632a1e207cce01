"""Builds libvsc_b200.so (the C-ABI library, include/vsc/vsc.h) from csrc/*.cu for sm_100a.

In-tree, explicit nvcc: the .so lands in video-stream-consistency_b200/lib/ (git-ignored, but it
travels to the GPU box with the repo snapshot).  No JIT, no torch extension machinery: the library
has no torch types in its interface.

    python video-stream-consistency_b200/build.py [--force] [--ptxas-v]
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libvsc_b200.so")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
HOSTCXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

# -fmad=false: expressions are evaluated as written (see csrc/vsc_common.cuh); FMAs are explicit.
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
    "-ccbin", HOSTCXX, "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall", "-I", os.path.join(ROOT, "include"),
    "-I", CSRC,
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    hdr = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdr.append(os.path.join(ROOT, "include", "vsc", "vsc.h"))
    return sorted(hdr)


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, ptxas_v: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    flags = NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_v else [])
    srcs, deps = _sources(), _deps()
    stamp = os.path.join(OBJDIR, "stamp")
    want = _digest(srcs + deps, " ".join(flags))
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == want:
        return LIB

    def compile_one(src):
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        r = subprocess.run([NVCC, *flags, "-c", src, "-o", obj], capture_output=True, text=True)
        return src, obj, r

    objs = []
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        for src, obj, r in ex.map(compile_one, srcs):
            if r.stdout.strip():
                print(r.stdout, file=sys.stderr)
            if r.stderr.strip():
                print(r.stderr, file=sys.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}")
            objs.append(obj)
    r = subprocess.run([NVCC, "-shared", "-ccbin", HOSTCXX, "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB,
                        *objs], capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stdout, r.stderr, file=sys.stderr)
        raise RuntimeError("link failed")
    with open(stamp, "w") as f:
        f.write(want)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, ptxas_v="--ptxas-v" in sys.argv))


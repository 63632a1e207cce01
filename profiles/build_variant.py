"""Builds a variant of libvsc_b200.so with extra preprocessor defines, for A/B timing of kernel variants on the GPU box:

    python profiles/build_variant.py fo -DVSC_STREAM_FREE_ORDER=1     -> video-stream-consistency_b200/lib/libvsc_b200_fo.so
    VSC_B200_LIB=$PWD/video-stream-consistency_b200/lib/libvsc_b200_fo.so python profiles/sweep_bands.py

Only the sources that mention one of the defined macros are recompiled; the other objects come from the default
build (video-stream-consistency_b200/build/).  Not part of the product build.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "video-stream-consistency_b200")
sys.path.insert(0, PKG)
import build as B  # noqa: E402


def main():
    tag, defs = sys.argv[1], sys.argv[2:]
    B.build()
    macros = [re.sub(r"^-D([A-Za-z0-9_]+).*", r"\1", d) for d in defs]
    objdir = os.path.join(PKG, "build", "variant_" + tag)
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for src in B._sources():
        text = open(src).read()
        base = os.path.basename(src)[:-3] + ".o"
        if any(m in text for m in macros):
            obj = os.path.join(objdir, base)
            subprocess.run([B.NVCC, *B.NVCC_FLAGS, *defs, "-c", src, "-o", obj], check=True)
        else:
            obj = os.path.join(B.OBJDIR, base)
        objs.append(obj)
    out = os.path.join(B.LIBDIR, f"libvsc_b200_{tag}.so")
    subprocess.run([B.NVCC, "-shared", "-ccbin", B.HOSTCXX, "-gencode", "arch=compute_100a,code=sm_100a", "-o", out,
                    *objs], check=True)
    print(out)


if __name__ == "__main__":
    main()

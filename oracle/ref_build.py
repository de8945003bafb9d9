"""Build recipe for oracle/_ref/libasora_ref.so: the reference's CUDA ray tracer, UNMODIFIED, compiled
from the sources where they lie under /root/reference/src/asora for sm_100, plus oracle/ref_shim.cu.

Flags follow src/asora/Makefile:9-26 (-std=c++14 -O2 -D PERIODIC -D LOCALRATES, -dc per translation
unit, then a -shared device link) with the architecture changed from sm_60 to sm_100.  The CPython
wrapper (python_module.cu) is not needed: ctypes binds the extern "C" shim instead.

No reference source is copied into this repository; outputs go to oracle/_ref/ only (git-ignored,
shipped to the GPU box with the tree).  TEST INFRASTRUCTURE ONLY.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src/asora"
OUT = os.path.join(HERE, "_ref")
SO = os.path.join(OUT, "libasora_ref.so")
SO_GREY = os.path.join(OUT, "libasora_ref_grey.so")   # the same sources with -D GREY_NOTABLES (rates.cu:44-64)


def _build_one(so, extra, tag):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = ["-std=c++14", "-O2", "-Xcompiler", "-fPIC", "-D", "PERIODIC", "-D", "LOCALRATES",
             "-gencode", "arch=compute_100,code=sm_100", "-lineinfo", "-I", REF_SRC] + extra
    objs = []
    for src in ("memory.cu", "rates.cu", "raytracing.cu"):
        o = os.path.join(OUT, src.replace(".cu", tag + ".o"))
        subprocess.check_call([nvcc] + flags + ["-dc", os.path.join(REF_SRC, src), "-o", o])
        objs.append(o)
    o = os.path.join(OUT, "ref_shim" + tag + ".o")
    subprocess.check_call([nvcc] + flags + ["-dc", os.path.join(HERE, "ref_shim.cu"), "-o", o])
    objs.append(o)
    subprocess.check_call([nvcc] + flags + ["-shared", "-o", so] + objs)
    return so


def build(force=False):
    if not os.path.isdir(REF_SRC):
        raise RuntimeError(f"{REF_SRC} not present (GPU box?): using the prebuilt {SO} if it exists")
    os.makedirs(OUT, exist_ok=True)
    if force or not os.path.exists(SO):
        _build_one(SO, [], "")
    if force or not os.path.exists(SO_GREY):
        _build_one(SO_GREY, ["-D", "GREY_NOTABLES"], "_grey")
    return SO


if __name__ == "__main__":
    print(build(force=True))

"""In-tree nvcc build of libasora_b200.so for sm_100a (no JIT cache: the .so travels with the tree)."""
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
LIB = os.path.join(_PKG, "lib", "libasora_b200.so")
SOURCES = ["asora_api.cu", "sweep_plan.cu", "sweep_kernels.cu", "chemistry.cu"]
HEADERS = ["asora_common.cuh", os.path.join("..", "..", "include", "asora_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force=False, verbose=False):
    """Compile every CUDA source of the package into pyc2ray_b200/lib/libasora_b200.so."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(_PKG, "lib", "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if verbose:
        print(log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log)
    return LIB


if __name__ == "__main__":
    build_native(force=True, verbose=True)

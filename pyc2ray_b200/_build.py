"""In-tree nvcc build of libasora_b200.so for sm_100a (no JIT cache: the .so travels with the tree).

Every CUDA source is compiled to its own object file (in parallel, only when it or a header changed) and the
objects are linked into pyc2ray_b200/lib/libasora_b200.so."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
LIB = os.path.join(_PKG, "lib", "libasora_b200.so")
OBJ_DIR = os.path.join(_PKG, "lib", "build")
OCT_DEFS = ["-DASORA_OCT_PROBE"] if os.environ.get("ASORA_OCT_PROBE", "1") == "1" else []
# (source, extra defines, object name); sweep_octant.cu is compiled once per group of launch shapes
UNITS = [("asora_api.cu", [], "asora_api.o"), ("sweep_plan.cu", [], "sweep_plan.o"), ("sweep_kernels.cu", [], "sweep_kernels.o"),
         ("sweep_octant.cu", ["-DASORA_OCT_TU=0"] + OCT_DEFS, "sweep_octant_0.o"),
         ("sweep_octant.cu", ["-DASORA_OCT_TU=1"] + OCT_DEFS, "sweep_octant_1.o"),
         ("sweep_octant.cu", ["-DASORA_OCT_TU=2"] + OCT_DEFS, "sweep_octant_2.o"),
         ("sweep_octant.cu", ["-DASORA_OCT_TU=3"] + OCT_DEFS, "sweep_octant_3.o"),
         ("sweep_octant.cu", ["-DASORA_OCT_TU=4"] + OCT_DEFS, "sweep_octant_4.o"),
         ("sweep_octant.cu", ["-DASORA_OCT_TU=5"] + OCT_DEFS, "sweep_octant_5.o"),
         ("sweep_octant.cu", ["-DASORA_OCT_TU=6"] + OCT_DEFS, "sweep_octant_6.o"),
         ("sweep_octant.cu", ["-DASORA_OCT_TU=7"] + OCT_DEFS, "sweep_octant_7.o"),
         ("sweep_cluster.cu", [], "sweep_cluster.o"), ("chemistry.cu", [], "chemistry.o"), ("deterministic.cu", [], "deterministic.o")]
SOURCES = sorted({u[0] for u in UNITS})
HEADERS = ["asora_common.cuh", "sweep_device.cuh", os.path.join("..", "..", "include", "asora_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _units():
    return [u for u in UNITS if os.path.exists(os.path.join(CSRC, u[0]))]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    deps = [os.path.join(CSRC, s) for s in _sources() + HEADERS] + [os.path.abspath(__file__)]
    return _stale(LIB, deps)


def build_native(force=False, verbose=False):
    """Compile every CUDA source of the package into pyc2ray_b200/lib/libasora_b200.so."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    common = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    logs = {}

    def compile_one(unit):
        src, defines, objname = unit
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, objname)
        if not force and not _stale(obj, [path] + common):
            return obj, 0
        cmd = [nvcc] + NVCC_FLAGS + defines + ["-c", path, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        logs[objname] = " ".join(cmd) + "\n" + res.stdout + res.stderr
        return obj, res.returncode

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, _units()))
    log = "".join(logs[u[2]] for u in _units() if u[2] in logs)
    # one log file per source, so that an incremental build keeps the ptxas report of the files it did not recompile
    for s, text in logs.items():
        with open(os.path.join(OBJ_DIR, s + ".log"), "w") as f:
            f.write(text)
    if any(rc != 0 for _, rc in results):
        raise RuntimeError("nvcc failed:\n" + log)
    cmd = [nvcc, "-shared", "-Xcompiler", "-fPIC", "-o", LIB] + [o for o, _ in results]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log += " ".join(cmd) + "\n" + res.stdout + res.stderr
    with open(os.path.join(_PKG, "lib", "build.log"), "w") as f:
        f.write("".join(open(os.path.join(OBJ_DIR, u[2] + ".log")).read() for u in _units()
                        if os.path.exists(os.path.join(OBJ_DIR, u[2] + ".log"))) + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(log)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + log)
    return LIB


if __name__ == "__main__":
    import sys
    build_native(force="--incremental" not in sys.argv, verbose=True)

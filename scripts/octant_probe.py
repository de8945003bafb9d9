"""Launch-shape probe of the mirror-image sweep (variant 3) on the bench workload (256^3, uniform box, R = 30 and
R = 10.76): kernel-only milliseconds (CUDA events around the sweep kernel) for every instantiated shape and option,
next to the one-cell-per-thread sweep (variant 1).   usage: python scripts/octant_probe.py [nsrc] [R ...]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ASORA_QUIET"] = "1"
import pyc2ray_b200 as p
from pyc2ray_b200.lib import _cabi, libasora
from pyc2ray_b200.lib._cabi import L, check
from tests.fields import f0_fields, MPC, SIG
from tests.test_gpu_octant import SHAPES, BIG

N = 256
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
radii = [float(x) for x in sys.argv[2:]] or [30.0, 10.76]
ONLY = [tuple(int(v) for v in t.split(",")) for t in os.environ.get("ASORA_PROBE_SHAPES", "").split(";") if t]
KNOBS = [int(x) for x in os.environ.get("ASORA_PROBE_KNOBS", "0,1,4,5,8,16,2").split(",")]
thin, thick, dlogtau = p.blackbody_tables(1e5, False, -20.0, 4.0, 20000)
p.device_init(N, 64); p.photo_table_to_device(thin, thick)
nd, xh = f0_fields(N)
libasora.density_to_device(np.ascontiguousarray(nd.ravel()), N)
check(L.asora_buffer_upload(_cabi.BUF_XH_AV, _cabi.dptr(np.ascontiguousarray(xh.ravel()))))


def run(R, n, reps=3):
    best, bestk = 1e30, 1e30
    for r in range(reps):
        check(L.asora_raytrace_device(R, SIG, 3 * MPC / N, 0, n, -20.0, dlogtau, 20000, 1)); check(L.asora_sync())
        ms, kms = ctypes.c_float(0), ctypes.c_float(0)
        L.asora_last_sweep_stats(None, None, None, None, None, ctypes.byref(ms))
        check(L.asora_last_sweep_kernel_ms(ctypes.byref(kms)))
        if r > 0 or reps == 1:
            best, bestk = min(best, ms.value), min(bestk, kms.value)
    return best, bestk


for R in radii:
    n = ns if R > 20 else 10 * ns
    srcpos = p.generate_test_sources(N, n, seed=100)
    flux = 10 ** np.random.default_rng(100).normal(0, 0.5, size=n)
    pos_flat, flux_flat = p.format_sources(srcpos, flux)
    libasora.source_data_to_device(pos_flat, flux_flat, n)
    cells = int(L.asora_cells_per_source(N, R))
    for sphere in (0, 1):
        check(L.asora_set_sphere_only(sphere))
        check(L.asora_set_sweep_variant(1))
        ms, kms = run(R, n)
        print(f"R={R:g} n={n} sphere_only={sphere} variant 1 (automatic shape): sweep {ms:.3f} ms, kernel {kms:.3f} ms, "
              f"{n*cells/kms/1e6:.1f} G updates/s (full cell count)", flush=True)
        check(L.asora_set_sweep_variant(3))
        for shape in SHAPES:
            if ONLY and shape not in ONLY:
                continue
            if R > 20 and shape not in BIG:
                continue
            if R < 20 and (shape[3] > 512 or shape[0] != 8):
                continue
            # knobs: 1 log2 copies, 2 toggle the z-face copies (automatic: on for large sweeps), 4 plan entry prefetched,
            # 8 no de-duplication, 16 table gathers by LDG
            for knobs in (KNOBS if (R > 20 and sphere == 0) else ((0, 4) if shape not in BIG else (0,))):
                check(L.asora_set_octant_shape(*shape[:3], shape[3] | (knobs << 16)))
                try:
                    ms, kms = run(R, n)
                except RuntimeError as e:
                    print(f"R={R:g} shape={shape} knobs={knobs}: {e}", flush=True)
                    continue
                print(f"R={R:g} n={n} sphere_only={sphere} variant 3 shape={shape} knobs={knobs}: sweep {ms:.3f} ms, kernel {kms:.3f} ms, "
                      f"{n*cells/kms/1e6:.1f} G updates/s", flush=True)
        check(L.asora_set_tuning(0, 0)); check(L.asora_set_octant_shape(0, 0, 0, 0)); check(L.asora_set_sweep_variant(0))
    check(L.asora_set_sphere_only(0))
p.device_close()

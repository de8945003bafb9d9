"""Timing of the reference's own CUDA kernel (oracle/_ref/libasora_ref.so: src/asora compiled unmodified for sm_100)
on the bench workload, next to this repository's sweep -- the like-for-like bar of SURVEY section 8(d).  Test
infrastructure (lives under tests/, not collected by pytest); run on a GPU box:

    python tests/ref_kernel_timing.py [nsrc]
"""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ASORA_QUIET"] = "1"
import pyc2ray_b200 as p
from pyc2ray_b200.lib import _cabi, libasora
from tests.test_gpu_vs_reference_kernel import REF_SO

N = 256
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
dp = ctypes.POINTER(ctypes.c_double)
R_ = ctypes.CDLL(REF_SO)
R_.ref_device_init.argtypes = [ctypes.c_int, ctypes.c_int]
R_.ref_density_to_device.argtypes = [dp, ctypes.c_int]
R_.ref_photo_table_to_device.argtypes = [dp, dp, ctypes.c_int]
R_.ref_source_data_to_device.argtypes = [ctypes.POINTER(ctypes.c_int32), dp, ctypes.c_int]
R_.ref_do_all_sources.argtypes = [ctypes.c_double, dp, ctypes.c_double, ctypes.c_double, dp, dp, dp, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_double, ctypes.c_double, ctypes.c_int]
thin, thick, dlogtau = p.blackbody_tables(1e5, False, -20.0, 4.0, 20000)
srcpos = p.generate_test_sources(N, ns, seed=100)
flux = 10 ** np.random.default_rng(100).normal(0, 0.5, size=ns)
pos_flat, flux_flat = p.format_sources(srcpos, flux)
ndens = np.full(N ** 3, 1e-3); xh = np.full(N ** 3, 2e-4)
dr, sig = 3 * 3.086e24 / N, 6.3e-18
phi = np.zeros(N ** 3); dummy = np.zeros(1)
for R in (10.0, 30.0):
    cells = int(_cabi.L.asora_cells_per_source(N, R))
    sphere = 4.0 / 3.0 * np.pi * R ** 3
    for batch in (32, 64, 96, 128):
        assert R_.ref_device_init(N, batch) == 0
        R_.ref_zero_coldens(N, batch)
        R_.ref_density_to_device(ndens.ctypes.data_as(dp), N)
        R_.ref_photo_table_to_device(thin.ctypes.data_as(dp), thick.ctypes.data_as(dp), 20000)
        R_.ref_source_data_to_device(pos_flat.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), flux_flat.ctypes.data_as(dp), ns)
        best = 1e30
        for rep in range(3):
            t0 = time.perf_counter()
            rc = R_.ref_do_all_sources(R, dummy.ctypes.data_as(dp), sig, dr, ndens.ctypes.data_as(dp), xh.ctypes.data_as(dp),
                                       phi.ctypes.data_as(dp), ns, N, -20.0, dlogtau, 20000)
            best = min(best, time.perf_counter() - t0)
            assert rc == 0
        R_.ref_device_close()
        print(f"reference kernel  R={R:g} batch={batch}: {1e3*best:9.2f} ms for {ns} sources (wall, incl. 2x134 MB PCIe as in its "
              f"benchmark) = {ns*cells/best/1e9:7.3f} G updates/s; paper unit 3t/(Ns 4 pi R^3) = {best/(ns*sphere)*1e9:.3f} ns", flush=True)
    libasora.device_init(N, 64)
    libasora.photo_table_to_device(thin, thick, 20001)
    libasora.density_to_device(ndens, N)
    libasora.source_data_to_device(pos_flat, flux_flat, ns)
    best = 1e30
    for rep in range(4):
        t0 = time.perf_counter()
        libasora.do_all_sources(R, dummy, sig, dr, dummy, xh, phi, ns, N, -20.0, dlogtau, 20000)
        best = min(best, time.perf_counter() - t0)
    libasora.device_close()
    print(f"this repository   R={R:g}          : {1e3*best:9.2f} ms for {ns} sources (same call, pageable host buffers)         "
          f"= {ns*cells/best/1e9:7.3f} G updates/s; paper unit = {best/(ns*sphere)*1e9:.4f} ns", flush=True)

"""Sphere-only vs full sweep timing (250^3, 10^5 sources, R=10.76) and chemistry pass timing."""
import ctypes, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ASORA_QUIET"] = "1"
import pyc2ray_b200 as p
from pyc2ray_b200.lib import _cabi, libasora
from pyc2ray_b200.lib._cabi import L, check
N, ns = 250, 100000
rng = np.random.default_rng(244)
srcpos = p.generate_test_sources(N, ns, seed=244); flux = 10 ** rng.normal(5.0, 0.5, size=ns)
ndens = 1.87e-4 * np.exp(0.5 * rng.normal(size=N**3) - 0.125); xh = np.full(N**3, 2e-4)
thin, thick, dlogtau = p.blackbody_tables(1e5, False, -20.0, 4.0, 20000)
dr = 244.0 / 0.7 * 3.086e24 / N / 10.0; R = 15.0 * N * 0.7 / 244.0
p.device_init(N, 96); p.photo_table_to_device(thin, thick)
pos_flat, flux_flat = p.format_sources(srcpos, flux)
libasora.source_data_to_device(pos_flat, flux_flat, ns); libasora.density_to_device(ndens, N)
check(L.asora_buffer_upload(_cabi.BUF_XH_AV, _cabi.dptr(xh)))
for sph in (0, 1):
    check(L.asora_set_sphere_only(sph))
    for S, block in ((0, 0), (1, 256), (2, 256), (1, 512), (2, 512)):
        try:
            check(L.asora_set_tuning(S, block)); best = 1e30
            for r in range(4):
                check(L.asora_raytrace_device(R, 6.3e-18, dr, 0, ns, -20.0, dlogtau, 20000, 1)); check(L.asora_sync())
                ms = ctypes.c_float(0); upd = ctypes.c_int64(0)
                L.asora_last_sweep_stats(None, None, ctypes.byref(upd), None, None, ctypes.byref(ms))
                if r > 0: best = min(best, ms.value)
            print(f"sphere_only={sph} S={S} block={block}: {best:.3f} ms, {upd.value/best/1e6:.2f} G updates/s", flush=True)
        except RuntimeError as e: print("skip", sph, S, block, e)
check(L.asora_set_sphere_only(0)); check(L.asora_set_tuning(0, 0))
for b in (_cabi.BUF_XH, _cabi.BUF_XH_INTERMED): check(L.asora_buffer_upload(b, _cabi.dptr(xh)))
check(L.asora_buffer_upload(_cabi.BUF_TEMP, _cabi.dptr(np.full(N**3, 1e4))))
f = ctypes.c_int(0); a = ctypes.c_double(0); b2 = ctypes.c_double(0)
for r in range(3):
    t0 = time.perf_counter()
    check(L.asora_global_pass_device(3.15576e14, 2.59e-13, -0.7, 5.8e-11, 157800.0, 7.1e-7, ctypes.byref(f), ctypes.byref(a), ctypes.byref(b2)))
    print(f"global_pass_device wall {1e3*(time.perf_counter()-t0):.3f} ms, conv_flag {f.value}")
p.device_close()

"""Where the time of one synthetic 250^3 / 10^5-source evolve3D step goes (bench.py: eor_step): whole call, convergence loop,
one sphere-only sweep, one chemistry pass.   usage: python scripts/eor_breakdown.py [reps]"""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ASORA_QUIET"] = "1"
import bench
import pyc2ray_b200 as p
from pyc2ray_b200.lib import _cabi
L, check = _cabi.L, _cabi.check
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
thin, thick, dlogtau, numtau = bench.tables()
N, nsrc = 250, 100000
srcpos, flux, ndens, xh, temp, dr, R = bench.eor_inputs(N, nsrc)
dt = 1e7 * 3.15576e7
p.device_init(N, 96)
p.photo_table_to_device(thin, thick)
for rep in range(reps):
    t0 = time.perf_counter()
    x, phi = p.evolve3D(dt, dr, flux, srcpos, True, 1000, 64, 1e-2, temp, ndens, xh, thin, thick, -20.0, dlogtau, R, 1e-4,
                        bench.SIG, *bench.CHEM, logfile=None, quiet=True)
    wall = time.perf_counter() - t0
    print(f"evolve3D: {1e3*wall:.1f} ms, loop {1e3*p.evolve3D.last_loop_seconds:.1f} ms, {p.evolve3D.last_niter} iterations", flush=True)
check(L.asora_set_sphere_only(1))
for rep in range(3):
    t0 = time.perf_counter()
    check(L.asora_raytrace_device(R, bench.SIG, dr, 0, nsrc, -20.0, dlogtau, thin.size, 1)); check(L.asora_sync())
    wall = time.perf_counter() - t0
    ms, kms, v = ctypes.c_float(0), ctypes.c_float(0), ctypes.c_int(0)
    L.asora_last_sweep_stats(ctypes.byref(v), None, None, None, None, ctypes.byref(ms)); L.asora_last_sweep_kernel_ms(ctypes.byref(kms))
    print(f"sphere-only sweep: wall {1e3*wall:.2f} ms, device {ms.value:.2f} ms, kernel {kms.value:.2f} ms, variant {v.value}", flush=True)
check(L.asora_set_sphere_only(0))
flag, s1, s0 = ctypes.c_int(0), ctypes.c_double(0), ctypes.c_double(0)
for rep in range(3):
    t0 = time.perf_counter()
    check(L.asora_global_pass_device(dt, *bench.CHEM, ctypes.byref(flag), ctypes.byref(s1), ctypes.byref(s0)))
    print(f"chemistry pass: {1e3*(time.perf_counter()-t0):.2f} ms", flush=True)
p.device_close()

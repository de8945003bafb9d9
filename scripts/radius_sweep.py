"""Throughput against ray-tracing radius and source count (the axes of the reference's raytracing_benchmark:
test/paper_tests/raytracing_benchmark/run_test.py:24,48,82-93), mesh 256 (BASELINE) or 250 (the paper's own; ASORA_SWEEP_MESH),
device-resident inputs, automatic launch shape.  Prints a markdown table.  Also reachable as `python bench.py --sweep`."""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ASORA_QUIET"] = "1"
import pyc2ray_b200 as p
from pyc2ray_b200.lib import _cabi, libasora
from pyc2ray_b200.lib._cabi import L, check
from tests.fields import f0_fields, MPC, SIG
thin, thick, dlogtau = p.blackbody_tables(1e5, False, -20.0, 4.0, 20000)
N = int(os.environ.get("ASORA_SWEEP_MESH", "256"))
p.device_init(N, 64); p.photo_table_to_device(thin, thick)
nd, xh = f0_fields(N)
libasora.density_to_device(np.ascontiguousarray(nd.ravel()), N)
check(L.asora_buffer_upload(_cabi.BUF_XH_AV, _cabi.dptr(np.ascontiguousarray(xh.ravel()))))
print("| R (cells) | sources | variant | levels | ms | us/source | G updates/s | ns per source and sphere cell (paper unit) |")
print("|---|---|---|---|---|---|---|---|")
for R, counts in ((5.0, (10000, 100000)), (10.0, (1, 100, 10000, 100000, 1000000)), (10.76, (100000,)), (20.0, (10000,)),
                  (30.0, (1, 100, 1000, 10000, 100000)), (36.0, (4000,)), (40.0, (2000,)), (45.0, (2000,)), (48.0, (1000,)), (50.0, (1000,)), (60.0, (500,)), (70.0, (300,)), (80.0, (128,)), (100.0, (64,)),
                  (1e4, (1, 16, 64))):
    for ns in counts:
        srcpos = p.generate_test_sources(N, ns); flux = np.ones(ns)
        pos_flat, flux_flat = p.format_sources(srcpos, flux)
        libasora.source_data_to_device(pos_flat, flux_flat, ns)
        best = 1e30
        for r in range(3):
            check(L.asora_raytrace_device(R, SIG, 3 * MPC / N, 0, ns, -20.0, dlogtau, 20000, 1)); check(L.asora_sync())
            ms = ctypes.c_float(0); v = ctypes.c_int(0); upd = ctypes.c_int64(0); lv = ctypes.c_int(0)
            L.asora_last_sweep_stats(ctypes.byref(v), None, ctypes.byref(upd), None, ctypes.byref(lv), ctypes.byref(ms))
            if r > 0: best = min(best, ms.value)
        Reff = min(R, N * 0.5 * 3 ** 0.5)
        paper = best * 1e-3 / (ns * 4.0 / 3.0 * np.pi * Reff ** 3) * 1e9
        print(f"| {R:g} | {ns} | {v.value} | {lv.value} | {best:.3f} | {best/ns*1e3:.2f} | {upd.value/best/1e6:.1f} | {paper:.4f} |", flush=True)
p.device_close()

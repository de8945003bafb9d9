"""Timing probe of the grid-cooperative (large radius) sweep."""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ASORA_QUIET"] = "1"
import pyc2ray_b200 as p
from pyc2ray_b200.lib import _cabi, libasora
from pyc2ray_b200.lib._cabi import L, check
from tests.fields import f0_fields, f1_fields, MPC, SIG

def run(N, R, ns, field="f0", variant=0, reps=2):
    srcpos = p.generate_test_sources(N, ns)
    flux = 10 ** np.random.default_rng(7).normal(0, 0.5, size=ns)
    nd, xh = f0_fields(N) if field == "f0" else f1_fields(N, srcpos)
    dr = 3 * MPC / N
    pos_flat, flux_flat = p.format_sources(srcpos, flux)
    libasora.source_data_to_device(pos_flat, flux_flat, ns)
    libasora.density_to_device(np.ascontiguousarray(nd.ravel()), N)
    check(L.asora_buffer_upload(_cabi.BUF_XH_AV, _cabi.dptr(np.ascontiguousarray(xh.ravel()))))
    check(L.asora_set_sweep_variant(variant))
    best = 1e30
    for r in range(reps + 1):
        check(L.asora_raytrace_device(R, SIG, dr, 0, ns, -20.0, dlogtau, NumTau, 1)); check(L.asora_sync())
        ms = ctypes.c_float(0); v = ctypes.c_int(0); upd = ctypes.c_int64(0); lv = ctypes.c_int(0)
        L.asora_last_sweep_stats(ctypes.byref(v), None, ctypes.byref(upd), None, ctypes.byref(lv), ctypes.byref(ms))
        if r > 0: best = min(best, ms.value)
    print(f"N={N} R={R} ns={ns} {field} variant={v.value} levels={lv.value}: {best:.3f} ms, {best/ns*1e3:.1f} us/source, "
          f"{upd.value/best/1e6:.2f} G updates/s", flush=True)

thin, thick, dlogtau = p.blackbody_tables(1e5, False, -20.0, 4.0, 20000)
NumTau = 20000
N = 256
p.device_init(N, 64); p.photo_table_to_device(thin, thick)
run(N, 1e4, 1, "f0"); run(N, 1e4, 4, "f0"); run(N, 1e4, 16, "f1")
run(N, 100.0, 32, "f1"); run(N, 50.0, 128, "f1", variant=2); run(N, 50.0, 1000, "f0", variant=2); run(N, 50.0, 1000, "f0"); run(N, 60.0, 500, "f0"); run(N, 40.0, 1000, "f0")
run(N, 30.0, 200, "f0", variant=2); run(N, 30.0, 200, "f0", variant=1)
run(N, 10.0, 200, "f0", variant=2)
p.device_close()
N = 128
p.device_init(N, 64); p.photo_table_to_device(thin, thick)
run(N, 1e4, 1, "f0"); run(N, 1e4, 5, "f0"); run(N, 1e4, 64, "f1")
p.device_close()

"""Exploratory timings of the sweep variants on one B200 (not the bench; numbers feed DESIGN.md)."""
import ctypes, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyc2ray_b200 as p
from pyc2ray_b200.lib import _cabi, libasora
from pyc2ray_b200.lib._cabi import L, check
from tests.fields import f0_fields, f1_fields, MPC, SIG

def run(N, R, ns, field="f0", S=0, block=0, variant=0, reps=3, dr=None):
    srcpos = p.generate_test_sources(N, ns)
    rng = np.random.default_rng(7)
    flux = 10 ** rng.normal(0, 0.5, size=ns)
    nd, xh = f0_fields(N) if field == "f0" else f1_fields(N, srcpos)
    dr = dr or 3 * MPC / N
    pos_flat, flux_flat = p.format_sources(srcpos, flux)
    libasora.source_data_to_device(pos_flat, flux_flat, ns)
    libasora.density_to_device(np.ascontiguousarray(nd.ravel()), N)
    check(L.asora_buffer_upload(_cabi.BUF_XH_AV, _cabi.dptr(np.ascontiguousarray(xh.ravel()))))
    check(L.asora_set_sweep_variant(variant)); check(L.asora_set_tuning(S, block))
    best = 1e30
    for r in range(reps + 1):
        check(L.asora_raytrace_device(R, SIG, dr, 0, ns, -20.0, dlogtau, NumTau, 1))
        check(L.asora_sync())
        ms = ctypes.c_float(0); v = ctypes.c_int(0); upd = ctypes.c_int64(0); lv = ctypes.c_int(0)
        L.asora_last_sweep_stats(ctypes.byref(v), None, ctypes.byref(upd), None, ctypes.byref(lv), ctypes.byref(ms))
        if r > 0: best = min(best, ms.value)
    print(f"N={N} R={R} ns={ns} {field} S={S} block={block} variant={v.value} levels={lv.value}: {best:.3f} ms, "
          f"{upd.value/best/1e6:.3f} G updates/s, {upd.value/best/1e6*32/6530.3*100:.2f}% of HBM roofline @32B", flush=True)

thin, thick, dlogtau = p.blackbody_tables(1e5, False, -20.0, 4.0, 20000)
NumTau = 20000
N = 256
p.device_init(N, 64)
p.photo_table_to_device(thin, thick)
for R in (10.0, 10.76, 30.0):
    for S, block in ((1, 256), (1, 512), (2, 256), (2, 512), (1, 768), (1, 896), (1, 1024), (2, 1024)):
        try:
            run(N, R, 10000, "f0", S, block)
        except RuntimeError as e:
            print("skip", R, S, block, e)
run(N, 10.76, 10000, "f1", 0, 0)
run(N, 30.0, 10000, "f1", 0, 0)
run(N, 10.0, 100, "f0", 0, 0, variant=2)
run(N, 1e4, 2, "f0", 0, 0, variant=2)
run(N, 100.0, 4, "f1", 0, 0, variant=2)
p.device_close()

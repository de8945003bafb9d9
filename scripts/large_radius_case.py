"""One large-radius workload for profiling: 256^3 uniform box, `nsrc` sources, radius R (1e4 = full box), forced
sweep variant.   usage: python scripts/large_radius_case.py [variant] [nsrc] [R] [reps]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ASORA_QUIET"] = "1"
import pyc2ray_b200 as p
from pyc2ray_b200.lib import _cabi, libasora
from pyc2ray_b200.lib._cabi import L, check
from tests.fields import f0_fields, MPC, SIG

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 16
R = float(sys.argv[3]) if len(sys.argv) > 3 else 1e4
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
N = int(os.environ.get("ASORA_CASE_N", "256"))
thin, thick, dlogtau = p.blackbody_tables(1e5, False, -20.0, 4.0, 20000)
p.device_init(N, 64); p.photo_table_to_device(thin, thick)
nd, xh = f0_fields(N)
libasora.density_to_device(np.ascontiguousarray(nd.ravel()), N)
check(L.asora_buffer_upload(_cabi.BUF_XH_AV, _cabi.dptr(np.ascontiguousarray(xh.ravel()))))
srcpos = p.generate_test_sources(N, ns, seed=100)
pos_flat, flux_flat = p.format_sources(srcpos, np.ones(ns))
libasora.source_data_to_device(pos_flat, flux_flat, ns)
check(L.asora_set_sweep_variant(variant))
if os.environ.get("ASORA_OCT_SHAPE"):  # "noct,opt,batch,block[,knobs]"
    f = [int(x) for x in os.environ["ASORA_OCT_SHAPE"].split(",")]
    check(L.asora_set_octant_shape(f[0], f[1], f[2], f[3] | ((f[4] if len(f) > 4 else 0) << 16)))
if os.environ.get("ASORA_CLUSTER_SHAPE"):  # "ctas,threads"
    check(L.asora_set_cluster_shape(*[int(x) for x in os.environ["ASORA_CLUSTER_SHAPE"].split(",")]))
for r in range(reps):
    check(L.asora_raytrace_device(R, SIG, 3 * MPC / N, 0, ns, -20.0, dlogtau, 20000, 1)); check(L.asora_sync())
    ms, kms = ctypes.c_float(0), ctypes.c_float(0)
    v, upd, lv = ctypes.c_int(0), ctypes.c_int64(0), ctypes.c_int(0)
    L.asora_last_sweep_stats(ctypes.byref(v), None, ctypes.byref(upd), None, ctypes.byref(lv), ctypes.byref(ms))
    check(L.asora_last_sweep_kernel_ms(ctypes.byref(kms)))
    print(f"N={N} R={R:g} sources={ns} variant={v.value} levels={lv.value}: sweep {ms.value:.3f} ms, kernel {kms.value:.3f} ms, "
          f"{kms.value/ns*1e3:.1f} us/source, {upd.value/kms.value/1e6:.2f} G updates/s", flush=True)
p.device_close()

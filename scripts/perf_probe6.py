"""Launch-option sweep of the shared-memory sweep kernel (256^3, 10^4 sources): log2 copies / texture gathers / offsets
word one cell ahead (bits 16-18 of asora_set_tuning's block_threads toggle the defaults).  Also checks phi against the
default options."""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ASORA_QUIET"] = "1"
import pyc2ray_b200 as p
from pyc2ray_b200.lib import _cabi, libasora
from pyc2ray_b200.lib._cabi import L, check
from tests.fields import f0_fields, MPC, SIG
thin, thick, dlogtau = p.blackbody_tables(1e5, False, -20.0, 4.0, 20000)
N, ns = 256, 10000
p.device_init(N, 64); p.photo_table_to_device(thin, thick)
srcpos = p.generate_test_sources(N, ns); flux = 10 ** np.random.default_rng(7).normal(0, 0.5, size=ns)
nd, xh = f0_fields(N); pos_flat, flux_flat = p.format_sources(srcpos, flux)
libasora.source_data_to_device(pos_flat, flux_flat, ns); libasora.density_to_device(np.ascontiguousarray(nd.ravel()), N)
check(L.asora_buffer_upload(_cabi.BUF_XH_AV, _cabi.dptr(np.ascontiguousarray(xh.ravel()))))
toggles = [int(a) for a in sys.argv[1:]] or list(range(8))
for R, S, block in ((30.0, 0, 0), (10.76, 0, 0)):  # automatic launch shape, options toggled
    ref = None
    for t in toggles:
        check(L.asora_set_tuning(S, block | (t << 16))); best = 1e30
        for r in range(4):
            check(L.asora_raytrace_device(R, SIG, 3 * MPC / N, 0, ns, -20.0, dlogtau, 20000, 1)); check(L.asora_sync())
            ms = ctypes.c_float(0); upd = ctypes.c_int64(0)
            L.asora_last_sweep_stats(None, None, ctypes.byref(upd), None, None, ctypes.byref(ms))
            if r > 0: best = min(best, ms.value)
        phi = np.empty(N ** 3); check(L.asora_buffer_download(_cabi.BUF_PHI_ION, _cabi.dptr(phi)))
        if ref is None: ref = phi
        err = np.max(np.abs(phi - ref) / np.maximum(np.abs(ref), 1e-300))
        print(f"R={R} S={S} block={block} toggle={t}: {best:.3f} ms, {upd.value/best/1e6:.2f} G updates/s, max rel diff vs first {err:.2e}", flush=True)
p.device_close()

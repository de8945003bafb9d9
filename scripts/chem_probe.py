"""Device-resident chemistry passes at 256^3 for ncu (first pass fills the temperature-factor cache)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ASORA_QUIET"] = "1"
import pyc2ray_b200 as p
from pyc2ray_b200.lib import _cabi
from pyc2ray_b200.lib._cabi import L, check, dptr
N = 256
rng = np.random.default_rng(0)
p.device_init(N, 1)
x = rng.uniform(1e-4, 0.5, size=N ** 3)
for b, a in ((_cabi.BUF_NDENS, 1e-3 * np.exp(rng.normal(size=N ** 3) * 0.3)), (_cabi.BUF_TEMP, np.full(N ** 3, 1e4)),
             (_cabi.BUF_XH, x), (_cabi.BUF_XH_AV, x), (_cabi.BUF_XH_INTERMED, x),
             (_cabi.BUF_PHI_ION, 10 ** rng.uniform(-16, -11, size=N ** 3))):
    check(L.asora_buffer_upload(b, dptr(np.ascontiguousarray(a))))
f = ctypes.c_int(0); s1 = ctypes.c_double(0); s0 = ctypes.c_double(0)
for it in range(4):
    check(L.asora_global_pass_device(3.15576e13, 2.59e-13, -0.7, 5.8e-11, 157800.0, 7.1e-7, ctypes.byref(f), ctypes.byref(s1), ctypes.byref(s0)))
    print("pass", it, "conv_flag", f.value, "sum x", s1.value)
p.device_close()

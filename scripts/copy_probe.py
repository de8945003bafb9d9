"""Host<->device copy rates of the C ABI with pageable (numpy) and pinned (torch) host buffers, 256^3 grids."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ASORA_QUIET"] = "1"
import torch
import pyc2ray_b200 as p
from pyc2ray_b200.lib import _cabi
from pyc2ray_b200.lib._cabi import L, check
N = 256
p.device_init(N, 8)
a = np.random.default_rng(0).uniform(size=N ** 3)
b = np.empty_like(a)
pin = torch.empty(N ** 3, dtype=torch.float64).pin_memory()
pin.numpy()[:] = a
for name, src, dst in (("pageable", a, b), ("pinned", pin.numpy(), pin.numpy())):
    for r in range(3):
        t0 = time.perf_counter(); check(L.asora_buffer_upload(_cabi.BUF_XH, _cabi.dptr(src))); t1 = time.perf_counter()
        check(L.asora_buffer_download(_cabi.BUF_XH, _cabi.dptr(dst))); t2 = time.perf_counter()
    gb = a.nbytes / 1e9
    print(f"{name}: upload {1e3*(t1-t0):.2f} ms ({gb/(t1-t0):.1f} GB/s), download {1e3*(t2-t1):.2f} ms ({gb/(t2-t1):.1f} GB/s)")
assert np.array_equal(a, b)
af = np.asfortranarray(a.reshape(N, N, N)); bf = np.empty((N, N, N), order="F")
check(L.asora_buffer_upload_f(_cabi.BUF_XH, _cabi.dptr(af))); check(L.asora_buffer_download_f(_cabi.BUF_XH, _cabi.dptr(bf)))
assert np.array_equal(af, bf)
c = np.empty_like(a); check(L.asora_buffer_download(_cabi.BUF_XH, _cabi.dptr(c)))
assert np.array_equal(c.reshape(N, N, N), af)  # device layout is the logical C order
print("round trips exact")
p.device_close()

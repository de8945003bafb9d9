"""Small sweeps + chemistry for compute-sanitizer (memcheck / racecheck): every sweep variant, split and
sphere-only modes, odd and even meshes."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ASORA_QUIET"] = "1"
from pyc2ray_b200.lib import _cabi, libasora, libc2ray
from tests.fields import make_case
L, check = _cabi.L, _cabi.check
for name in ("small_r5", "odd_n15_full", "multi_n32"):
    c = make_case(name)
    libasora.device_init(c["N"], 8)
    libasora.photo_table_to_device(c["thin"], c["thick"], c["NumTau"])
    libasora.density_to_device(np.ascontiguousarray(c["ndens"].ravel()), c["N"])
    libasora.source_data_to_device(c["pos_flat"], c["flux_flat"], c["flux_flat"].size)
    xh = np.ascontiguousarray(c["xh"].ravel()); ref = None
    for variant, parts, sph in ((1, 0, 0), (2, 0, 0), (1, 2, 0), (1, 8, 1), (2, 0, 1), (1, 4, 0)):
        check(L.asora_set_sweep_variant(variant)); check(L.asora_set_tuning(0, parts << 20)); check(L.asora_set_sphere_only(sph))
        phi = np.zeros(c["N"] ** 3)
        libasora.do_all_sources(c["R"], np.zeros(1), c["sig"], c["dr"], np.zeros(1), xh, phi, c["flux_flat"].size, c["N"],
                                c["minlogtau"], c["dlogtau"], c["NumTau"])
        if ref is None: ref = phi
        assert np.allclose(phi, ref, rtol=1e-10, atol=1e-14 * ref.max()), (name, variant, parts, sph)
    check(L.asora_set_sweep_variant(0)); check(L.asora_set_tuning(0, 0)); check(L.asora_set_sphere_only(0))
    libasora.device_close()
rng = np.random.default_rng(0); shape = (9, 10, 11)
a = lambda lo, hi: np.asfortranarray(rng.uniform(lo, hi, size=shape))
x = a(1e-4, 0.9); xa, xi = x.copy(order="F"), x.copy(order="F")
print("conv_flag", libc2ray.chemistry.global_pass(3e13, a(1e-4, 1e-2), a(5e3, 2e4), x, xa, xi, a(0, 1e-12), 2.59e-13, -0.7, 5.8e-11, 157800.0, 7.1e-7))
print("sanitize_small ok")

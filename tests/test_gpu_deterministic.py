"""Deterministic accumulation mode (asora_set_deterministic: 128-bit fixed-point sums by integer reductions): phi_ion is
bit-identical from run to run and across splits / launch shapes, and agrees with the default fp64 reductions and with
the oracle to rounding."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dense_case():
    """Many overlapping sources: 300 sources at R = 7.5 in a 32^3 box (every cell receives ~100 contributions)."""
    from tests.fields import make_case
    from pyc2ray_b200.utils.sourceutils import format_sources, generate_test_sources
    c = make_case("multi_n32")
    ns = 300
    srcpos = generate_test_sources(c["N"], ns, seed=77)
    flux = 10 ** np.random.default_rng(77).normal(0, 1.0, size=ns)   # three decades of fluxes
    c["pos_flat"], c["flux_flat"] = format_sources(srcpos, flux)
    return c


def test_deterministic_mode_is_bit_reproducible_and_matches_default():
    import oracle
    from pyc2ray_b200.lib import _cabi, libasora
    from tests.test_gpu_parity import _setup, _sweep, _assert_close
    c = _dense_case()
    ref, _, _ = oracle.asora_do_all_sources(c["R"], c["sig"], c["dr"], c["ndens"].ravel(), c["xh"].ravel(), c["pos_flat"],
                                            c["flux_flat"], c["N"], c["thin"], c["thick"], c["minlogtau"], c["dlogtau"], c["NumTau"])
    _setup(libasora, c)
    try:
        default, _, _ = _sweep(libasora, _cabi, c, 1)
        _cabi.check(_cabi.L.asora_set_deterministic(1))
        first, used, _ = _sweep(libasora, _cabi, c, 1)
        assert used == 1
        for rep in range(5):
            again, _, _ = _sweep(libasora, _cabi, c, 1)
            assert np.array_equal(first, again), "deterministic mode: two runs differ"
        for parts in (2, 4, 8):  # half-spaces, quadrants, octants as separate CTAs
            _cabi.check(_cabi.L.asora_set_tuning(0, parts << 20))
            split, _, _ = _sweep(libasora, _cabi, c, 1)
            assert np.array_equal(first, split), f"deterministic mode: parts = {parts} differs from the unsplit sweep"
        _cabi.check(_cabi.L.asora_set_tuning(0, 0))
        for shape in ((8, 4, 2, 512), (4, 4, 4, 256), (2, 2, 2, 192), (8, 2, 1, 128)):  # mirror-image sweep
            _cabi.check(_cabi.L.asora_set_octant_shape(*shape))
            oct_, used, _ = _sweep(libasora, _cabi, c, 3)
            assert used == 3
            assert np.array_equal(first, oct_), f"deterministic mode: mirror-image sweep {shape} differs"
        _cabi.check(_cabi.L.asora_set_octant_shape(0, 0, 0, 0))
        _cabi.check(_cabi.L.asora_set_sphere_only(1))
        sph, _, _ = _sweep(libasora, _cabi, c, 1)
        _cabi.check(_cabi.L.asora_set_sphere_only(0))
        assert np.array_equal(first, sph), "deterministic mode: sphere-only differs"
        _assert_close(first, default, "deterministic vs default accumulation", rtol=1e-13, floor=1e-15)
        _assert_close(first, ref, "deterministic mode vs oracle")
        # accumulating on top of earlier rates
        ns = c["flux_flat"].size
        h = ns // 3
        _cabi.check(_cabi.L.asora_set_sweep_variant(1))
        _cabi.check(_cabi.L.asora_buffer_upload(_cabi.BUF_XH_AV, _cabi.dptr(np.ascontiguousarray(c["xh"].ravel()))))
        args = (c["R"], c["sig"], c["dr"])
        tail = (c["minlogtau"], c["dlogtau"], c["NumTau"])
        _cabi.check(_cabi.L.asora_raytrace_device(*args, 0, h, *tail, 1))
        _cabi.check(_cabi.L.asora_raytrace_device(*args, h, ns - h, *tail, 0))
        acc = np.empty(c["N"] ** 3)
        _cabi.check(_cabi.L.asora_buffer_download(_cabi.BUF_PHI_ION, _cabi.dptr(acc)))
        _assert_close(acc, ref, "deterministic mode, two accumulating sweeps")
    finally:
        _cabi.L.asora_set_sweep_variant(0)
        _cabi.L.asora_set_deterministic(0)
        _cabi.L.asora_set_tuning(0, 0)
        _cabi.L.asora_set_octant_shape(0, 0, 0, 0)
        _cabi.L.asora_set_sphere_only(0)
        libasora.device_close()


def test_deterministic_mode_with_heating():
    from pyc2ray_b200.lib import _cabi, libasora
    from tests.test_gpu_parity import _setup, _assert_close
    from tests.test_heating import heat_case
    c = heat_case("multi_n32")
    _setup(libasora, c)
    try:
        libasora.heat_table_to_device(c["heat_thin"], c["heat_thick"], c["NumTau"])
        n3 = c["N"] ** 3
        xh = np.ascontiguousarray(c["xh"].ravel())

        def run():
            phi, heat = np.zeros(n3), np.zeros(n3)
            libasora.do_all_sources_heat(c["R"], np.zeros(1), c["sig"], c["dr"], np.zeros(1), xh, phi, heat, c["flux_flat"].size,
                                         c["N"], c["minlogtau"], c["dlogtau"], c["NumTau"])
            return phi, heat
        phi0, heat0 = run()
        _cabi.check(_cabi.L.asora_set_deterministic(1))
        phi1, heat1 = run()
        phi2, heat2 = run()
        assert np.array_equal(phi1, phi2) and np.array_equal(heat1, heat2)
        _assert_close(phi1, phi0, "deterministic phi_ion with heating", rtol=1e-13, floor=1e-15)
        _assert_close(heat1, heat0, "deterministic phi_heat", rtol=1e-13, floor=1e-15)
    finally:
        _cabi.L.asora_set_deterministic(0)
        libasora.device_close()

"""GPU parity of the large-radius sweep (variant 4, csrc/sweep_cluster.cu: one thread-block cluster per wedge, level buffers in
distributed shared memory): against the oracle on every seeded case and cluster shape, against the reference's own CUDA kernel
per cell on a full 128^3 box, the debug column densities, sphere-only, heating, deterministic accumulation."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [(8, 256), (8, 384), (8, 512), (4, 256), (4, 512), (2, 256), (2, 512), (1, 256), (1, 512)]
CASE_NAMES = ["small_r5", "clip_full_n24", "odd_n15_full", "r_int5", "multi_n32", "bench_like_n32", "thin_n24", "mid_n48_r14"]
CASE_RTOL = {"thin_n24": 1e-6}   # see tests/test_gpu_parity.py


def _oracle(c):
    import oracle
    return oracle.asora_do_all_sources(c["R"], c["sig"], c["dr"], c["ndens"].ravel(), c["xh"].ravel(), c["pos_flat"],
                                       c["flux_flat"], c["N"], c["thin"], c["thick"], c["minlogtau"], c["dlogtau"], c["NumTau"])


@pytest.mark.parametrize("name", CASE_NAMES)
def test_cluster_sweep_vs_oracle(name):
    from pyc2ray_b200.lib import _cabi, libasora
    from tests.fields import make_case
    from tests.test_gpu_parity import _setup, _sweep, _assert_close
    c = make_case(name)
    ref, _, n = _oracle(c)
    _setup(libasora, c)
    try:
        for shape in SHAPES:
            _cabi.check(_cabi.L.asora_set_cluster_shape(*shape))
            for sphere_only in (0, 1):
                _cabi.check(_cabi.L.asora_set_sphere_only(sphere_only))
                phi, used, upd = _sweep(libasora, _cabi, c, 4)
                assert used == 4
                if not sphere_only:
                    assert upd == n
                assert np.isfinite(phi).all()
                _assert_close(phi, ref, f"{name} cluster {shape} sphere_only={sphere_only}", rtol=CASE_RTOL.get(name, 1e-9))
    finally:
        _cabi.L.asora_set_cluster_shape(0, 0)
        _cabi.L.asora_set_sphere_only(0)
        libasora.device_close()


@pytest.mark.parametrize("name", ["small_r5", "clip_full_n24", "odd_n15_full"])
def test_cluster_column_density_vs_oracle(name):
    from pyc2ray_b200.lib import _cabi, libasora
    from tests.fields import make_case
    from tests.test_gpu_parity import _setup, _assert_close
    c = make_case(name)
    ref_phi, ref_cdh, _ = _oracle(c)
    _setup(libasora, c)
    try:
        _cabi.check(_cabi.L.asora_set_sweep_variant(4))
        cdh, phi = np.zeros(c["N"] ** 3), np.zeros(c["N"] ** 3)
        xh = np.ascontiguousarray(c["xh"].ravel())
        _cabi.check(_cabi.L.asora_debug_single_source(c["R"], c["sig"], c["dr"], _cabi.dptr(xh), 0, c["minlogtau"], c["dlogtau"],
                                                      c["NumTau"], _cabi.dptr(cdh), _cabi.dptr(phi)))
    finally:
        _cabi.L.asora_set_sweep_variant(0)
        libasora.device_close()
    assert ((cdh != 0) == (ref_cdh != 0)).all(), "visited-cell sets differ"
    np.testing.assert_allclose(cdh, ref_cdh, rtol=1e-12, atol=0)
    _assert_close(phi, ref_phi, f"{name} cluster sweep, debug path")


def test_cluster_full_box_128_vs_reference_kernel():
    """Full 128^3 box (q_max = 193 clipped to 65 levels), 6 sources on non-trivial fields, per cell against the reference's
    own kernel (oracle/_ref) and the oracle; also the eight-octant shared-memory sweep, which the automatic selection
    picks for four or more sources while an octant's levels fit one SM."""
    from pyc2ray_b200.lib import _cabi, libasora
    from tests.fields import f1_fields, tables, SIG
    from tests.test_gpu_parity import _setup, _sweep, _assert_close
    from tests.test_gpu_vs_reference_kernel import REF_SO, run_reference
    from pyc2ray_b200.utils.sourceutils import format_sources, generate_test_sources
    import os
    N, ns, R = 128, 6, 1e4
    srcpos = generate_test_sources(N, ns, seed=100)
    flux = 10 ** np.random.default_rng(12).normal(0, 0.5, size=ns)
    ndens, xh = f1_fields(N, srcpos, mean_dens=2e-4)
    thin, thick, dlogtau, numtau = tables("bb1e5")
    pos_flat, flux_flat = format_sources(srcpos, flux)
    c = dict(N=N, R=R, sig=SIG, dr=4e20, ndens=ndens, xh=xh, thin=thin, thick=thick, minlogtau=-20.0, dlogtau=dlogtau,
             NumTau=numtau, pos_flat=pos_flat, flux_flat=flux_flat)
    _setup(libasora, c)
    try:
        phi, used, upd = _sweep(libasora, _cabi, c, 4)
        assert used == 4 and upd == ns * N ** 3
        grid, used2, _ = _sweep(libasora, _cabi, c, 2)
        assert used2 == 2
        # q_max > 127: the shared-memory sweep in its eight-octant form (what many sources at such radii get)
        octs, used1, upd1 = _sweep(libasora, _cabi, c, 0)
        assert used1 == 1 and upd1 == upd
    finally:
        libasora.device_close()
    _assert_close(phi, grid, "cluster sweep vs grid-cooperative sweep, full 128^3 box", rtol=1e-11)
    _assert_close(phi, octs, "cluster sweep vs eight-octant shared-memory sweep, full 128^3 box", rtol=1e-11)
    ref, _, n = _oracle(c)
    _assert_close(phi, ref, "cluster sweep vs oracle, full 128^3 box")
    if os.path.exists(REF_SO):
        L = ctypes.CDLL(REF_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        L.ref_device_init.argtypes = [ctypes.c_int, ctypes.c_int]
        L.ref_density_to_device.argtypes = [dp, ctypes.c_int]
        L.ref_photo_table_to_device.argtypes = [dp, dp, ctypes.c_int]
        L.ref_source_data_to_device.argtypes = [ctypes.POINTER(ctypes.c_int32), dp, ctypes.c_int]
        L.ref_do_all_sources.argtypes = [ctypes.c_double, dp, ctypes.c_double, ctypes.c_double, dp, dp, dp, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int]
        L.ref_copy_coldens.argtypes = [dp, ctypes.c_int]
        L.ref_zero_coldens.argtypes = [ctypes.c_int, ctypes.c_int]
        ref_phi, _ = run_reference(L, c, batch=6)
        _assert_close(phi, ref_phi, "cluster sweep vs the reference kernel, full 128^3 box")


def test_cluster_heating_and_deterministic():
    from pyc2ray_b200.lib import _cabi, libasora
    from tests.test_gpu_parity import _setup, _sweep, _assert_close
    from tests.test_heating import heat_case, _oracle_heat
    c = heat_case("multi_n32")
    ref_phi, ref_heat = _oracle_heat(c)[:2]
    _setup(libasora, c)
    try:
        libasora.heat_table_to_device(c["heat_thin"], c["heat_thick"], c["NumTau"])
        _cabi.check(_cabi.L.asora_set_sweep_variant(4))
        n3 = c["N"] ** 3
        xh = np.ascontiguousarray(c["xh"].ravel())

        def run():
            phi, heat = np.zeros(n3), np.zeros(n3)
            libasora.do_all_sources_heat(c["R"], np.zeros(1), c["sig"], c["dr"], np.zeros(1), xh, phi, heat, c["flux_flat"].size,
                                         c["N"], c["minlogtau"], c["dlogtau"], c["NumTau"])
            return phi, heat
        phi, heat = run()
        _assert_close(phi, ref_phi, "cluster sweep with heating: phi_ion")
        _assert_close(heat, ref_heat, "cluster sweep with heating: phi_heat")
        _cabi.check(_cabi.L.asora_set_deterministic(1))
        p1, h1 = run()
        p2, h2 = run()
        assert np.array_equal(p1, p2) and np.array_equal(h1, h2)
        _assert_close(p1, phi, "cluster sweep, deterministic phi_ion", rtol=1e-13, floor=1e-15)
        for shape in ((4, 256), (1, 512)):   # bit-identical for every cluster shape
            _cabi.check(_cabi.L.asora_set_cluster_shape(*shape))
            p3, h3 = run()
            assert np.array_equal(p1, p3) and np.array_equal(h1, h3)
    finally:
        _cabi.L.asora_set_cluster_shape(0, 0)
        _cabi.L.asora_set_deterministic(0)
        _cabi.L.asora_set_sweep_variant(0)
        libasora.device_close()

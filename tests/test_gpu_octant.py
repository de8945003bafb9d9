"""GPU parity of the mirror-image sweep (variant 3, csrc/sweep_octant.cu) against the oracle and against the one-cell-per-
thread sweep: every instantiated launch shape (octants per CTA, mirror images per thread, batch, threads), with and
without the de-duplication of plane cells, the z-face grid copies, sphere-only, accumulating sweeps."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import os
import re


def _instantiated_shapes():
    """(octants per CTA, images per thread, batch, threads, big) of every instantiated launch shape, read from the
    ASORA_OCT_SHAPES_* tables of csrc/sweep_octant.cu."""
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pyc2ray_b200", "csrc",
                            "sweep_octant.cu")).read()
    out = []
    for line in src.splitlines():
        if line.startswith("#define ASORA_OCT_SHAPES_"):
            for m in re.finditer(r"X\((\d+), (\d+), (\d+), (\d+), (\d+), (true|false)\)", line):
                no, op, ba, bl, mb, big = m.groups()
                out.append((int(no), int(op), int(ba), int(bl), big == "true"))
    return out


ALL = _instantiated_shapes()
assert len(ALL) >= 16, "could not read the shape tables of csrc/sweep_octant.cu"
SHAPES = [s[:4] for s in ALL]
BIG = [s[:4] for s in ALL if s[4]]


def _oracle(c):
    import oracle
    return oracle.asora_do_all_sources(c["R"], c["sig"], c["dr"], c["ndens"].ravel(), c["xh"].ravel(), c["pos_flat"],
                                       c["flux_flat"], c["N"], c["thin"], c["thick"], c["minlogtau"], c["dlogtau"], c["NumTau"])


@pytest.mark.parametrize("name", ["r_int5", "odd_n15_full", "multi_n32"])
@pytest.mark.parametrize("shape", SHAPES)
def test_octant_shapes_vs_oracle(name, shape):
    from pyc2ray_b200.lib import _cabi, libasora
    from tests.fields import make_case
    from tests.test_gpu_parity import _setup, _sweep, _assert_close
    c = make_case(name)
    ref, _, n = _oracle(c)
    _setup(libasora, c)
    try:
        for sphere_only in (0, 1):
            _cabi.check(_cabi.L.asora_set_sphere_only(sphere_only))
            for knobs in (0, 8, 4):  # bit 3: every entry as class A (no de-duplication of plane cells), bit 2: plan entry prefetched
                _cabi.check(_cabi.L.asora_set_octant_shape(*shape[:3], shape[3] | (knobs << 16)))
                phi, used, upd = _sweep(libasora, _cabi, c, 3)
                assert used == 3
                if not sphere_only:
                    assert upd == n
                _assert_close(phi, ref, f"{name} shape={shape} sphere_only={sphere_only} knobs={knobs}")
    finally:
        _cabi.L.asora_set_octant_shape(0, 0, 0, 0)
        _cabi.L.asora_set_tuning(0, 0)
        _cabi.L.asora_set_sphere_only(0)
        libasora.device_close()


def _r30_case():
    from tests.fields import f1_fields, tables, SIG
    from pyc2ray_b200.utils.sourceutils import format_sources, generate_test_sources
    N, ns, R = 112, 10, 30.0          # q_max = 52 <= N/2 - 1: mirror-symmetric; 53 levels as in the bench workload
    srcpos = generate_test_sources(N, ns, seed=100)
    flux = 10 ** np.random.default_rng(11).normal(0, 0.5, size=ns)
    ndens, xh = f1_fields(N, srcpos)
    thin, thick, dlogtau, _ = tables("bb1e5")
    pos_flat, flux_flat = format_sources(srcpos, flux)
    return dict(N=N, R=R, sig=SIG, dr=6e20, ndens=ndens, xh=xh, thin=thin, thick=thick, minlogtau=-20.0, dlogtau=dlogtau,
                NumTau=thin.size, pos_flat=pos_flat, flux_flat=flux_flat)


def test_octant_bench_radius_vs_oracle_and_variant1():
    """R = 30 (the bench radius): every large launch shape, with the (k,i,j)-ordered z-face copies on and off and eight log2
    copies, against the oracle; the automatic choice; accumulation on top of earlier rates; agreement with variant 1."""
    from pyc2ray_b200.lib import _cabi, libasora
    from tests.test_gpu_parity import _setup, _sweep, _assert_close
    c = _r30_case()
    N, ns, R = c["N"], c["flux_flat"].size, c["R"]
    ref, _, n = _oracle(c)
    _setup(libasora, c)
    try:
        v1, used, _ = _sweep(libasora, _cabi, c, 1)
        assert used == 1
        _assert_close(v1, ref, "variant 1 at N=112, R=30")
        for shape in BIG:
            if shape[0] != 8:
                continue  # all eight octants of R = 30 fit in one CTA; the split shapes are exercised below
            for knobs in (0, 2, 3, 8, 10, 6, 7, 5, 13):  # bit 1: z-face copies on, bit 0: log2 copies, bit 3: no de-duplication,
                                                     # bit 2: plan entry prefetched
                _cabi.check(_cabi.L.asora_set_octant_shape(*shape[:3], shape[3] | (knobs << 16)))
                phi, used, upd = _sweep(libasora, _cabi, c, 3)
                assert used == 3 and upd == n
                _assert_close(phi, ref, f"R=30 shape={shape} knobs={knobs}")
                _assert_close(phi, v1, f"R=30 shape={shape} knobs={knobs} vs variant 1", rtol=1e-11)
        for shape in [s for s in BIG if s[0] != 8]:
            for knobs in (0, 2):
                _cabi.check(_cabi.L.asora_set_octant_shape(*shape[:3], shape[3] | (knobs << 16)))
                phi, used, upd = _sweep(libasora, _cabi, c, 3)
                _assert_close(phi, ref, f"R=30 split shape={shape} knobs={knobs}")
        # automatic shape, sphere-only, z-face copies, two accumulating sweeps
        _cabi.check(_cabi.L.asora_set_octant_shape(0, 0, 0, 2 << 16))
        _cabi.check(_cabi.L.asora_set_sphere_only(1))
        phi, used, _ = _sweep(libasora, _cabi, c, 3)
        _cabi.check(_cabi.L.asora_set_sphere_only(0))
        _assert_close(phi, ref, "R=30 automatic shape, sphere only, z-face copies")
        _cabi.check(_cabi.L.asora_set_sweep_variant(3))
        _cabi.check(_cabi.L.asora_buffer_upload(_cabi.BUF_XH_AV, _cabi.dptr(np.ascontiguousarray(c["xh"].ravel()))))
        h = ns // 2
        _cabi.check(_cabi.L.asora_raytrace_device(R, c["sig"], c["dr"], 0, h, -20.0, c["dlogtau"], c["NumTau"], 1))
        _cabi.check(_cabi.L.asora_raytrace_device(R, c["sig"], c["dr"], h, ns - h, -20.0, c["dlogtau"], c["NumTau"], 0))
        acc = np.empty(N ** 3)
        _cabi.check(_cabi.L.asora_buffer_download(_cabi.BUF_PHI_ION, _cabi.dptr(acc)))
        _assert_close(acc, ref, "R=30, two accumulating mirror-image sweeps")
    finally:
        _cabi.L.asora_set_sweep_variant(0)
        _cabi.L.asora_set_octant_shape(0, 0, 0, 0)
        _cabi.L.asora_set_tuning(0, 0)
        _cabi.L.asora_set_sphere_only(0)
        libasora.device_close()


def test_octant_refuses_asymmetric_region_and_auto_falls_back():
    """Even mesh with q_max > N/2 - 1: variant 3 cannot be forced; the automatic choice still sweeps correctly."""
    from pyc2ray_b200.lib import _cabi, libasora
    from tests.fields import make_case
    from tests.test_gpu_parity import _setup, _sweep, _assert_close
    c = make_case("small_r5")
    ref, _, _ = _oracle(c)
    _setup(libasora, c)
    try:
        with pytest.raises(RuntimeError):
            _sweep(libasora, _cabi, c, 3)
        _cabi.L.asora_set_sweep_variant(0)
        phi, used, _ = _sweep(libasora, _cabi, c, 0)
        assert used in (1, 2)
        _assert_close(phi, ref, "automatic fallback")
    finally:
        _cabi.L.asora_set_sweep_variant(0)
        libasora.device_close()


def test_plan_cache_hits():
    """Repeated sweeps -- also with an automatic split into parts (the cache used to miss whenever parts > 1)
    and when alternating between full and sphere-only sweeps -- do not rebuild their plans."""
    from pyc2ray_b200.lib import _cabi, libasora
    from tests.fields import make_case
    from tests.test_gpu_parity import _setup, _sweep
    c = make_case("mid_n48_r14")
    _setup(libasora, c)
    try:
        _cabi.check(_cabi.L.asora_set_tuning(0, 4 << 20))  # parts = 4
        _sweep(libasora, _cabi, c, 1)
        _cabi.check(_cabi.L.asora_set_tuning(0, 0))
        _sweep(libasora, _cabi, c, 1)
        _cabi.check(_cabi.L.asora_set_sphere_only(1))
        _sweep(libasora, _cabi, c, 1)
        _cabi.check(_cabi.L.asora_set_sphere_only(0))
        before = _cabi.L.asora_plan_builds()
        for rep in range(3):
            _cabi.check(_cabi.L.asora_set_tuning(0, 4 << 20))
            _sweep(libasora, _cabi, c, 1)
            _cabi.check(_cabi.L.asora_set_tuning(0, 0))
            _sweep(libasora, _cabi, c, 1)
            _cabi.check(_cabi.L.asora_set_sphere_only(1))
            _sweep(libasora, _cabi, c, 1)
            _cabi.check(_cabi.L.asora_set_sphere_only(0))
        assert _cabi.L.asora_plan_builds() == before
    finally:
        _cabi.L.asora_set_tuning(0, 0)
        _cabi.L.asora_set_sphere_only(0)
        libasora.device_close()


def test_octant_quadrants_automatic_at_r40():
    """R = 40 on a 160^3 mesh (q_max = 70: half-spaces no longer fit two to an SM): the automatic selection takes the
    mirror-image sweep with quadrants as CTAs; against the oracle and against variant 1."""
    import ctypes
    from pyc2ray_b200.lib import _cabi, libasora
    from tests.fields import f1_fields, tables, SIG
    from tests.test_gpu_parity import _setup, _sweep, _assert_close
    from pyc2ray_b200.utils.sourceutils import format_sources, generate_test_sources
    N, ns, R = 160, 3, 40.0
    srcpos = generate_test_sources(N, ns, seed=40)
    flux = 10 ** np.random.default_rng(40).normal(0, 0.5, size=ns)
    ndens, xh = f1_fields(N, srcpos)
    thin, thick, dlogtau, _ = tables("bb1e5")
    pos_flat, flux_flat = format_sources(srcpos, flux)
    c = dict(N=N, R=R, sig=SIG, dr=6e20, ndens=ndens, xh=xh, thin=thin, thick=thick, minlogtau=-20.0, dlogtau=dlogtau,
             NumTau=thin.size, pos_flat=pos_flat, flux_flat=flux_flat)
    ref, _, n = _oracle(c)
    _setup(libasora, c)
    try:
        phi, used, upd = _sweep(libasora, _cabi, c, 0)
        assert used == 3 and upd == n
        v1, used1, _ = _sweep(libasora, _cabi, c, 1)
        assert used1 == 1
    finally:
        libasora.device_close()
    _assert_close(phi, ref, "R=40 automatic (quadrant CTAs) vs oracle")
    _assert_close(phi, v1, "R=40 automatic (quadrant CTAs) vs variant 1", rtol=1e-11)

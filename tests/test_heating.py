"""Photo-heating rates: the oracle's restatement of photorates.f90:118,124 / raytracing.f90:530,537
pinned by properties of the reference's own formulae and tables, and the CUDA path against the oracle."""
import os

import numpy as np
import pytest

from tests.fields import GOLDEN, make_case, tables


def heat_case(name):
    """A parity case on the Teff = 5e4 K tables, for which the golden fixture holds heating tables made by the
    reference's BlackBodySource.make_heat_table (tests/golden/make_golden.py)."""
    c = make_case(name)
    thin, thick, dlogtau, _ = tables("bb5e4")
    g = np.load(os.path.join(GOLDEN, "ref_heat_tables.npz"))
    c.update(thin=thin, thick=thick, dlogtau=dlogtau, NumTau=thin.size,
             heat_thin=np.ascontiguousarray(g["bb5e4_heat_thin"]), heat_thick=np.ascontiguousarray(g["bb5e4_heat_thick"]))
    return c


def _oracle_heat(c, nthreads=1):
    import oracle
    return oracle.asora_do_all_sources_heat(c["R"], c["sig"], c["dr"], c["ndens"].ravel(), c["xh"].ravel(), c["pos_flat"],
                                            c["flux_flat"], c["N"], c["thin"], c["thick"], c["heat_thin"], c["heat_thick"],
                                            c["minlogtau"], c["dlogtau"], c["NumTau"], nthreads=nthreads)


def test_oracle_heating_leaves_phi_unchanged_and_threads_agree():
    import oracle
    c = heat_case("multi_n32")
    phi, heat, n = _oracle_heat(c)
    ref, _, n0 = oracle.asora_do_all_sources(c["R"], c["sig"], c["dr"], c["ndens"].ravel(), c["xh"].ravel(), c["pos_flat"],
                                             c["flux_flat"], c["N"], c["thin"], c["thick"], c["minlogtau"], c["dlogtau"],
                                             c["NumTau"])
    assert n == n0
    np.testing.assert_array_equal(phi, ref)
    assert (heat >= 0).all() and ((heat > 0) == (phi > 0)).all()
    phi4, heat4, _ = _oracle_heat(c, nthreads=4)
    np.testing.assert_allclose(heat4, heat, rtol=1e-13, atol=0)
    np.testing.assert_allclose(phi4, phi, rtol=1e-13, atol=0)


def test_oracle_heating_thin_limit_is_the_table_ratio():
    """Optically thin cells (photorates.f90:121,124): phi_heat / phi_ion = H_thin(tau) / T_thin(tau) for a single
    source, i.e. the mean excess energy per ionisation of the unattenuated spectrum where tau -> 0."""
    c = heat_case("thin_n24")
    c["pos_flat"], c["flux_flat"] = c["pos_flat"][:3].copy(), c["flux_flat"][:1].copy()
    phi, heat, _ = _oracle_heat(c)
    m = phi > 0
    ratio = heat[m] / phi[m]
    expect = c["heat_thin"][0] / c["thin"][0]  # table entry 0 is tau = 0 (radiation/common.py:33-36)
    # tau <~ 1e-4 everywhere in this box: the thin tables have moved by less than that from their tau = 0 values
    np.testing.assert_allclose(ratio, expect, rtol=2e-3)
    # ... and the value is physical: h (<nu> - nu_HI) of a 5e4 K black body, a few eV
    assert 1.0 < expect / 1.602176634e-12 < 15.0


def test_oracle_heating_fortran_flavour_matches_asora_flavour():
    """The Fortran traversal (cube, plane by plane) and the ASORA traversal (octahedral shells) of the oracle agree
    on phi_heat like they do on phi_ion (SURVEY notes N1, N4, N5: single-precision constants, thin-cell argument)."""
    import oracle
    c = heat_case("bench_like_n32")
    phi_a, heat_a, _ = _oracle_heat(c)
    N = c["N"]
    out = oracle.fortran_do_all_sources(c["flux"], c["srcpos"], 1000, N, c["sig"], c["dr"], c["ndens"], c["xh"], 0.0,
                                        c["thin"], c["thick"], c["minlogtau"], c["dlogtau"], c["R"], NumTau=c["NumTau"],
                                        use_subbox=False, heat_thin=c["heat_thin"], heat_thick=c["heat_thick"])
    phi_f, heat_f = out[0], out[5]
    a = heat_a.reshape(N, N, N)
    assert ((a > 0) == (heat_f > 0)).all()
    m = a > 0
    np.testing.assert_allclose(heat_f[m], a[m], rtol=2e-6)
    np.testing.assert_allclose(phi_f[m], phi_a.reshape(N, N, N)[m], rtol=2e-6)


# ---- CUDA path ----------------------------------------------------------------------------------------------------

def _gpu_heat(c, variant=0, tuning=None):
    import pyc2ray_b200  # noqa: F401
    from pyc2ray_b200.lib import _cabi, libasora
    N = c["N"]
    libasora.device_init(N, 8)
    try:
        libasora.photo_table_to_device(c["thin"], c["thick"], c["NumTau"])
        libasora.heat_table_to_device(c["heat_thin"], c["heat_thick"], c["NumTau"])
        libasora.density_to_device(np.ascontiguousarray(c["ndens"].ravel()), N)
        libasora.source_data_to_device(c["pos_flat"], c["flux_flat"], c["flux_flat"].size)
        _cabi.check(_cabi.L.asora_set_sweep_variant(variant))
        if tuning:
            _cabi.check(_cabi.L.asora_set_tuning(*tuning))
        phi, heat = np.zeros(N ** 3), np.zeros(N ** 3)
        libasora.do_all_sources_heat(c["R"], np.zeros(1), c["sig"], c["dr"], np.zeros(1),
                                     np.ascontiguousarray(c["xh"].ravel()), phi, heat, c["flux_flat"].size, N,
                                     c["minlogtau"], c["dlogtau"], c["NumTau"])
        # the plain entry point afterwards must give the same phi and leave heating off
        phi2 = np.zeros(N ** 3)
        libasora.do_all_sources(c["R"], np.zeros(1), c["sig"], c["dr"], np.zeros(1), np.ascontiguousarray(c["xh"].ravel()),
                                phi2, c["flux_flat"].size, N, c["minlogtau"], c["dlogtau"], c["NumTau"])
    finally:
        _cabi.L.asora_set_sweep_variant(0)
        _cabi.L.asora_set_tuning(0, 0)
        libasora.device_close()
    return phi, heat, phi2


def _close(a, b, rtol, what):
    atol = 1e-12 * np.abs(b).max()
    bad = np.abs(a - b) > rtol * np.abs(b) + atol
    assert not bad.any(), f"{what}: {bad.sum()} cells differ, max rel {(np.abs(a - b) / np.maximum(np.abs(b), atol)).max():.2e}"


@pytest.mark.gpu
@pytest.mark.parametrize("name,variant,tuning", [
    ("multi_n32", 1, None), ("multi_n32", 1, (2, 256)), ("multi_n32", 2, None), ("bench_like_n32", 1, None),
    ("mid_n48_r14", 1, (1, 896)), ("mid_n48_r14", 1, (1, 896 | (8 << 16))), ("clip_full_n24", 2, None), ("thin_n24", 1, None)])
def test_gpu_heating_vs_oracle(name, variant, tuning):
    c = heat_case(name)
    phi, heat, phi2 = _gpu_heat(c, variant, tuning)
    ref_phi, ref_heat, _ = _oracle_heat(c)
    rtol = 1e-6 if name == "thin_n24" else 1e-9  # thin cells: cancellation in tau_out - tau_in (test_gpu_parity.py)
    _close(phi, ref_phi, rtol, f"{name} phi_ion")
    _close(heat, ref_heat, rtol, f"{name} phi_heat")
    _close(phi2, ref_phi, rtol, f"{name} phi_ion without heating")


@pytest.mark.gpu
def test_gpu_do_raytracing_returns_heating():
    import pyc2ray_b200 as p
    c = heat_case("multi_n32")
    p.device_init(c["N"], 8)
    try:
        p.photo_table_to_device(c["thin"], c["thick"])
        args = (c["dr"], c["flux"], c["srcpos"], True, 1000, 64, 1e-2, np.asfortranarray(c["ndens"]), np.asfortranarray(c["xh"]),
                c["thin"], c["thick"])
        tail = (c["minlogtau"], c["dlogtau"], c["R"], c["sig"])
        phi, heat = p.do_raytracing(*args, c["heat_thin"], c["heat_thick"], *tail, quiet=True)
        phi0, heat0 = p.do_raytracing(*args, np.zeros_like(c["thin"]), np.zeros_like(c["thin"]), *tail, quiet=True)
    finally:
        p.device_close()
    ref_phi, ref_heat, _ = _oracle_heat(c)
    N = c["N"]
    assert heat0 is None and heat.shape == (N, N, N)
    _close(phi.ravel(), ref_phi, 1e-9, "do_raytracing phi")
    _close(heat.ravel(), ref_heat, 1e-9, "do_raytracing heat")
    _close(phi0.ravel(), ref_phi, 1e-9, "do_raytracing phi (no heating)")


@pytest.mark.gpu
def test_gpu_heating_errors():
    from pyc2ray_b200.lib import _cabi, libasora
    c = heat_case("small_r5")
    libasora.device_init(c["N"], 8)
    try:
        with pytest.raises(RuntimeError, match="photo_table_to_device first"):
            libasora.heat_table_to_device(c["heat_thin"], c["heat_thick"], c["NumTau"])
        libasora.photo_table_to_device(c["thin"], c["thick"], c["NumTau"])
        with pytest.raises(RuntimeError, match="no heating tables"):
            _cabi.check(_cabi.L.asora_set_heating(1))
        with pytest.raises(RuntimeError, match="length of the photo tables"):
            libasora.heat_table_to_device(c["heat_thin"][:-1].copy(), c["heat_thick"][:-1].copy(), c["NumTau"] - 1)
    finally:
        libasora.device_close()


@pytest.mark.gpu
def test_gpu_heating_device_resident_entry_points():
    """asora_set_heating + asora_raytrace_device leave the heating rates in ASORA_BUF_PHI_HEAT (no host traffic)."""
    from pyc2ray_b200.lib import _cabi, libasora
    L, check = _cabi.L, _cabi.check
    c = heat_case("multi_n32")
    N = c["N"]
    libasora.device_init(N, 8)
    try:
        libasora.photo_table_to_device(c["thin"], c["thick"], c["NumTau"])
        libasora.heat_table_to_device(c["heat_thin"], c["heat_thick"], c["NumTau"])
        libasora.density_to_device(np.ascontiguousarray(c["ndens"].ravel()), N)
        libasora.source_data_to_device(c["pos_flat"], c["flux_flat"], c["flux_flat"].size)
        check(L.asora_buffer_upload(_cabi.BUF_XH_AV, _cabi.dptr(np.ascontiguousarray(c["xh"].ravel()))))
        check(L.asora_set_heating(1))
        check(L.asora_raytrace_device(c["R"], c["sig"], c["dr"], 0, c["flux_flat"].size, c["minlogtau"], c["dlogtau"], c["NumTau"], 1))
        phi, heat = np.empty(N ** 3), np.empty(N ** 3)
        check(L.asora_buffer_download(_cabi.BUF_PHI_ION, _cabi.dptr(phi)))
        check(L.asora_buffer_download(_cabi.BUF_PHI_HEAT, _cabi.dptr(heat)))
        with pytest.raises(RuntimeError, match="zero_phi"):
            check(L.asora_raytrace_device(c["R"], c["sig"], c["dr"], 0, 1, c["minlogtau"], c["dlogtau"], c["NumTau"], 0))
        check(L.asora_set_heating(0))
        # accumulating sweeps (zero_phi = 0) without heating: two halves of the list add up to the whole
        ns = c["flux_flat"].size
        check(L.asora_raytrace_device(c["R"], c["sig"], c["dr"], 0, ns // 2, c["minlogtau"], c["dlogtau"], c["NumTau"], 1))
        check(L.asora_raytrace_device(c["R"], c["sig"], c["dr"], ns // 2, ns - ns // 2, c["minlogtau"], c["dlogtau"], c["NumTau"], 0))
        phi_halves = np.empty(N ** 3)
        check(L.asora_buffer_download(_cabi.BUF_PHI_ION, _cabi.dptr(phi_halves)))
    finally:
        libasora.device_close()
    ref_phi, ref_heat, _ = _oracle_heat(c)
    _close(phi, ref_phi, 1e-9, "device-resident phi_ion")
    _close(heat, ref_heat, 1e-9, "device-resident phi_heat")
    _close(phi_halves, ref_phi, 1e-9, "phi_ion accumulated over two sweeps")


@pytest.mark.gpu
def test_gpu_pageable_grids_take_the_threaded_copy_path():
    """Grids above 8 MB in pageable (numpy) memory are staged by several host threads (asora_api.cu: host_copy);
    sizes that do not divide into whole bounce buffers, both memory orders, and pinned memory for comparison."""
    import torch
    from pyc2ray_b200.lib import _cabi, libasora
    L, check = _cabi.L, _cabi.check
    N = 113  # 113^3 * 8 B = 11.5 MB: three 4 MB chunks, the last one partial, unevenly spread over four threads
    rng = np.random.default_rng(5)
    a = rng.uniform(size=N ** 3)
    libasora.device_init(N, 8)
    try:
        b = np.empty_like(a)
        check(L.asora_buffer_upload(_cabi.BUF_XH, _cabi.dptr(a)))
        check(L.asora_buffer_download(_cabi.BUF_XH, _cabi.dptr(b)))
        np.testing.assert_array_equal(a, b)
        af = np.asfortranarray(a.reshape(N, N, N))
        bf = np.empty((N, N, N), order="F")
        check(L.asora_buffer_upload_f(_cabi.BUF_TEMP, _cabi.dptr(af)))
        check(L.asora_buffer_download_f(_cabi.BUF_TEMP, _cabi.dptr(bf)))
        np.testing.assert_array_equal(af, bf)
        c = np.empty(N ** 3)
        check(L.asora_buffer_download(_cabi.BUF_TEMP, _cabi.dptr(c)))
        np.testing.assert_array_equal(c.reshape(N, N, N), af)  # the device holds the logical C order
        pin = torch.empty(N ** 3, dtype=torch.float64).pin_memory()
        check(L.asora_buffer_download(_cabi.BUF_XH, ctypes_ptr(pin)))
        np.testing.assert_array_equal(pin.numpy(), a)
    finally:
        libasora.device_close()


def ctypes_ptr(t):
    import ctypes
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_double))

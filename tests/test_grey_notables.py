"""The reference's -D GREY_NOTABLES build (raytracing.cu:317-318, rates.cu:44-64; raytracing.f90:499-501,
photorates.f90:13-57): analytic grey-opacity rates instead of the table lookups.  CPU: the oracle's restatement against
the closed form along a grid axis and between its two traversals.  GPU: asora_set_grey_notables against the oracle and
against the reference's own kernel compiled with -D GREY_NOTABLES (oracle/_ref/libasora_ref_grey.so)."""
import ctypes
import os

import numpy as np
import pytest

import oracle
from tests.fields import make_case

REF_GREY_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libasora_ref_grey.so")
FOURPI = 12.566370614359172463991853874177


def _oracle_grey(c):
    phi, cdh, n = oracle.asora_do_all_sources(c["R"], c["sig"], c["dr"], c["ndens"].ravel(), c["xh"].ravel(), c["pos_flat"],
                                              c["flux_flat"], c["N"], c["thin"], c["thick"], c["minlogtau"], c["dlogtau"],
                                              c["NumTau"], grey_notables=True)
    return phi, cdh


def test_oracle_grey_rates_closed_form_along_an_axis():
    """Uniform medium, one source: along a grid axis the column is exact (d - 1/2 cells in, d + 1/2 out), so the rate of the
    cell at distance d is strength * 1e48 / (4 pi d^2 dr^3) * (exp(-tau_in) - exp(-tau_out)) / nHI."""
    N, R, dr, sig = 24, 9.0, 2.0e21, 6.30e-18
    nd, xh = np.full(N ** 3, 3e-4), np.full(N ** 3, 0.25)
    pos = np.array([12, 11, 13], dtype=np.int32)
    flux = np.array([0.7])
    dummy = np.zeros(4)
    phi, _, _ = oracle.asora_do_all_sources(R, sig, dr, nd, xh, pos, flux, N, dummy, dummy, -20.0, 1.0, 3, grey_notables=True)
    phi = phi.reshape(N, N, N)
    nhi = 3e-4 * 0.75
    for d in range(1, 9):
        tin, tout = (d - 0.5) * nhi * dr * sig, (d + 0.5) * nhi * dr * sig
        want = 0.7 * 1e48 / (FOURPI * d * d * dr ** 3) * (np.exp(-tin) - np.exp(-tout)) / nhi
        for cell in ((12 + d, 11, 13), (12, 11 - d, 13), (12, 11, 13 + d)):
            assert abs(phi[cell] - want) <= 1e-12 * want, (d, cell, phi[cell], want)
    # the source cell: path dr/2, volume dr^3 (raytracing.cu:285-294)
    t0 = 0.5 * nhi * dr * sig
    want0 = 0.7 * 1e48 / dr ** 3 * (1.0 - np.exp(-t0)) / nhi
    assert abs(phi[12, 11, 13] - want0) <= 1e-12 * want0
    # the tables were never read: a second run with different dummy tables gives the same bits
    phi2, _, _ = oracle.asora_do_all_sources(R, sig, dr, nd, xh, pos, flux, N, dummy + 5.0, dummy - 3.0, -20.0, 1.0, 3,
                                             grey_notables=True)
    assert np.array_equal(phi.ravel(), phi2)


def test_oracle_grey_fortran_and_asora_traversals_agree():
    c = make_case("small_r5")
    phi_a, _ = _oracle_grey(c)
    N = c["N"]
    out = oracle.fortran_do_all_sources(c["flux"], c["srcpos"], N, N, c["sig"], c["dr"], c["ndens"], c["xh"], 0.0, c["thin"], c["thick"],
                                        c["minlogtau"], c["dlogtau"], c["R"], NumTau=c["NumTau"], use_subbox=False,
                                        grey_notables=True)
    phi_f = np.ascontiguousarray(out[0]).ravel()
    inside = phi_a != 0          # the Fortran also rates cells outside the octahedron's reach of the CUDA sweep
    np.testing.assert_allclose(phi_f[inside], phi_a[inside], rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["small_r5", "clip_full_n24", "odd_n15_full", "multi_n32", "thin_n24"])
def test_grey_notables_vs_oracle_and_reference_kernel(name):
    from pyc2ray_b200.lib import _cabi, libasora
    from tests.test_gpu_vs_reference_kernel import _close
    c = make_case(name)
    ref, _ = _oracle_grey(c)
    libasora.device_init(c["N"], 8)
    try:
        # no tables uploaded: the grey rates must not need them
        libasora.density_to_device(np.ascontiguousarray(c["ndens"].ravel()), c["N"])
        libasora.source_data_to_device(c["pos_flat"], c["flux_flat"], c["flux_flat"].size)
        _cabi.check(_cabi.L.asora_set_grey_notables(1))
        phi = np.zeros(c["N"] ** 3)
        libasora.do_all_sources(c["R"], np.zeros(1), c["sig"], c["dr"], np.zeros(1), np.ascontiguousarray(c["xh"].ravel()), phi,
                                c["flux_flat"].size, c["N"], c["minlogtau"], c["dlogtau"], c["NumTau"])
        used = ctypes.c_int(0)
        _cabi.L.asora_last_sweep_stats(ctypes.byref(used), None, None, None, None, None)
        assert used.value == 2
        # with the switch off and no tables the call must fail loudly, not fall back
        _cabi.check(_cabi.L.asora_set_grey_notables(0))
        with pytest.raises(RuntimeError):
            libasora.do_all_sources(c["R"], np.zeros(1), c["sig"], c["dr"], np.zeros(1), np.ascontiguousarray(c["xh"].ravel()),
                                    np.zeros(c["N"] ** 3), c["flux_flat"].size, c["N"], c["minlogtau"], c["dlogtau"], c["NumTau"])
    finally:
        libasora.device_close()
    assert ((phi != 0) == (ref != 0)).all()
    rtol = 1e-6 if name == "thin_n24" else 1e-9   # thin cells: cancellation in tau_out - tau_in
    _close(phi, ref, f"{name}: grey rates vs oracle", rtol=rtol)
    if os.path.exists(REF_GREY_SO):
        from tests.test_gpu_vs_reference_kernel import run_reference
        L = ctypes.CDLL(REF_GREY_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        L.ref_device_init.argtypes = [ctypes.c_int, ctypes.c_int]
        L.ref_density_to_device.argtypes = [dp, ctypes.c_int]
        L.ref_photo_table_to_device.argtypes = [dp, dp, ctypes.c_int]
        L.ref_source_data_to_device.argtypes = [ctypes.POINTER(ctypes.c_int32), dp, ctypes.c_int]
        L.ref_do_all_sources.argtypes = [ctypes.c_double, dp, ctypes.c_double, ctypes.c_double, dp, dp, dp, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int]
        L.ref_copy_coldens.argtypes = [dp, ctypes.c_int]
        L.ref_zero_coldens.argtypes = [ctypes.c_int, ctypes.c_int]
        # batch 1: under GREY_NOTABLES the reference reads coldensh_out without the batch offset (raytracing.cu:318), which is
        # only the cell's own value for the first source of a batch
        phi_ref, _ = run_reference(L, c, batch=1)
        _close(ref, phi_ref, f"{name}: oracle grey rates vs the reference kernel", rtol=rtol)
        _close(phi, phi_ref, f"{name}: grey rates vs the reference kernel", rtol=rtol)

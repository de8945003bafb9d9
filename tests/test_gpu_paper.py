"""GPU end-to-end tests: the reference's paper / regression tests run through pyc2ray_b200.evolve3D
(device-resident ray tracing + chemistry loop), checked against the reference's printed known answers,
the analytic Stroemgren solution, and the CPU oracle with the reference's own tolerance ladder."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests.fields import GOLDEN

KAT = json.load(open(os.path.join(GOLDEN, "kat.json")))


@pytest.mark.parametrize("tag,Teff,grey", [("grey", 5e4, True), ("Teff5e3", 5e3, False), ("Teff5e4", 5e4, False),
                                           ("Teff1e5", 1e5, False)])
def test_paper_test3_multisource_known_answers(tag, Teff, grey):
    """test/paper_tests/test3_multisource/make_plot.ipynb cell 5: volume-mean x after 10 x 1 Myr for four
    spectra, printed to 7-8 digits.  pyC2Ray itself differs from the original C2Ray by ~1e-6 relative;
    we require 5e-6 against pyC2Ray's numbers."""
    from tests.paper_tests import run_test3_multisource
    k = KAT["test3_multisource_mean_x"]
    ref = k["pyc2ray"][k["order"].index(tag)]
    mean_x, _ = run_test3_multisource("gpu", Teff=Teff, grey=grey)
    assert abs(mean_x - ref) / ref < 5e-6, (tag, mean_x, ref)


def test_paper_test1_stromgren_256_within_reference_band():
    """test/paper_tests/test1_Ifront, coarse time steps (10 x 50 Myr) at the full 256^3: r_N / r_A must stay in
    the band the reference plots, [0.985, 1.005] (make_plot.ipynb cell 10)."""
    from tests.paper_tests import run_test1_stromgren
    t, ratio, (r_S, t_rec) = run_test1_stromgren("gpu", N=256, nsteps=10)
    lo, hi = KAT["test1_stromgren"]["band"]
    assert np.all(ratio[1:] >= lo) and np.all(ratio[1:] <= hi), ratio


def test_hackathon_test1_tolerance_ladder_vs_oracle():
    """test/unit_tests_hackathon/1_single_black_body/run_test.py:91-115 with the CPU oracle standing in for
    the missing c2ray_xfrac_reference.refbin."""
    import oracle
    from tests.paper_tests import run_hackathon_test1
    x_gpu, phi_gpu = run_hackathon_test1("gpu")
    x_cpu, phi_cpu = run_hackathon_test1("oracle", nthreads=1)
    tol = KAT["hackathon_test1_tolerances"]
    abserr = x_gpu - x_cpu
    relerr = abserr / x_cpu
    assert abs(abserr.mean()) <= tol["abs"]["mean"] and abserr.std() <= tol["abs"]["std"]
    assert np.abs(abserr).max() <= tol["abs"]["max"]
    assert abs(relerr.mean()) <= tol["rel"]["mean"] and relerr.std() <= tol["rel"]["std"]
    assert np.abs(relerr).max() <= tol["rel"]["max"]
    assert x_gpu.mean() > 0.01  # the source did ionise a bubble
    # far tighter than the reference's own ladder: fp64 end to end
    np.testing.assert_allclose(x_gpu, x_cpu, rtol=1e-8, atol=1e-14)


def test_driver_class_runs_reference_style_script():
    """The reference's driver pattern (test/paper_tests/test3_multisource/run_test.py:21-66) on the C2Ray_Test
    class of this package: same calls, same known answer."""
    import pyc2ray_b200 as pc2r
    here = os.path.dirname(os.path.abspath(__file__))
    sim = pc2r.C2Ray_Test(os.path.join(here, "golden", "params_test3.yml"), 128, True)
    try:
        zred_array = sim.generate_redshift_array(2, 1e7)
        srcpos, srcflux = sim.read_sources(os.path.join(here, "golden", "src_test3.txt"), 5)
        for k in range(len(zred_array) - 1):
            zi, zf = zred_array[k], zred_array[k + 1]
            dt = sim.set_timestep(zi, zf, 10)
            sim.set_constant_average_density(1.0e-6, 0)
            sim.zred = zi
            for t in range(10):
                sim.cosmo_evolve(dt)
                sim.evolve3D(dt, srcflux, srcpos)
        sim.write_output(zf)
        assert f"{zf:.3f}" == "8.835"
        k3 = KAT["test3_multisource_mean_x"]
        ref = k3["pyc2ray"][k3["order"].index("Teff5e4")]
        assert abs(sim.xh.mean() - ref) / ref < 5e-6, sim.xh.mean()
        assert sim.phi_ion.shape == (128, 128, 128)
    finally:
        sim._gpu_close()


def test_paper_test2_cosmological_ifront():
    """test/paper_tests/test2_Ifront_cosmo through the driver class (cosmological: 1 -> density dilution and
    a new cell size, hence a rebuilt sweep plan, every step): comoving I-front radius against the analytic
    solution y(t) = lam e^{lam t_i/t} [t/t_i E2(lam t_i/t) - E2(lam)] (make_plot.ipynb cell 5).  128^3 instead of
    256^3, so the reference's band [0.985, 1.005] is widened to +-2 %."""
    from scipy.special import expn
    import pyc2ray_b200 as pc2r
    here = os.path.dirname(os.path.abspath(__file__))
    N, numzred, t_evol = 128, 10, 5e8
    sim = pc2r.C2Ray_Test(os.path.join(here, "golden", "params_test2.yml"), N, True)
    try:
        zred_array = sim.generate_redshift_array(numzred + 1, t_evol / numzred)
        srcpos = np.array([[N // 2], [N // 2], [N // 2]])
        srcflux = np.array([1e54 / 1e48])
        fronts = []
        boxsize = 22685.455026110553 / 10.0          # kpc, a = 1 at z = 9 (make_plot.ipynb cell 1)
        x = np.linspace(0, boxsize / 2, N // 2 + 1)
        for k in range(numzred):
            zi, zf = zred_array[k], zred_array[k + 1]
            dt = sim.set_timestep(zi, zf, 1)
            sim.density_init(zi)
            sim.zred = zi
            sim.cosmo_evolve(dt)
            sim.evolve3D(dt, srcflux, srcpos)
            prof = np.asarray(sim.xh)[N // 2 - 1:, N // 2 - 1, N // 2 - 1]
            fronts.append(np.interp(0.5, np.flip(prof), np.flip(x[:prof.size])))
    finally:
        sim._gpu_close()
    kk = KAT["test2_cosmo_Ifront"]
    ti = kk["age_z9_Myr"]
    nH, alpha_B, kpc, year = 1.87e-4, 2.59e-13, 3.086e21, 3.15576e7
    r_S = ((3 * 1e54) / (4 * np.pi * alpha_B * nH ** 2)) ** (1. / 3) / kpc
    lam = ti / (1.0 / (alpha_B * nH * year * 1e6))
    t = ti + np.linspace(50, 500, 10)
    y = lam * np.exp(lam * ti / t) * (t / ti * expn(2, lam * ti / t) - expn(2, lam))
    ratio = np.array(fronts) / (r_S * y ** (1. / 3))
    assert np.all(np.abs(ratio[1:] - 1.0) < 0.02), ratio

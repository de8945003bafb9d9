"""The 244 Mpc/h driver class (pyc2ray_b200/c2ray_244paper.py; reference: pyc2ray/c2ray_244paper.py): host-side conventions on
CPU, and a miniature run + resume through the GPU path."""
import os

import numpy as np
import pytest
import yaml

from pyc2ray_b200.utils import c2ray_files as cf


def _bare():
    """An instance without a device: only the host-side methods are exercised."""
    from pyc2ray_b200.c2ray_244paper import C2Ray_244Test
    from pyc2ray_b200.cosmology import FlatLambdaCDM
    sim = object.__new__(C2Ray_244Test)
    sim._ld = {"Grid": {"boxsize": 244.0}, "Cosmology": {"h": 0.7}}
    sim.cosmology = FlatLambdaCDM(70.0, 0.27, 2.726, Ob0=0.044)
    sim.fgamma_hm, sim.fgamma_lm = 30.0, 0.0
    sim.zred_0 = 21.062
    sim.age_0 = 2. * (1. + sim.zred_0) ** (-1.5) / (3. * 70.0 * 1e5 / 3.086e24 * np.sqrt(0.27))
    sim.logfile = os.devnull
    return sim


def test_tools21cm_conversion_factors():
    from pyc2ray_b200.c2ray_244paper import C2Ray_244Test as T
    # rho_crit_0 for h = 0.7 with tools21cm's constants: 1.88e-29 h^2 g/cm^3 to three digits
    assert abs(T._rho_crit_0() / (1.8791e-29 * 0.49) - 1.0) < 2e-3
    # a 250^3 coarse cell holds (8000/250)^3 fine cells: mean raw value 32^3 * Omega0/... -> density factor scales as mesh^3
    assert np.isclose(T.gridmass_to_cgs_density(250) / T.gridmass_to_cgs_density(125), 8.0)
    sim = _bare()
    # mass of the whole box in fine-grid units = nbox_fine^3  ->  rho_matter * L^3 in solar masses
    L = 244.0 / 0.7 * 3.086e24
    assert np.isclose(sim.gridmass_to_msun() * 8000.0 ** 3, T._rho_crit_0() * 0.27 * L ** 3 * 5.02785431e-34, rtol=1e-12)
    # a uniform box at the cosmic mean: raw = (nbox_fine/mesh)^3 * Omega0-weighted mass units per coarse cell gives rho_b
    raw_mean = (8000.0 / 250.0) ** 3
    assert np.isclose(raw_mean * T.gridmass_to_cgs_density(250), T._rho_crit_0() * 0.044, rtol=1e-12)


def test_time_redshift_relations_and_timestep():
    sim = _bare()
    for z in (21.062, 15.0, 9.5, 6.0):
        assert np.isclose(sim.time2zred(sim.zred2time(z)), z, rtol=1e-13)
    assert np.isclose(sim.zred2time(sim.zred_0), sim.age_0)
    dt = sim.set_timestep(12.0, 11.5, 2)
    assert dt > 0 and np.isclose(2 * dt, sim.zred2time(11.5) - sim.zred2time(12.0))


def test_text_catalogue_and_density_file(tmp_path):
    sim = _bare()
    cat = tmp_path / "9.938-coarsest_wsubgrid_sources.dat"
    rows = np.array([[3, 4, 5, 120.0, 7.0], [10, 1, 24, 35.5, 0.0], [7, 7, 7, 980.0, 12.0]])
    with open(cat, "w") as f:
        f.write("3\n")
        for r in rows:
            f.write("%d %d %d %.6e %.6e\n" % tuple(r))
    ts = 20e6 * 3.15576e7
    srcpos, normflux = sim.read_sources(str(cat), mass="hm", ts=ts)
    assert srcpos.shape == (3, 3) and (srcpos == rows[:, :3].T).all()
    mass_msun = rows[:, 3] * sim.gridmass_to_msun()
    _, expect = cf.sources_from_catalogue(rows[:, :3], mass_msun, 30.0, 0.044, 0.27, ts)
    np.testing.assert_allclose(normflux, expect, rtol=1e-13)
    _, lm = sim.read_sources(str(cat), mass="lm", ts=ts)
    assert lm[1] == 0.0 and lm[0] > 0
    # density directory
    d = tmp_path / "coarser_densities"
    d.mkdir()
    rng = np.random.default_rng(0)
    for z in (9.938, 10.110, 10.290):
        cf.save_cbin(str(d / ("%.3fn_all.dat" % z)), rng.uniform(1e4, 5e4, size=(6, 6, 6)), bits=32, order="F")
    from pyc2ray_b200.c2ray_244paper import C2Ray_244Test as T
    np.testing.assert_allclose(T.get_dens_redshifts(str(d)), [9.938, 10.110, 10.290])


PARAMS = """
Grid: {boxsize: 244, resume: %(resume)d}
Material: {temp0: 1e4, xh0: 2.0e-4, avg_dens: 1.981e-07}
CGS: {albpow: -0.7, bh00: 2.59e-13, alcpow: -0.672, eth0: 13.598, ethe0: 24.587, ethe1: 54.416, xih0: 1.0, fh0: 0.83, colh0_fact: 1.3e-8}
Abundances: {abu_h: 0.926, abu_he: 0.074, abu_c: 7.1e-7}
Photo: {sigma_HI_at_ion_freq: 6.30e-18, minlogtau: -20, maxlogtau: 4, NumTau: 2000, grey: 0, SourceType: blackbody,
        compute_heating_rates: 0, R_max_cMpc: 60.0}
BlackBodySource: {Teff: 5e4, cross_section_pl_index: 2.8}
Sources: {fgamma_hm: 30, fgamma_lm: 0., ts: 20.0}
Cosmology: {cosmological: 1, h: 0.7, Omega0: 0.27, Omega_B: 0.044, cmbtemp: 2.726, zred_0: 12.0}
Output: {results_basename: %(out)s/, inputs_basename: %(inp)s/, logfile: pyC2Ray.log}
Raytracing: {loss_fraction: 1e-2, subboxsize: 5, max_subbox: 1000, source_batch_size: 8, convergence_fraction: 1e-4}
"""


@pytest.mark.gpu
def test_miniature_244_run_and_resume(tmp_path):
    """The time loop of test/paper_eor_simulation/run_test.py on a 24^3 mesh with synthetic inputs in the 244 Mpc layout
    (coarse-grained density files, text source catalogues), then a resumed run from the written cbin files."""
    import pyc2ray_b200 as pc2r
    from pyc2ray_b200.c2ray_244paper import C2Ray_244Test
    N = 24
    inp, out = tmp_path / "inputs", tmp_path / "results"
    (inp / "coarser_densities").mkdir(parents=True)
    (inp / "sources").mkdir()
    out.mkdir()
    zs = [12.0, 11.7, 11.4]
    rng = np.random.default_rng(5)
    raw_mean = (8000.0 / N) ** 3
    for z in zs:
        cf.save_cbin(str(inp / "coarser_densities" / ("%.3fn_all.dat" % z)),
                     raw_mean * np.exp(rng.normal(size=(N, N, N)) * 0.5 - 0.125), bits=32, order="F")
        with open(inp / "sources" / ("%.3f-coarsest_wsubgrid_sources.dat" % z), "w") as f:
            f.write("5\n")
            for s in range(5):
                i, j, k = rng.integers(1, N + 1, size=3)
                f.write("%d %d %d %.6e 0.0\n" % (i, j, k, 10 ** rng.uniform(5.5, 6.5)))
    par = tmp_path / "parameters.yml"
    par.write_text(PARAMS % dict(resume=0, out=out, inp=inp))
    sim = C2Ray_244Test(paramfile=str(par), Nmesh=N, use_gpu=True)
    try:
        assert np.isclose(sim.R_max_LLS, 60.0 * N * 0.7 / 244.0)
        steps = 2
        for k in range(len(zs) - 1):
            zi, zf = zs[k], zs[k + 1]
            dt = sim.set_timestep(zi, zf, steps)
            sim.write_output(zi)
            sim.read_density(z=zi)
            srcpos, normflux = sim.read_sources(file="%ssources/%.3f-coarsest_wsubgrid_sources.dat" % (sim.inputs_basename, zi),
                                                mass="hm", ts=steps * dt)
            sim.zred = zi
            for t in range(steps):
                sim.cosmo_evolve(dt)
                sim.evolve3D(dt, normflux, srcpos)
            sim.cosmo_evolve_to_now()
            assert np.isclose(sim.zred, zf, rtol=1e-10)
        sim.write_output(zs[-1])
        x_end, phi_end, mean_dens = sim.xh.copy(), sim.phi_ion.copy(), sim.ndens.mean()
    finally:
        sim._gpu_close()
    assert x_end.max() > 0.5 and x_end.min() >= 2.0e-4 * 0.99 and np.isfinite(phi_end).all()   # the sources ionised their surroundings
    # mean number density follows the comoving baryon density at the final redshift
    rho_b = C2Ray_244Test._rho_crit_0() * 0.044 / ((0.926 + 4 * 0.074) * cf.M_P)
    assert abs(mean_dens / (rho_b * (1 + zs[-1]) ** 3) - 1.0) < 0.1
    # resume: picks the lowest redshift with an xfrac file and restores the grids from the cbin files
    par.write_text(PARAMS % dict(resume=1, out=out, inp=inp))
    sim2 = C2Ray_244Test(paramfile=str(par), Nmesh=N, use_gpu=True)
    try:
        assert np.isclose(sim2.zred, zs[-1])
        np.testing.assert_array_equal(sim2.xh, x_end)                      # 64-bit file
        np.testing.assert_allclose(sim2.phi_ion, phi_end, rtol=1e-6)       # 32-bit file
        assert sim2.xh.flags.f_contiguous
    finally:
        sim2._gpu_close()
    assert "Resuming" in open(out / "pyC2Ray.log").read()

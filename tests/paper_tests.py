"""The reference's self-contained paper tests, restated without astropy / tools21cm
(test/paper_tests/test1_Ifront, test3_multisource) and runnable on either backend:
  backend="gpu"     pyc2ray_b200.evolve3D (the product path, through the C ABI)
  backend="oracle"  the evolve3D loop of pyc2ray/evolve.py:125-245 with the CPU oracle in both roles
"""
import numpy as np

YEAR = 3.15576e7          # c2ray_base.py:75
MPC = 3.086e24            # c2ray_base.py:74-79
EV2K = 1.0 / 8.617e-05    # c2ray_base.py:77
SIG = 6.30e-18
CHEM = dict(bh00=2.59e-13, albpow=-0.7, colh0=1.3e-8 * 0.83 * 1.0 / 13.598 ** 2, temph0=13.598 * EV2K, abu_c=7.1e-7)


def make_tables(Teff, grey, NumTau):
    from pyc2ray_b200.radiation import blackbody_tables
    thin, thick, dlogtau = blackbody_tables(Teff, grey, -20.0, 4.0, NumTau)
    return thin, thick, dlogtau


def oracle_evolve3D(dt, dr, src_flux, src_pos, temp, ndens, xh, thin, thick, minlogtau, dlogtau, R, conv_frac, sig,
                    nthreads=1):
    """pyc2ray/evolve.py:125-245 on the CPU oracle (ASORA flavour ray tracer + chemistry)."""
    import oracle
    from pyc2ray_b200.utils.sourceutils import format_sources
    N = temp.shape[0]
    NumSrc = src_flux.shape[0]
    pos_flat, flux_flat = format_sources(src_pos, src_flux)
    conv_criterion = min(int(conv_frac * N ** 3), (NumSrc - 1) / 3)
    prev1 = prev0 = 2 * N ** 3
    ndens_c, temp_c, xh_c = (np.ascontiguousarray(a, dtype=np.float64) for a in (ndens, temp, xh))
    xh_av, xh_int = xh_c.copy(), xh_c.copy()
    niter = 0
    while True:
        niter += 1
        phi, _, _ = oracle.asora_do_all_sources(R, sig, dr, ndens_c.ravel(), xh_av.ravel(), pos_flat, flux_flat, N, thin,
                                                thick, minlogtau, dlogtau, thin.size, nthreads=nthreads)
        phi = phi.reshape(N, N, N)
        flag = oracle.global_pass(dt, ndens_c, temp_c, xh_c, xh_av, xh_int, phi, CHEM["bh00"], CHEM["albpow"],
                                  CHEM["colh0"], CHEM["temph0"], CHEM["abu_c"])
        s1, s0 = xh_int.sum(), (1.0 - xh_int).sum()
        r1 = abs((s1 - prev1) / s1) if s1 > 0 else 1.0
        r0 = abs((s0 - prev0) / s0) if s0 > 0 else 1.0
        prev1, prev0 = s1, s0
        if flag < conv_criterion or (r1 < conv_frac and r0 < conv_frac):
            return xh_int, phi, niter


def _evolve(backend, dt, dr, flux, pos, temp, ndens, xh, thin, thick, dlogtau, R, nthreads):
    if backend == "gpu":
        import pyc2ray_b200 as p
        return p.evolve3D(dt, dr, flux, pos, True, 1000, 64, 1e-2, temp, ndens, xh, thin, thick, -20.0, dlogtau, R,
                          1e-4, SIG, CHEM["bh00"], CHEM["albpow"], CHEM["colh0"], CHEM["temph0"], CHEM["abu_c"],
                          logfile=None, quiet=True)[:2]
    x, phi, _ = oracle_evolve3D(dt, dr, flux, pos, temp, ndens, xh, thin, thick, -20.0, dlogtau, R, 1e-4, SIG, nthreads)
    return x, phi


def run_test3_multisource(backend, Teff=5e4, grey=False, N=128, nsteps=10, nthreads=1):
    """test/paper_tests/test3_multisource: 128^3, box 14 kpc, n_H = 1e-3, 5 sources of 5e48 /s, 10 x 1 Myr.
    Returns the volume-mean ionised fraction (known answers: tests/golden/kat.json)."""
    dr = 0.014 * MPC / N
    pos = np.array([[64, 64, 64], [32, 96, 64], [32, 32, 64], [96, 32, 64], [96, 96, 64]], dtype=np.int64).T
    if N != 128:
        pos = np.maximum(1, (pos * N) // 128)
    flux = np.full(5, 5e48 / 1e48)
    ndens = np.full((N, N, N), 1.0e-6 * (1 + 9.0) ** 3, order="F")
    temp = np.full((N, N, N), 1e4, order="F")
    xh = np.full((N, N, N), 1.2e-3, order="F")
    thin, thick, dlogtau = make_tables(Teff, grey, 10000)
    R = 15.0 * N / 0.014
    dt = 1e7 * YEAR / nsteps
    if backend == "gpu":
        import pyc2ray_b200 as p
        p.device_init(N, 8)
        p.photo_table_to_device(thin, thick)
    try:
        for _ in range(nsteps):
            xh, phi = _evolve(backend, dt, dr, flux, pos, temp, ndens, xh, thin, thick, dlogtau, R, nthreads)
    finally:
        if backend == "gpu":
            p.device_close()
    return float(np.mean(xh)), xh


def run_test1_stromgren(backend, N=256, nsteps=10, t_evol_yr=5e8, nthreads=1):
    """test/paper_tests/test1_Ifront: grey Stroemgren sphere, 1 source of 1e54 /s in n_H = 1.87e-4, box 1.62 Mpc.
    Returns (times_Myr, front radius / analytic radius) after each step (make_plot.ipynb cells 5-9)."""
    boxsize_mpc = 1.62022035
    dr = boxsize_mpc * MPC / N
    c = N // 2
    pos = np.array([[c, c, c]], dtype=np.int64).T  # 1-indexed; (128,128,128) at N = 256
    flux = np.array([1e54 / 1e48])
    nH = 1.87e-7 * (1 + 9.0) ** 3
    ndens = np.full((N, N, N), nH, order="F")
    temp = np.full((N, N, N), 1e4, order="F")
    xh = np.full((N, N, N), 1.2e-3, order="F")
    thin, thick, dlogtau = make_tables(5e4, True, 20000)
    R = 15.0 * N / boxsize_mpc
    dt = t_evol_yr * YEAR / nsteps
    kpc = 3.086e21
    r_S = ((3 * 1e54) / (4 * np.pi * 2.59e-13 * nH ** 2)) ** (1. / 3) / kpc
    t_rec = 1.0 / (2.59e-13 * nH * YEAR * 1e6)
    xcoord = np.linspace(0, boxsize_mpc * 1e3 / 2, N // 2 + 1)  # kpc, make_plot.ipynb cell 6
    out_t, out_ratio = [], []
    if backend == "gpu":
        import pyc2ray_b200 as p
        p.device_init(N, 1)
        p.photo_table_to_device(thin, thick)
    try:
        for k in range(nsteps):
            xh, phi = _evolve(backend, dt, dr, flux, pos, temp, ndens, xh, thin, thick, dlogtau, R, nthreads)
            prof = np.asarray(xh)[c - 1:, c - 1, c - 1]
            front = np.interp(0.5, np.flip(prof), np.flip(xcoord[:prof.size]))
            t_myr = (k + 1) * t_evol_yr / nsteps / 1e6
            r_an = r_S * (1.0 - np.exp(-t_myr / t_rec)) ** (1. / 3)
            out_t.append(t_myr)
            out_ratio.append(front / r_an)
    finally:
        if backend == "gpu":
            p.device_close()
    return np.array(out_t), np.array(out_ratio), (r_S, t_rec)


def run_hackathon_test1(backend, N=128, nsteps=10, nthreads=1):
    """test/unit_tests_hackathon/1_single_black_body: one source (96,96,64) of 1e49 /s in a uniform
    n_H = 1e-3 box of 14 kpc, Teff = 5e4 K, NumTau = 10000, 10 x 1 Myr.  Returns the final xh."""
    dr = 0.014 * MPC / N
    pos = np.array([[96, 96, 64]], dtype=np.int64).T
    flux = np.array([10e48 / 1e48])
    ndens = np.full((N, N, N), 1e-3, order="F")
    temp = np.full((N, N, N), 1e4, order="F")
    xh = np.full((N, N, N), 1.2e-3, order="F")
    thin, thick, dlogtau = make_tables(5e4, False, 10000)
    R = 0.01640625 * N / 0.014
    dt = 1e7 * YEAR / nsteps
    if backend == "gpu":
        import pyc2ray_b200 as p
        p.device_init(N, 1)
        p.photo_table_to_device(thin, thick)
    try:
        for _ in range(nsteps):
            xh, phi = _evolve(backend, dt, dr, flux, pos, temp, ndens, xh, thin, thick, dlogtau, R, nthreads)
    finally:
        if backend == "gpu":
            p.device_close()
    return np.asarray(xh), np.asarray(phi)

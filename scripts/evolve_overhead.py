"""Where the fixed cost of one evolve3D call goes (250^3, 10^5 sources)."""
import cProfile, pstats, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ASORA_QUIET"] = "1"
import pyc2ray_b200 as p
N, nsrc = 250, 100000
rng = np.random.default_rng(244)
srcpos = p.generate_test_sources(N, nsrc, seed=244)
flux = 10 ** rng.normal(5.0, 0.5, size=nsrc)
ndens = 1.87e-4 * np.exp(0.5 * rng.normal(size=(N, N, N)) - 0.125)
xh = np.full((N, N, N), 2e-4); temp = np.full((N, N, N), 1e4)
thin, thick, dlogtau = p.blackbody_tables(1e5, False, -20.0, 4.0, 20000)
dr = 244.0 / 0.7 * 3.086e24 / N / 10.0; R = 15.0 * N * 0.7 / 244.0
chem = (2.59e-13, -0.7, 1.3e-8 * 0.83 / 13.598 ** 2, 13.598 / 8.617e-5, 7.1e-7)
p.device_init(N, 96); p.photo_table_to_device(thin, thick)
def step():
    return p.evolve3D(1e7 * 3.15576e7, dr, flux, srcpos, True, 1000, 64, 1e-2, temp, ndens, xh, thin, thick, -20.0, dlogtau,
                      R, 1e-4, 6.3e-18, *chem, logfile=None, quiet=True)
step()
t0 = time.perf_counter(); step(); print("evolve3D wall", time.perf_counter() - t0, "iterations", p.evolve3D.last_niter)
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
p.device_close()

"""Seeded synthetic fields and sources for the parity tests (fields F0 / F1)."""
import os

import numpy as np

from pyc2ray_b200.utils.sourceutils import format_sources, generate_test_sources

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MPC = 3.086e24
SIG = 6.30e-18
MINLOGTAU, MAXLOGTAU = -20.0, 4.0

_tables = None


def tables(tag="bb1e5"):
    """(thin, thick, dlogtau, NumTau_yaml) from the reference-generated fixture."""
    global _tables
    if _tables is None:
        _tables = np.load(os.path.join(GOLDEN, "ref_tables.npz"))
    thin = np.ascontiguousarray(_tables[f"{tag}_thin"])
    thick = np.ascontiguousarray(_tables[f"{tag}_thick"])
    return thin, thick, float(_tables[f"{tag}_dlogtau"]), thin.size - 1


def box_smooth(g):
    out = np.zeros_like(g)
    for a in (-1, 0, 1):
        for b in (-1, 0, 1):
            for c in (-1, 0, 1):
                out += np.roll(g, (a, b, c), axis=(0, 1, 2))
    return out / 27.0


def f1_fields(N, srcpos, seed=20240229, mean_dens=1e-3):
    """F1: log-normal density, neutral background with ionised bubbles around every 10th source."""
    rng = np.random.default_rng(seed)
    g = box_smooth(rng.normal(size=(N, N, N)))
    ndens = mean_dens * np.exp(g * 3.0 - 0.5)
    xh = np.full((N, N, N), 2e-4)
    ax = np.arange(N)
    for s in range(0, srcpos.shape[1], 10):
        c = srcpos[:, s] - 1
        r = rng.uniform(3, min(12, N / 3))
        d = [np.minimum(np.abs(ax - c[i]), N - np.abs(ax - c[i])) for i in range(3)]
        m = (d[0][:, None, None] ** 2 + d[1][None, :, None] ** 2 + d[2][None, None, :] ** 2) <= r * r
        xh[m] = 0.999
    return ndens, xh


def f0_fields(N):
    """F0: the benchmark's uniform box (raytracing_benchmark/run_test.py:32, parameters.yml:27)."""
    return np.full((N, N, N), 1e-3), np.full((N, N, N), 2e-4)


CASES = {
    # name: (N, R, numsrc, field, dr, flux mode)
    "small_r5": (16, 5.3, 1, "f1", 4e20, "one"),
    "clip_full_n24": (24, 1e3, 1, "f1", 2e20, "one"),
    "odd_n15_full": (15, 1e3, 1, "f1", 3e20, "one"),
    "r_int5": (20, 5.0, 3, "f1", 3 * MPC / 250, "ones"),
    "multi_n32": (32, 7.5, 20, "f1", 3e20, "lognormal"),
    "bench_like_n32": (32, 10.0, 5, "f0", 3 * MPC / 250, "ones"),
    "thin_n24": (24, 8.2, 4, "f1", 1e16, "lognormal"),
    "mid_n48_r14": (48, 14.0, 6, "f1", 2e20, "lognormal"),
}


def make_case(name):
    N, R, ns, field, dr, fmode = CASES[name]
    srcpos = generate_test_sources(N, ns, seed=100)  # 1-indexed, (3, ns)
    rng = np.random.default_rng(7)
    if fmode == "lognormal":
        flux = 10 ** rng.normal(0, 0.5, size=ns)
    else:
        flux = np.ones(ns)
    ndens, xh = f1_fields(N, srcpos) if field == "f1" else f0_fields(N)
    thin, thick, dlogtau, numtau = tables("bb1e5")
    pos_flat, flux_flat = format_sources(srcpos, flux)
    return dict(name=name, N=N, R=R, sig=SIG, dr=dr, ndens=ndens, xh=xh, srcpos=srcpos, flux=flux,
                pos_flat=pos_flat, flux_flat=flux_flat, thin=thin, thick=thick, minlogtau=MINLOGTAU,
                dlogtau=dlogtau, NumTau=thin.size)

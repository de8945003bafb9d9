"""CPU tests of the mirror-image sweep's plan (csrc/sweep_plan.cu: build_octant_plan) -- no GPU needed.

The plan is exported through the C ABI (asora_plan_export) and driven by a plain-Python emulation of
sweep_octant_kernel's control structure (csrc/sweep_octant.cu): octants per CTA, mirror images per thread, class A /
class B entries, the stores into both images bordering a plane, the owner-only rate deposit.  The arithmetic of a
cell is the reference's own form (raytracing.cu:285-329,397-444, rates.cu:16-83), so the result must agree with the
oracle's restatement of the ASORA kernel to rounding.  Level buffers start as NaN: reading a slot that was never
written, even with weight zero, poisons the result exactly as it would on the GPU.
"""
import ctypes
import math

import numpy as np
import pytest

import oracle
from pyc2ray_b200.lib._cabi import L, c_dp
from tests.fields import make_case, CASES

PC_RATED, PC_SOURCE, PC_DIAG2, PC_DIAG3 = 1, 2, 4, 8
SQRT3, SQRT2 = 1.73205080757, 1.41421356237
FOURPI = 12.566370614359172463991853874177


def export_plan(N, R, dr, sphere_only, octant, parts=1):
    info = (ctypes.c_int * 6)()
    n = L.asora_plan_export(N, R, dr, int(sphere_only), int(octant), parts, 0, None, None, None, None, None, None, None, None, info)
    assert n >= 0
    nlev = info[0]
    path, inv_np = np.zeros(n), np.zeros(n)
    nb = np.zeros(4 * n, dtype=np.uint16)
    d = np.zeros(3 * n, dtype=np.uint8)
    flags = np.zeros(n, dtype=np.uint8)
    ab = np.zeros(2 * n, dtype=np.uint8)
    ls = np.zeros(max(1, info[5]) * (nlev + 1), dtype=np.int32)
    lm = np.zeros(3 * nlev, dtype=np.int32)
    u8, u16, ip = ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_uint16), ctypes.POINTER(ctypes.c_int)
    n2 = L.asora_plan_export(N, R, dr, int(sphere_only), int(octant), parts, n, path.ctypes.data_as(c_dp),
                             inv_np.ctypes.data_as(c_dp), nb.ctypes.data_as(u16), d.ctypes.data_as(u8),
                             flags.ctypes.data_as(u8), ab.ctypes.data_as(u8), ls.ctypes.data_as(ip), lm.ctypes.data_as(ip), info)
    assert n2 == n
    return dict(n=n, nlevels=nlev, lmax=info[1], lo=info[2], side=info[3], q_max=info[4], path=path, inv_np=inv_np,
                nb=nb.reshape(n, 4), d=d.reshape(n, 3), flags=flags, ab=ab.reshape(n, 2), level_start=ls,
                level_mid=lm.reshape(3, nlev) if octant else None)


def lookup(table, tau, minlogtau, dlogtau, NumTau):
    """rates.cu:70-83 with the index clamped to the table (SURVEY note N6)."""
    logtau = math.log10(max(1.0e-20, tau))
    real_i = min(float(NumTau), max(0.0, 1.0 + (logtau - minlogtau) / dlogtau))
    i0 = int(real_i)
    i0 = min(i0, table.size - 1)
    i1 = min(min(NumTau, i0 + 1), table.size - 1)
    return table[i0] + (real_i - i0) * (table[i1] - table[i0])


def emulate(plan, c, noct, opt, dedup):
    """phi_ion of all sources of case `c` by the kernel's control structure."""
    N, sig, dr = c["N"], c["sig"], c["dr"]
    hi = (plan["side"] - 1) // 2
    nd, xh = c["ndens"], c["xh"]
    nhi = nd * (1.0 - xh)
    phi = np.zeros((N, N, N))
    parts = 8 // noct
    R2 = c["R"] ** 2
    row = {8: 0, 4: 1, 2: 2}.get(opt)
    lmax = plan["lmax"]
    thin, thick = c["thin"], c["thick"]
    for s in range(c["flux_flat"].size):
        i0, j0, k0 = (int(v) for v in c["pos_flat"][3 * s:3 * s + 3])
        strength = c["flux_flat"][s]
        for part in range(parts):
            bufs = np.full((2, noct, lmax), np.nan)
            bufs[:, :, lmax - 1] = 0.0   # the zero slot: zero-weight corners and the source cell
            for m in range(plan["nlevels"]):
                beg, end = plan["level_start"][m], plan["level_start"][m + 1]
                mid = plan["level_mid"][row][m] if (dedup and opt >= 2) else end
                cur, prev = bufs[m & 1], bufs[(m & 1) ^ 1]
                for e in range(beg, end):
                    fl = int(plan["flags"][e])
                    zmask = fl >> 5
                    di, dj, dk = (int(v) for v in plan["d"][e])
                    a, b = (int(v) for v in plan["ab"][e])
                    wA, wB = (a / m, b / m) if m > 0 else (0.0, 0.0)
                    s1, s2, s3, s4 = wA * wB, wB * (1 - wA), wA * (1 - wB), (1 - wA) * (1 - wB)
                    nbs = plan["nb"][e]
                    for sub in range(noct // opt):
                        obase = sub * opt
                        gbase = part * noct + obase
                        if e < mid:
                            images, zb = list(range(opt)), 0
                        else:
                            zm = zmask & (opt - 1)
                            assert zm != 0, "class B entry without a zero offset inside the thread's bits"
                            zb = 4 if zm & 4 else (2 if zm & 2 else 1)
                            images = []
                            for u in range(opt // 2):
                                low = u & (zb - 1)
                                images.append(((u - low) << 1) | low)
                        for t in images:
                            og = gbase + t
                            sx, sy, sz = (og >> 2) & 1, (og >> 1) & 1, og & 1
                            i, j, k = (i0 + (-di if sx else di)) % N, (j0 + (-dj if sy else dj)) % N, (k0 + (-dk if sz else dk)) % N
                            pv = prev[obase + t]
                            cs = [pv[nbs[0]], pv[nbs[1]], pv[nbs[2]], pv[nbs[3]]]
                            ws = [sw / max(0.6, cc * sig) for sw, cc in zip((s1, s2, s3, s4), cs)]
                            if fl & PC_SOURCE:
                                cin = 0.0
                            else:
                                cin = sum(w * cc for w, cc in zip(ws, cs)) / sum(ws)   # NaN if an unwritten slot was read
                                if fl & PC_DIAG3:
                                    cin *= SQRT3
                                elif fl & PC_DIAG2:
                                    cin *= SQRT2
                            path = plan["path"][e] * dr
                            cout = cin + nhi[i, j, k] * path
                            cur[obase + t][e - beg] = cout
                            if zb:
                                cur[obase + t + zb][e - beg] = cout
                            owned = (zmask & og) == 0
                            if (fl & PC_RATED) and owned and cin <= 2e30:
                                vol = dr ** 3 if (fl & PC_SOURCE) else FOURPI * dr ** 3 / plan["inv_np"][e]
                                tin, tout = cin * sig, cout * sig
                                pre = strength / vol
                                if abs(tout - tin) > 1e-7:
                                    rate = pre * (lookup(thick, tin, c["minlogtau"], c["dlogtau"], c["NumTau"]) -
                                                  lookup(thick, tout, c["minlogtau"], c["dlogtau"], c["NumTau"]))
                                else:
                                    rate = pre * (tout - tin) * lookup(thin, tout, c["minlogtau"], c["dlogtau"], c["NumTau"])
                                phi[i, j, k] += rate / nhi[i, j, k]
    return phi


def oracle_phi(c):
    N = c["N"]
    phi, _, _ = oracle.asora_do_all_sources(c["R"], c["sig"], c["dr"], c["ndens"].ravel(), c["xh"].ravel(), c["pos_flat"],
                                            c["flux_flat"], N, c["thin"], c["thick"], c["minlogtau"], c["dlogtau"], c["NumTau"])
    return phi.reshape(N, N, N)


def small_case(name, nsrc=None):
    c = make_case(name)
    if nsrc is not None:
        c["pos_flat"] = c["pos_flat"][:3 * nsrc]
        c["flux_flat"] = c["flux_flat"][:nsrc]
    return c


@pytest.mark.parametrize("name,nsrc", [("r_int5", 2), ("odd_n15_full", 1)])
@pytest.mark.parametrize("noct,opt,dedup", [(8, 8, True), (8, 8, False), (8, 4, True), (8, 2, True), (4, 4, True), (4, 2, True),
                                            (2, 2, True)])
def test_octant_plan_reproduces_the_oracle(name, nsrc, noct, opt, dedup):
    c = small_case(name, nsrc)
    plan = export_plan(c["N"], c["R"], c["dr"], False, True)
    phi = emulate(plan, c, noct, opt, dedup)
    ref = oracle_phi(c)
    assert np.isfinite(phi).all()
    np.testing.assert_allclose(phi, ref, rtol=1e-10, atol=1e-14 * np.abs(ref).max())
    assert np.array_equal(phi != 0, ref != 0)


def test_octant_plan_sphere_only_same_rates():
    c = small_case("r_int5", 1)
    full = emulate(export_plan(c["N"], c["R"], c["dr"], False, True), c, 8, 8, True)
    sph = emulate(export_plan(c["N"], c["R"], c["dr"], True, True), c, 8, 8, True)
    assert np.array_equal(full, sph)


def test_octant_plan_structure():
    """Counts, class boundaries and ordering at the bench shape (256^3, R = 30)."""
    plan = export_plan(256, 30.0, 3 * 3.086e24 / 256, False, True)
    assert plan["q_max"] == 52 and plan["nlevels"] == 53
    z = plan["flags"] >> 5
    nz = 3 - ((z >> 2) & 1) - ((z >> 1) & 1) - (z & 1)          # non-zero offsets of an entry
    assert int((1 << nz.astype(np.int64)).sum()) == L.asora_cells_per_source(256, 30.0) == 193025
    ls, lm = plan["level_start"], plan["level_mid"]
    for m in range(plan["nlevels"]):
        zz = z[ls[m]:ls[m + 1]]
        a8, a4, a2 = lm[0][m] - ls[m], lm[1][m] - ls[m], lm[2][m] - ls[m]
        assert (zz[:a8] == 0).all() and (zz[a8:] != 0).all()
        assert ((zz[:a4] & 3) == 0).all() and ((zz[a4:] & 3) != 0).all()
        assert ((zz[:a2] & 1) == 0).all() and ((zz[a2:] & 1) != 0).all()
    # asymmetric cell sets are refused (even mesh, octahedron beyond N/2 - 1)
    info = (ctypes.c_int * 6)()
    assert L.asora_plan_export(16, 5.3, 1.0, 0, 1, 1, 0, None, None, None, None, None, None, None, None, info) == -1


def test_whole_sweep_plan_export_counts():
    for name in ("small_r5", "r_int5", "odd_n15_full"):
        N, R = CASES[name][0], CASES[name][1]
        plan = export_plan(N, R, 1.0, False, False, 1)
        assert plan["n"] == L.asora_cells_per_source(N, R) == oracle.cells_per_source(N, R)

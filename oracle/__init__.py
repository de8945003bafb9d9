"""CPU oracle for the ASORA + chemistry hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package.  Nothing under ``pyc2ray_b200/`` imports it: the product path is
the CUDA library and raises when that library is missing.

The arithmetic lives in ``c2ray_oracle.c`` (a C restatement of the reference's Fortran CPU path and of
its CUDA path, each function citing the reference file:line it follows); this module is the ctypes
binding plus the array-layout glue the f2py / CPython wrappers of the reference provide
(``pyc2ray/evolve.py:142-155,187-194,210``).

Parity pinning: see ``tests/golden/README.md`` -- chemistry is pinned by the reference tutorial's
printed known answer, the radiation tables / source wire format by fixtures generated from the
reference's own Python modules, the ray tracer by the paper tests' printed results (Stroemgren
radius, 5-source mean ionised fractions) and, on the GPU box, by the reference CUDA kernel itself
compiled unmodified for sm_100 (``oracle/_ref``).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libc2ray_oracle.so")

FORTRAN = 0
ASORA = 1
OPT_NORMFLUX_BUG = 1
OPT_USE_SUBBOX = 2
OPT_FMA_DIST2 = 4
OPT_GREY_NOTABLES = 8   # -DGREY_NOTABLES: analytic grey-opacity rates (rates.cu:48-64, photorates.f90:13-57)

_lib = None


def build(force=False):
    """Compile the C restatement (gcc, no -march=native: the .so travels to the GPU box)."""
    src = os.path.join(_HERE, "c2ray_oracle.c")
    if (not force) and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    cmd = ["gcc", "-O3", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
           "-o", _SO, src, "-lm"]
    subprocess.check_call(cmd)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int32)
        L.oracle_do_all_sources.restype = ctypes.c_long
        L.oracle_do_all_sources.argtypes = [
            ctypes.c_int, ctypes.c_int, dp, ip, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
            ctypes.c_double, dp, dp, dp, dp, dp, dp, ctypes.c_int, ctypes.c_double, ctypes.c_double,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, dp]
        L.oracle_do_all_sources_heat.restype = ctypes.c_long
        L.oracle_do_all_sources_heat.argtypes = [
            ctypes.c_int, ctypes.c_int, dp, ip, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
            ctypes.c_double, dp, dp, dp, dp, dp, dp, dp, dp, dp, ctypes.c_int, ctypes.c_double, ctypes.c_double,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, dp]
        L.oracle_global_pass.restype = ctypes.c_int
        L.oracle_global_pass.argtypes = [ctypes.c_double, dp, dp, dp, dp, dp, dp, ctypes.c_double,
                                         ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                         ctypes.c_long, ctypes.POINTER(ctypes.c_long)]
        L.oracle_cells_per_source.restype = ctypes.c_long
        L.oracle_cells_per_source.argtypes = [ctypes.c_int, ctypes.c_double]
        L.oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def max_threads():
    return int(lib().oracle_max_threads())


def cells_per_source(N, R):
    """|octahedron(q_max) & cube| -- the per-source unit count of the updates/s metric."""
    return int(lib().oracle_cells_per_source(int(N), float(R)))


def asora_do_all_sources_heat(R, sig, dr, ndens_flat, xh_av_flat, srcpos_flat, srcflux, N, thin, thick, heat_thin,
                              heat_thick, minlogtau, dlogtau, NumTau, nthreads=1):
    """asora_do_all_sources with photo-heating rates (photorates.f90:118,124 in the ASORA kernel's conventions).
    Returns (phi_ion_flat, phi_heat_flat, n_updates)."""
    n3 = N * N * N
    nd = np.ascontiguousarray(ndens_flat, dtype=np.float64).ravel()
    xa = np.ascontiguousarray(xh_av_flat, dtype=np.float64).ravel()
    pos = np.ascontiguousarray(srcpos_flat, dtype=np.int32)
    flux = np.ascontiguousarray(srcflux, dtype=np.float64)
    tabs = [np.ascontiguousarray(t, dtype=np.float64) for t in (thin, thick, heat_thin, heat_thick)]
    phi, heat, cdh, stats = np.zeros(n3), np.zeros(n3), np.zeros(n3), np.zeros(2)
    n = lib().oracle_do_all_sources_heat(ASORA, OPT_FMA_DIST2, _dp(flux), pos.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                         flux.size, N, float(R), float(sig), float(dr), _dp(nd), _dp(xa), _dp(phi),
                                         _dp(heat), _dp(cdh), _dp(tabs[0]), _dp(tabs[1]), _dp(tabs[2]), _dp(tabs[3]),
                                         tabs[0].size, float(minlogtau), float(dlogtau), int(NumTau), 0, 0, 0.0,
                                         int(nthreads), _dp(stats))
    return phi, heat, int(n)


def asora_do_all_sources(R, sig, dr, ndens_flat, xh_av_flat, srcpos_flat, srcflux, N, thin, thick,
                         minlogtau, dlogtau, NumTau, fma_dist2=True, nthreads=1, grey_notables=False):
    """CPU restatement of libasora.do_all_sources (src/asora/raytracing.cu:79-148).

    Inputs use the ASORA wire format: flat C-ordered float64 grids, int32 interleaved 0-indexed
    source positions (sourceutils.py:30).  Returns (phi_ion_flat, coldensh_out_flat, n_updates);
    coldensh_out is only meaningful for a single source and nthreads == 1.
    """
    n3 = N * N * N
    ndens_flat = np.ascontiguousarray(ndens_flat, dtype=np.float64).ravel()
    xh_av_flat = np.ascontiguousarray(xh_av_flat, dtype=np.float64).ravel()
    assert ndens_flat.size == n3 and xh_av_flat.size == n3
    pos = np.ascontiguousarray(srcpos_flat, dtype=np.int32)
    flux = np.ascontiguousarray(srcflux, dtype=np.float64)
    thin = np.ascontiguousarray(thin, dtype=np.float64)
    thick = np.ascontiguousarray(thick, dtype=np.float64)
    phi = np.zeros(n3)
    cdh = np.zeros(n3)
    stats = np.zeros(2)
    opts = (OPT_FMA_DIST2 if fma_dist2 else 0) | (OPT_GREY_NOTABLES if grey_notables else 0)
    n = lib().oracle_do_all_sources(ASORA, opts, _dp(flux), pos.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                    flux.size, N, float(R), float(sig), float(dr), _dp(ndens_flat),
                                    _dp(xh_av_flat), _dp(phi), _dp(cdh), _dp(thin), _dp(thick), thin.size,
                                    float(minlogtau), float(dlogtau), int(NumTau), 0, 0, 0.0, int(nthreads), _dp(stats))
    return phi, cdh, int(n)


def fortran_do_all_sources(normflux, srcpos, max_subbox, subboxsize, sig, dr, ndens, xh_av, loss_fraction,
                           thin, thick, minlogtau, dlogtau, R_max_LLS, NumTau=None, use_subbox=True,
                           normflux_bug=False, nthreads=1, heat_thin=None, heat_thick=None, grey_notables=False):
    """CPU restatement of libc2ray.raytracing.do_all_sources (src/c2ray/raytracing.f90:52-119).

    srcpos is (3, NumSrc), 1-indexed; ndens / xh_av are (N,N,N) logical arrays (any memory order; they
    are handed to the C code Fortran-ordered exactly as f2py would).  Returns
    (phi_ion (N,N,N) F-ordered, coldensh_out (N,N,N) F-ordered, nsubbox, photon_loss, n_updates).
    """
    N = ndens.shape[0]
    nd = np.asfortranarray(ndens, dtype=np.float64)
    xa = np.asfortranarray(xh_av, dtype=np.float64)
    pos = np.ascontiguousarray(np.asarray(srcpos).T, dtype=np.int32).ravel()  # srcpos(3,NumSrc) column-major
    flux = np.ascontiguousarray(normflux, dtype=np.float64)
    thin = np.ascontiguousarray(thin, dtype=np.float64)
    thick = np.ascontiguousarray(thick, dtype=np.float64)
    if NumTau is None:
        NumTau = thin.size
    phi = np.zeros((N, N, N), order="F")
    cdh = np.zeros((N, N, N), order="F")
    stats = np.zeros(2)
    opts = ((OPT_USE_SUBBOX if use_subbox else 0) | (OPT_NORMFLUX_BUG if normflux_bug else 0) |
            (OPT_GREY_NOTABLES if grey_notables else 0))
    if heat_thin is not None:
        # with heating tables: returns phi_heat as a sixth value (raytracing.f90:52-110 fills it in place)
        ht = np.ascontiguousarray(heat_thin, dtype=np.float64)
        hk = np.ascontiguousarray(heat_thick, dtype=np.float64)
        heat = np.zeros((N, N, N), order="F")
        n = lib().oracle_do_all_sources_heat(FORTRAN, opts, _dp(flux), pos.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                             flux.size, N, float(R_max_LLS), float(sig), float(dr), _dp(nd), _dp(xa),
                                             _dp(phi), _dp(heat), _dp(cdh), _dp(thin), _dp(thick), _dp(ht), _dp(hk),
                                             thin.size, float(minlogtau), float(dlogtau), int(NumTau), int(max_subbox),
                                             int(subboxsize), float(loss_fraction), int(nthreads), _dp(stats))
        return phi, cdh, int(stats[0]), float(stats[1]), int(n), heat
    n = lib().oracle_do_all_sources(FORTRAN, opts, _dp(flux), pos.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                    flux.size, N, float(R_max_LLS), float(sig), float(dr), _dp(nd), _dp(xa),
                                    _dp(phi), _dp(cdh), _dp(thin), _dp(thick), thin.size, float(minlogtau),
                                    float(dlogtau), int(NumTau), int(max_subbox), int(subboxsize),
                                    float(loss_fraction), int(nthreads), _dp(stats))
    return phi, cdh, int(stats[0]), float(stats[1]), int(n)


def global_pass(dt, ndens, temp, xh, xh_av, xh_intermed, phi_ion, bh00, albpow, colh0, temph0, abu_c):
    """CPU restatement of libc2ray.chemistry.global_pass (src/c2ray/chemistry.f90:13-48).

    xh_av and xh_intermed must be float64 arrays sharing one memory order with the other grids; they
    are updated in place (intent(inout)).  Returns conv_flag.
    """
    arrs = [ndens, temp, xh, xh_av, xh_intermed, phi_ion]
    for a in arrs:
        assert a.dtype == np.float64
    order = "F" if all(a.flags.f_contiguous for a in arrs) else "C"
    if order == "C":
        assert all(a.flags.c_contiguous for a in arrs), "grids must share one contiguous memory order"
    n = ndens.size
    nit = ctypes.c_long(0)
    flag = lib().oracle_global_pass(float(dt), _dp(ndens), _dp(temp), _dp(xh), _dp(xh_av), _dp(xh_intermed),
                                    _dp(phi_ion), float(bh00), float(albpow), float(colh0), float(temph0),
                                    float(abu_c), n, ctypes.byref(nit))
    return int(flag)

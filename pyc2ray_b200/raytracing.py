"""Stand-alone ray tracing (reference: pyc2ray/raytracing.py:34-108)."""
import time

import numpy as np

from .asora_core import cuda_is_init
from .load_extensions import load_asora
from .utils import printlog
from .utils.sourceutils import format_sources

libasora = load_asora()

__all__ = ["do_raytracing"]


def do_raytracing(dr, src_flux, src_pos, use_gpu, max_subbox, subboxsize, loss_fraction, ndens, xh_av,
                  photo_thin_table, photo_thick_table, heat_thin_table, heat_thick_table, minlogtau, dlogtau,
                  R_max_LLS, sig, logfile="pyC2Ray.log", quiet=False, stats=False):
    """Photo-ionisation rate of every cell for the current ionised fractions (no chemistry).

    Same arguments as the reference.  Only ``use_gpu=True`` exists in this build; the CPU-only
    arguments (max_subbox, subboxsize, loss_fraction) are accepted and unused.
    Returns (phi_ion, phi_heat).  The reference's GPU branch has no heating (it returns an undefined name
    there: raytracing.py:106-108; TODO at c2ray_base.py:424-426); here non-zero heating tables switch the
    photo-heating rates on (photorates.f90:118,124 evaluated in the sweep kernel), all-zero or missing tables
    -- what C2Ray passes when ``compute_heating_rates`` is off, c2ray_base.py:431-433 -- give phi_heat = None.
    """
    if not use_gpu:
        raise NotImplementedError("CPU ray tracing is not part of this build (use_gpu must be True)")
    if not cuda_is_init():
        raise RuntimeError("GPU not initialized. Please initialize it by calling device_init(N)")
    NumSrc = src_flux.shape[0]
    N = ndens.shape[0]
    NumTau = photo_thin_table.shape[0]

    xh_av_flat = np.ravel(xh_av).astype("float64", copy=True)
    ndens_flat = np.ravel(ndens).astype("float64", copy=True)
    srcpos_flat, normflux_flat = format_sources(src_pos, src_flux)
    libasora.source_data_to_device(srcpos_flat, normflux_flat, NumSrc)
    coldensh_out_flat = np.zeros(1, dtype="float64")  # ignored by the library (raytracing.cu:116)
    phi_ion_flat = np.zeros(N * N * N, dtype="float64")
    libasora.density_to_device(ndens_flat, N)
    printlog("Copied source data to device.", logfile, quiet)
    printlog(f"dr [Mpc]: {dr/3.086e24:.3e}", logfile, quiet)
    printlog(f"Running on {NumSrc:n} source(s), total normalized ionizing flux: {src_flux.sum():.2e}", logfile, quiet)
    printlog(f"Mean density (cgs): {ndens.mean():.3e}, Mean ionized fraction: {xh_av.mean():.3e}", logfile, quiet)

    trt0 = time.time()
    printlog("Doing Raytracing...", logfile, quiet, " ")
    heating = (heat_thin_table is not None and heat_thick_table is not None
               and (np.any(heat_thin_table) or np.any(heat_thick_table)))
    if heating:
        libasora.heat_table_to_device(np.ascontiguousarray(heat_thin_table, dtype=np.float64),
                                      np.ascontiguousarray(heat_thick_table, dtype=np.float64), NumTau)
        phi_heat_flat = np.zeros(N * N * N, dtype="float64")
        libasora.do_all_sources_heat(R_max_LLS, coldensh_out_flat, sig, dr, ndens_flat, xh_av_flat, phi_ion_flat,
                                     phi_heat_flat, NumSrc, N, minlogtau, dlogtau, NumTau)
    else:
        libasora.do_all_sources(R_max_LLS, coldensh_out_flat, sig, dr, ndens_flat, xh_av_flat, phi_ion_flat, NumSrc, N,
                                minlogtau, dlogtau, NumTau)
    printlog(f"took {(time.time()-trt0) : .1f} s.", logfile, quiet)
    return np.reshape(phi_ion_flat, (N, N, N)), (np.reshape(phi_heat_flat, (N, N, N)) if heating else None)

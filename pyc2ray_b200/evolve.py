"""The C2Ray time-step: iterate ray tracing and chemistry until the ionised fractions converge.

Reference: pyc2ray/evolve.py -- evolve3D :38-245, evolve3D_MPI :249-498.  Same signatures, same
convergence logic; what changes is where the data lives.  The reference crosses PCIe three times per
iteration (H2D xh_av, D2H phi_ion, then the Fortran chemistry on the host plus two transposing
copies).  Here ndens, temp, xh, xh_av, xh_intermed and phi_ion stay in HBM for the whole convergence
loop; per iteration the host sees three scalars (conv_flag, sum x, sum 1-x).
"""
import ctypes
import os
import time

import numpy as np

from .asora_core import cuda_is_init
from .lib import _cabi
from .lib._cabi import L, check, dptr, iptr
from .parallel import (shard_bounds, allreduce_sum_, reduce_scatter_sum_, allgather_chunks_, device_tensor, slab_edges,
                       SlabHalo)
from .utils import printlog
from .utils.sourceutils import format_sources

__all__ = ["evolve3D", "evolve3D_MPI", "evolve3D_dist"]


def _flat(a):
    """float64, flat, logical C order (index i*N*N + j*N + k), as evolve.py:142-143; a view when the input
    already has that layout (the library only reads it)."""
    return np.ascontiguousarray(np.ravel(a), dtype=np.float64)


def _is_f(a):
    """Fortran-ordered float64 3-D grid (what the reference's driver classes hold)?"""
    return (isinstance(a, np.ndarray) and a.ndim == 3 and a.dtype == np.float64 and a.flags.f_contiguous
            and not a.flags.c_contiguous)


def _upload(buf, a):
    """Whole-grid upload; Fortran-ordered grids go up as they are and are re-ordered on the device."""
    if _is_f(a):
        check(L.asora_buffer_upload_f(buf, dptr(a)))
    else:
        check(L.asora_buffer_upload(buf, dptr(_flat(a))))


def _evolve_device(dt, dr, src_flux, src_pos, temp, ndens, xh, photo_thin_table, minlogtau, dlogtau, R_max_LLS,
                   convergence_fraction, sig, bh00, albpow, colh0, temph0, abu_c, logfile, quiet, shard=None,
                   group=None, max_iter=10000, decomposition="list", io_rank=None):
    if not cuda_is_init():
        raise RuntimeError("GPU not initialized. Please initialize it by calling device_init(N)")
    NumSrc_total = src_flux.shape[0]
    N = temp.shape[0]
    NumCells = N * N * N
    NumTau = photo_thin_table.shape[0]
    conv_criterion = min(int(convergence_fraction * NumCells), (NumSrc_total - 1) / 3)  # evolve.py:127
    prev_sum_xh1_int = 2 * NumCells
    prev_sum_xh0_int = 2 * NumCells

    halo = None
    if shard is not None:
        rank, nprocs = shard
        edges = None
        if decomposition in ("auto", "slab"):
            # shard by position: planes of the slowest axis with equal source counts (parallel.py)
            edges, h = slab_edges(np.asarray(src_pos)[0] - 1, N, nprocs, R_max_LLS)
            if edges is None and decomposition == "slab":
                raise ValueError("slab decomposition impossible: a halo does not fit the neighbouring slab")
        if edges is not None:
            # halo planes read straight from the neighbours' GPUs over NVLink (CUDA IPC); ASORA_PEER_HALO=0: NCCL send/recv
            halo = SlabHalo(edges, h, N, rank, nprocs, group, peer=os.environ.get("ASORA_PEER_HALO", "1") != "0")
            x0 = np.mod(np.asarray(src_pos)[0].astype(np.int64) - 1, N)
            mine = (x0 >= halo.lo) & (x0 < halo.hi)
            srcpos_flat, normflux_flat = format_sources(np.asarray(src_pos)[:, mine], np.asarray(src_flux)[mine])
        else:
            i_start, i_end = shard_bounds(NumSrc_total, rank, nprocs)   # list order: evolve.py:362-367
            srcpos_flat, normflux_flat = format_sources(src_pos[:, i_start:i_end], src_flux[i_start:i_end])
    else:
        rank, nprocs = 0, 1
        srcpos_flat, normflux_flat = format_sources(src_pos, src_flux)
    NumSrc = normflux_flat.shape[0]
    # List-order sharding, optimised exchange: reduce-scatter of phi_ion, chemistry on the rank's own
    # N^3/nprocs cells, all-gather of the new xh_av -- the bytes of one all-reduce, 1/nprocs of the chemistry.
    rsag = (nprocs > 1 and halo is None and decomposition in ("auto", "rsag") and NumCells % nprocs == 0)
    if decomposition == "rsag" and nprocs > 1 and not rsag:
        raise ValueError("rsag decomposition needs N^3 divisible by the number of ranks")

    check(L.asora_source_data_to_device(iptr(srcpos_flat), dptr(normflux_flat), NumSrc))
    if nprocs > 1 and io_rank is not None:
        # one rank reads the grids from its host; the others receive them from that GPU over NVLink (three broadcasts of
        # 8 N^3 bytes) instead of nprocs uploads through the host's memory system
        import torch
        import torch.distributed as dist
        for buf, arr in ((_cabi.BUF_NDENS, ndens), (_cabi.BUF_TEMP, temp), (_cabi.BUF_XH, xh)):
            if rank == io_rank:
                _upload(buf, arr)
            t = device_tensor(L.asora_device_buffer(buf), NumCells)
            dist.broadcast(t, src=dist.get_global_rank(group, io_rank) if group is not None else io_rank, group=group)
        torch.cuda.synchronize()
        check(L.asora_invalidate_temperature())
    elif halo is None or _is_f(ndens) or _is_f(temp) or _is_f(xh):
        _upload(_cabi.BUF_NDENS, ndens)
        _upload(_cabi.BUF_TEMP, temp)
        _upload(_cabi.BUF_XH, xh)
    else:
        # a rank only ever reads its own planes and the halos: upload just those (at most two segments)
        first, count = halo.active_range()
        segs = [(first, min(count, N - first)), (0, count - min(count, N - first))]
        for buf, arr in ((_cabi.BUF_NDENS, _flat(ndens)), (_cabi.BUF_TEMP, _flat(temp)), (_cabi.BUF_XH, _flat(xh))):
            for b, c in segs:
                if c > 0:
                    check(L.asora_buffer_upload_range(buf, dptr(arr), b * N * N, c * N * N))
    for b in (_cabi.BUF_XH_AV, _cabi.BUF_XH_INTERMED):  # xh_av = xh_intermed = copy(xh): evolve.py:136-137
        check(L.asora_buffer_copy(b, _cabi.BUF_XH))
    phi_t = xav_t = None
    if nprocs > 1:
        import torch
        phi_t = device_tensor(L.asora_device_buffer(_cabi.BUF_PHI_ION), NumCells)
        if rsag:
            xav_t = device_tensor(L.asora_device_buffer(_cabi.BUF_XH_AV), NumCells)
            scal = torch.zeros(3, dtype=torch.float64, device="cuda")
            chunk = NumCells // nprocs
        if halo is not None:
            xav_t = device_tensor(L.asora_device_buffer(_cabi.BUF_XH_AV), NumCells)
            scal = torch.zeros(3, dtype=torch.float64, device="cuda")

    if rank == 0 and not (quiet and logfile is None):  # (the two grid means below cost 22 ms at 250^3)
        printlog("Calling evolve3D..." if nprocs == 1 else f"Calling evolve3D with {nprocs:n} ranks...", logfile, quiet)
        printlog(f"dr [Mpc]: {dr/3.086e24:.3e}", logfile, quiet)
        printlog(f"dt [years]: {dt/3.15576E+07:.3e}", logfile, quiet)
        printlog(f"Running on {NumSrc_total:n} source(s), total normalized ionizing flux: {src_flux.sum():.2e}", logfile, quiet)
        printlog(f"Mean density (cgs): {ndens.mean():.3e}, Mean ionized fraction: {xh.mean():.3e}", logfile, quiet)
        printlog(f"Convergence Criterion (Number of points): {conv_criterion : n}", logfile, quiet, end="\n\n")

    converged = False
    niter = 0
    flag = ctypes.c_int(0)
    s1 = ctypes.c_double(0.0)
    s0 = ctypes.c_double(0.0)
    t_loop0 = time.perf_counter()
    phases = {"sweep": 0.0, "exchange_phi": 0.0, "chemistry": 0.0, "exchange_xh": 0.0}  # seconds, summed over the iterations
    try:
        # process-global sweep settings: switched on inside the try so that the finally below always resets them
        # The evolve loop only consumes phi_ion, so the sweep may skip the cells outside the R_max sphere
        # (identical rates, see asora_set_sphere_only in include/asora_b200.h).
        check(L.asora_set_sphere_only(1))
        if halo is not None:
            check(L.asora_set_active_slab(*halo.active_range()))
        while not converged:
            niter += 1
            trt0 = time.time()
            trt0p = time.perf_counter()
            check(L.asora_raytrace_device(float(R_max_LLS), float(sig), float(dr), 0, NumSrc, float(minlogtau),
                                          float(dlogtau), int(NumTau), 1))
            check(L.asora_sync())
            tph = time.perf_counter()
            phases["sweep"] += tph - trt0p
            if nprocs > 1:
                torch.cuda.nvtx.range_push("asora:exchange_phi")
            if halo is not None:
                halo.reduce_phi_(phi_t)       # neighbours' rates for my planes: 2 halos of h*N^2 doubles
                torch.cuda.synchronize()
            elif rsag:
                reduce_scatter_sum_(phi_t, rank, nprocs, group)  # my chunk of the summed rates
                torch.cuda.synchronize()
            elif nprocs > 1:
                allreduce_sum_(phi_t, group)  # evolve.py:433-437 (Reduce + Bcast) as one NCCL all-reduce
                torch.cuda.synchronize()
            if nprocs > 1:
                torch.cuda.nvtx.range_pop()
            trt = time.time() - trt0
            phases["exchange_phi"] += time.perf_counter() - tph
            tch0 = time.time()
            tch0p = time.perf_counter()
            if halo is not None:
                o, cnt = halo.own_cells()
                check(L.asora_global_pass_device_range(float(dt), float(bh00), float(albpow), float(colh0), float(temph0),
                                                       float(abu_c), o, cnt, ctypes.byref(flag), ctypes.byref(s1),
                                                       ctypes.byref(s0)))
                phases["chemistry"] += time.perf_counter() - tch0p
                tch0p = time.perf_counter()
                scal.copy_(torch.tensor([flag.value, s1.value, s0.value], dtype=torch.float64))
                allreduce_sum_(scal, group)   # conv_flag, sum x, sum 1-x over all planes
                halo.gather_xh_(xav_t, synced=True)   # my neighbours' new xh_av inside my ray-tracing reach
                torch.cuda.synchronize()
                g = scal.tolist()
                conv_flag, sum_xh1_int, sum_xh0_int = int(round(g[0])), g[1], g[2]
                phases["exchange_xh"] += time.perf_counter() - tch0p
                tch0p = None
            elif rsag:
                check(L.asora_global_pass_device_range(float(dt), float(bh00), float(albpow), float(colh0), float(temph0),
                                                       float(abu_c), rank * chunk, chunk, ctypes.byref(flag), ctypes.byref(s1),
                                                       ctypes.byref(s0)))
                scal.copy_(torch.tensor([flag.value, s1.value, s0.value], dtype=torch.float64))
                allreduce_sum_(scal, group)                     # conv_flag, sum x, sum 1-x over all chunks
                allgather_chunks_(xav_t, rank, nprocs, group)   # every rank sweeps with the whole new xh_av
                torch.cuda.synchronize()
                g = scal.tolist()
                conv_flag, sum_xh1_int, sum_xh0_int = int(round(g[0])), g[1], g[2]
            else:
                check(L.asora_global_pass_device(float(dt), float(bh00), float(albpow), float(colh0), float(temph0),
                                                 float(abu_c), ctypes.byref(flag), ctypes.byref(s1), ctypes.byref(s0)))
                conv_flag, sum_xh1_int, sum_xh0_int = flag.value, s1.value, s0.value
            tch = time.time() - tch0
            if tch0p is not None:
                phases["chemistry"] += time.perf_counter() - tch0p
            # evolve.py:216-232
            rel_change_xh1 = abs((sum_xh1_int - prev_sum_xh1_int) / sum_xh1_int) if sum_xh1_int > 0.0 else 1.0
            rel_change_xh0 = abs((sum_xh0_int - prev_sum_xh0_int) / sum_xh0_int) if sum_xh0_int > 0.0 else 1.0
            if rank == 0:
                printlog(f"Raytracing took {trt*1e3:.2f} ms, chemistry {tch*1e3:.2f} ms. Number of non-converged points: "
                         f"{conv_flag} of {NumCells} ({conv_flag / NumCells * 100 : .3f} % ), Relative change in "
                         f"ionfrac: {rel_change_xh1 : .2e}", logfile, quiet)
            converged = (conv_flag < conv_criterion) or ((rel_change_xh1 < convergence_fraction) and
                                                         (rel_change_xh0 < convergence_fraction))
            prev_sum_xh1_int = sum_xh1_int
            prev_sum_xh0_int = sum_xh0_int
            if niter >= max_iter:
                raise RuntimeError("evolve3D: no convergence")
    finally:
        L.asora_set_sphere_only(0)
        L.asora_set_active_slab(0, 0)
        if halo is not None:
            halo.close()   # unmap the neighbours' buffers (no collective: safe on the error path too)
    evolve3D.last_loop_seconds = time.perf_counter() - t_loop0  # the convergence loop alone: no host<->device grid copies
    evolve3D.last_phase_seconds = phases
    if rsag:
        # once per time step: every rank gets the whole grids back (the reference API returns full arrays)
        allgather_chunks_(device_tensor(L.asora_device_buffer(_cabi.BUF_XH_INTERMED), NumCells), rank, nprocs, group)
        allgather_chunks_(phi_t, rank, nprocs, group)
        torch.cuda.synchronize()
    if halo is not None:
        # once per time step: every rank gets the whole grids back (the reference API returns full arrays)
        halo.assemble_(device_tensor(L.asora_device_buffer(_cabi.BUF_XH_INTERMED), NumCells))
        halo.assemble_(phi_t)
        torch.cuda.synchronize()
    if rank == 0:
        printlog("Multiple source convergence reached.", logfile, quiet)
    evolve3D.last_niter = niter
    if nprocs > 1 and io_rank is not None and rank != io_rank:
        return None, None  # only io_rank wanted the grids on its host
    phi_ion = np.empty(NumCells)
    check(L.asora_buffer_download(_cabi.BUF_PHI_ION, dptr(phi_ion)))
    if _is_f(xh):
        # np.copy(xh) keeps the order of xh (evolve.py:136-137,244): hand back a Fortran-ordered grid,
        # re-ordered on the device
        xh_new = np.empty((N, N, N), order="F")
        check(L.asora_buffer_download_f(_cabi.BUF_XH_INTERMED, dptr(xh_new)))
    else:
        xh_new = np.empty(NumCells)
        check(L.asora_buffer_download(_cabi.BUF_XH_INTERMED, dptr(xh_new)))
        xh_new = xh_new.reshape(N, N, N)
    evolve3D.last_niter = niter
    return xh_new, phi_ion.reshape(N, N, N)


def evolve3D(dt, dr, src_flux, src_pos, use_gpu, max_subbox, subboxsize, loss_fraction, temp, ndens, xh,
             photo_thin_table, photo_thick_table, minlogtau, dlogtau, R_max_LLS, convergence_fraction, sig, bh00,
             albpow, colh0, temph0, abu_c, logfile="pyC2Ray.log", quiet=False):
    """Evolve the ionised fraction of the whole grid over one time step (evolve.py:38-245).

    Arguments and return values as in the reference (xh_new, phi_ion).  ``use_gpu`` must be True;
    the tables must have been copied with photo_table_to_device().  max_subbox, subboxsize and
    loss_fraction only affect the reference's CPU ray tracer and are ignored.
    """
    if not use_gpu:
        raise NotImplementedError("CPU ray tracing is not part of this build (use_gpu must be True)")
    return _evolve_device(dt, dr, src_flux, src_pos, temp, ndens, xh, photo_thin_table, minlogtau, dlogtau,
                          R_max_LLS, convergence_fraction, sig, bh00, albpow, colh0, temph0, abu_c, logfile, quiet)


evolve3D.last_niter = 0
evolve3D.last_loop_seconds = 0.0


def evolve3D_dist(dt, dr, src_flux, src_pos, temp, ndens, xh, photo_thin_table, photo_thick_table, minlogtau,
                  dlogtau, R_max_LLS, convergence_fraction, sig, bh00, albpow, colh0, temph0, abu_c,
                  logfile="pyC2Ray.log", quiet=False, group=None, decomposition="auto", io_rank=None):
    """Source-sharded time step over the ranks of an initialised torch.distributed process group
    (backend nccl, one rank per GPU).  Every rank passes the full source list and gets the full
    result.

    decomposition: "list" -- contiguous blocks of the source list, one all-reduce of phi_ion per iteration and the
    chemistry of the whole grid on every rank (the reference's scheme, evolve.py:360-373,433-437, minus its rank-0
    chemistry and broadcasts); "rsag" -- the same sharding with a reduce-scatter of phi_ion, chemistry on the rank's
    N^3/nprocs cells and an all-gather of xh_av; "slab" -- sources sharded by position, halo exchanges
    instead of N^3 collectives (parallel.SlabHalo); "auto" -- slab when the slabs are wide enough for the ray-tracing
    radius, else rsag when N^3 divides by the number of ranks, else list.

    io_rank: None -- every rank passes the full grids and gets the full result (the reference's semantics); r -- only rank
    r's ``temp``, ``ndens``, ``xh`` are read (the other ranks receive them GPU to GPU over NVLink) and only rank r gets the
    result (the others return ``(None, None)``): one set of host<->device copies per time step instead of one per rank."""
    import torch.distributed as dist
    rank, nprocs = dist.get_rank(group), dist.get_world_size(group)
    shard = (rank, nprocs) if src_flux.shape[0] >= nprocs else None  # c2ray_base.py:185
    return _evolve_device(dt, dr, src_flux, src_pos, temp, ndens, xh, photo_thin_table, minlogtau, dlogtau,
                          R_max_LLS, convergence_fraction, sig, bh00, albpow, colh0, temph0, abu_c, logfile,
                          quiet or rank != 0, shard=shard, group=group, decomposition=decomposition, io_rank=io_rank)


def evolve3D_MPI(dt, dr, src_flux, src_pos, use_gpu, max_subbox, subboxsize, loss_fraction, use_mpi, comm, rank,
                 nprocs, temp, ndens, xh, photo_thin_table, photo_thick_table, minlogtau, dlogtau, R_max_LLS,
                 convergence_fraction, sig, bh00, albpow, colh0, temph0, abu_c, logfile="pyC2Ray.log", quiet=False):
    """Signature of the reference's MPI variant (evolve.py:249-258).  The mpi4py arguments are
    accepted for compatibility; the exchange runs over the default torch.distributed group (NCCL),
    which must have ``nprocs`` ranks with this process as ``rank``."""
    import torch.distributed as dist
    if not use_gpu:
        raise NotImplementedError("CPU ray tracing is not part of this build (use_gpu must be True)")
    if not dist.is_initialized() or dist.get_world_size() != nprocs or dist.get_rank() != rank:
        raise RuntimeError("evolve3D_MPI: initialise torch.distributed (backend='nccl') with the same rank/size")
    return evolve3D_dist(dt, dr, src_flux, src_pos, temp, ndens, xh, photo_thin_table, photo_thick_table, minlogtau,
                         dlogtau, R_max_LLS, convergence_fraction, sig, bh00, albpow, colh0, temph0, abu_c, logfile, quiet)
evolve3D.last_phase_seconds = {}

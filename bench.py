#!/usr/bin/env python
"""bench.py -- ASORA source-cell updates/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--R 30] [--nsrc 10000] [--mesh 256]
    python bench.py --sweep [--mesh 250]        radius x source-count table of the reference's raytracing_benchmark

Workload (config.workload): the reference's ray-tracing benchmark
(test/paper_tests/raytracing_benchmark/run_test.py) at the BASELINE size: 256^3 mesh, uniform
ndens = 1e-3, xh = 2e-4, black-body Teff = 1e5 tables with NumTau = 20000, 10^4 seeded random sources
per GPU, ray-tracing radius R = 30 cells (the radius of the reference's published 3.156 ns asymptote).
A step = one do_all_sources pass over the rank's 10^4 sources; with N > 1 ranks every rank sweeps its
own 10^4 sources (weak scaling) and the rate grids are summed by one NCCL all-reduce inside the step.

One JSON line is printed by rank 0:
  value            source-cell updates/s, all ranks, inputs resident in HBM (CUDA events, max over ranks)
  e2e              same metric through the reference-facing call libasora.do_all_sources with pageable HOST (numpy)
                   buffers, as a drop-in caller passes them (H2D of xh_av and D2H of phi_ion inside the timed region);
                   e2e_pinned: the same with page-locked buffers
  roofline         sweep kernel: 32 algorithmic bytes per update / mean launch duration (CUDA events on the
                   launching stream) vs the measured HBM copy bandwidth of MEASURED_PEAKS.json
  parity           phi_ion of the first 8 sources of THIS run's inputs, default launch shape, per cell against the
                   reference's own CUDA kernel (oracle/_ref, compiled unmodified for sm_100) or, without it, the C
                   oracle; the run exits non-zero above 1e-9
  reference_gpu    the reference's own kernel on the same GPU and workload (batch 64 and 128; N = 1 only)
  cpu_baseline     the CPU port of the reference's Fortran ray tracer (oracle/, Fortran flavour with the
                   benchmark's sub-box settings) on all host cores, on a bounded sample of the same sources
  N > 1 adds       allreduce_check (all-reduced phi_ion against the per-rank partials summed on rank 0),
                   strong (the fixed 10^4-source problem and a 512^3 / 10^5-source evolve step split over the ranks),
                   multi_gpu_parity (evolve3D_dist list / rsag / slab against single-GPU evolve3D)
`--impl reference` times only the CPU port (the Fortran itself cannot be built: no Fortran compiler) and imports
nothing of pyc2ray_b200's CUDA library.
"""
import argparse
import ctypes
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("ASORA_QUIET", "1")  # keep stdout to the one JSON line

MPC = 3.086e24
SIG = 6.30e-18
BYTES_PER_UPDATE = 32  # read ndens 8 + read xh_av 8 + read-modify-write phi_ion 16
PARITY_TOL = 1e-9
CHEM = (2.59e-13, -0.7, 1.3e-8 * 0.83 / 13.598 ** 2, 13.598 / 8.617e-5, 7.1e-7)


def _load(name, *path):
    """A pure-Python module of the package by file path: importing the package itself loads the CUDA library, which
    the reference arm must not do."""
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "pyc2ray_b200", *path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def workload(N, nsrc, seed):
    su = _load("_bench_sourceutils", "utils", "sourceutils.py")
    srcpos = su.generate_test_sources(N, nsrc, seed=seed)
    flux = 10 ** np.random.default_rng(seed).normal(0, 0.5, size=nsrc)
    ndens = np.full(N ** 3, 1e-3)
    xh = np.full(N ** 3, 2e-4)
    pos_flat, flux_flat = su.format_sources(srcpos, flux)
    return srcpos, flux, pos_flat, flux_flat, ndens, xh


_TABLES = None


def tables():
    global _TABLES
    if _TABLES is None:
        rad = _load("_bench_radiation", "radiation.py")
        thin, thick, dlogtau = rad.blackbody_tables(1e5, False, -20.0, 4.0, 20000)
        _TABLES = (thin, thick, dlogtau, 20000)
    return _TABLES


class quiet_stdout:
    """The reference's device_init / device_close print to stdout (memory.cu:52-59,77-78); bench.py must print one line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *exc):
        ctypes.CDLL(None).fflush(None)
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)


def host_threads():
    """Host cores this process may use -- not OMP_NUM_THREADS, which torchrun sets to 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def bind_to_gpu_numa(local_rank):
    """Best effort: restrict this process (and the staging threads of the library's host copies) to the cores of the
    NUMA node its GPU hangs off, so that 8 ranks do not stage 8 x 268 MB through one socket.  Returns a description."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return f"gpu {bus}: no NUMA node reported"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"gpu {bus}: NUMA node {node} has no permitted cores"
        os.sched_setaffinity(0, cpus)
        return f"gpu {bus}: NUMA node {node}, {len(cpus)} cores"
    except Exception as e:  # no sysfs entry, no permission, ...
        return f"not bound ({type(e).__name__})"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm
# ---------------------------------------------------------------------------------------------------------------------

def cpu_reference_rate(R, nsrc_sample, threads, N, seed=100):
    """Time the CPU port of src/c2ray/raytracing.f90 (oracle, Fortran flavour) exactly as the reference's
    benchmark drives it: sub-boxes of size R, max_subbox 1000, loss_fraction 1e-2
    (raytracing_benchmark/run_test.py:38,88; parameters.yml).  Returns a dict: rate in the metric's unit (octahedron
    cells of the GPU path per source), the cells the CPU really swept (its sub-box cube is larger), seconds."""
    import oracle
    srcpos, flux, _, _, ndens, xh = workload(N, nsrc_sample, seed)
    thin, thick, dlogtau, numtau = tables()
    nd3 = ndens.reshape(N, N, N)
    xh3 = xh.reshape(N, N, N)
    t0 = time.perf_counter()
    out = oracle.fortran_do_all_sources(flux, srcpos, 1000, max(1, int(R)), SIG, 3 * MPC / N, nd3, xh3, 1e-2, thin, thick,
                                        -20.0, dlogtau, R, NumTau=numtau, use_subbox=True, nthreads=threads)
    dt = time.perf_counter() - t0
    units = nsrc_sample * oracle.cells_per_source(N, R)
    return {"rate": units / dt, "rate_cells_swept": out[4] / dt, "cells_swept_per_source": out[4] / nsrc_sample,
            "seconds": dt, "threads": threads, "sources": nsrc_sample}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    N = args.mesh
    threads = host_threads()
    # bounded: a calibration step sizes the sample so that warm-up + K steps stay below ~60 s of wall clock
    cal = cpu_reference_rate(args.R, 4 * threads, threads, N)
    per_source = cal["seconds"] / cal["sources"]
    budget = 50.0 / max(1, args.steps + (1 if args.warmup > 0 else 0))
    sample = int(max(threads, min(args.cpu_sample, budget / per_source)))
    if args.warmup > 0:
        cpu_reference_rate(args.R, sample, threads, N)
    runs = [cpu_reference_rate(args.R, sample, threads, N) for _ in range(args.steps)]
    secs = sum(r["seconds"] for r in runs)
    cells = oracle.cells_per_source(N, args.R)
    value = args.steps * sample * cells / secs
    swept = runs[-1]["cells_swept_per_source"]
    line = {
        "impl": "reference", "metric": "ASORA source-cell updates/s", "value": value, "unit": "updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"raytracing_benchmark {N}^3 uniform ndens=1e-3 xh=2e-4 R={args.R:g}, CPU port of "
                               f"src/c2ray/raytracing.f90 (subboxsize=R, loss_fraction=1e-2), {sample} sources per step",
                   "mesh": N, "R_cells": args.R, "sources_per_step": sample,
                   "cells_credited_per_source": cells, "cells_swept_per_source": swept,
                   "note": "the CPU sweeps its sub-box cube (more cells than the GPU path's octahedron); `value` credits the "
                           "octahedron cells of the metric, value_cells_swept the cells the CPU really visited"},
        "value_cells_swept": args.steps * sample * swept / secs,
        "cpu_baseline": {"value": value, "unit": "updates/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} of the 10^4 benchmark sources per step, {args.steps} steps, {secs:.1f} s"},
        "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# the reference's own CUDA kernel (oracle/_ref), for parity and as the like-for-like bar
# ---------------------------------------------------------------------------------------------------------------------

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libasora_ref.so")


class ReferenceKernel:
    """ctypes handle on oracle/_ref/libasora_ref.so: src/asora/*.cu compiled unmodified for sm_100 (oracle/ref_build.py)."""

    def __init__(self):
        dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
        self.dp, self.ip = dp, ip
        L = ctypes.CDLL(REF_SO)
        L.ref_device_init.argtypes = [ctypes.c_int, ctypes.c_int]
        L.ref_density_to_device.argtypes = [dp, ctypes.c_int]
        L.ref_photo_table_to_device.argtypes = [dp, dp, ctypes.c_int]
        L.ref_source_data_to_device.argtypes = [ip, dp, ctypes.c_int]
        L.ref_do_all_sources.argtypes = [ctypes.c_double, dp, ctypes.c_double, ctypes.c_double, dp, dp, dp, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int]
        L.ref_zero_coldens.argtypes = [ctypes.c_int, ctypes.c_int]
        self.L = L

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def setup(self, N, batch, ndens, thin, thick, pos_flat, flux_flat, ns):
        with quiet_stdout():
            assert self.L.ref_device_init(N, batch) == 0
            assert self.L.ref_zero_coldens(N, batch) == 0  # the reference reads its scratch uninitialised (oracle/ref_shim.cu)
            self.L.ref_density_to_device(ndens.ctypes.data_as(self.dp), N)
            self.L.ref_photo_table_to_device(thin.ctypes.data_as(self.dp), thick.ctypes.data_as(self.dp), thin.size)
            self.L.ref_source_data_to_device(pos_flat.ctypes.data_as(self.ip), flux_flat.ctypes.data_as(self.dp), ns)

    def sweep(self, R, dr, ndens, xh, phi, ns, N, dlogtau, numtau):
        dummy = np.zeros(1)
        with quiet_stdout():
            rc = self.L.ref_do_all_sources(R, dummy.ctypes.data_as(self.dp), SIG, dr, ndens.ctypes.data_as(self.dp),
                                           xh.ctypes.data_as(self.dp), phi.ctypes.data_as(self.dp), ns, N, -20.0, dlogtau, numtau)
        assert rc == 0, "reference kernel failed"

    def close(self):
        with quiet_stdout():
            self.L.ref_device_close()


def max_rel(a, b, floor=1e-12):
    """max |a - b| / max(|b|, floor * max|b|): the tolerance form of the parity tests (tests/test_gpu_parity.py)."""
    scale = np.maximum(np.abs(b), floor * np.abs(b).max())
    return float((np.abs(a - b) / scale).max())


def reference_gpu_timing(N, R, dr, ndens, xh, thin, thick, dlogtau, numtau, cells, nsrc_ref=2000):
    """The reference's do_all_sources (its benchmark's timed call: pageable host buffers, 2 x 8 N^3 bytes over PCIe
    per call, one launch + device synchronise per batch) on this GPU; a call with zero sources times its copies alone."""
    if not ReferenceKernel.available():
        return {"unavailable": "oracle/_ref/libasora_ref.so not built"}
    ref = ReferenceKernel()
    su = _load("_bench_sourceutils", "utils", "sourceutils.py")
    srcpos = su.generate_test_sources(N, nsrc_ref, seed=100)
    flux = 10 ** np.random.default_rng(100).normal(0, 0.5, size=nsrc_ref)
    pos_flat, flux_flat = su.format_sources(srcpos, flux)
    phi = np.zeros(N ** 3)
    out = {"sources": nsrc_ref, "what": "src/asora compiled unmodified for sm_100 (oracle/_ref), its do_all_sources call with "
                                        "pageable host buffers as its benchmark times it; kernel_ms = that minus the same call "
                                        "with zero sources (its two PCIe copies)"}
    for batch in (64, 128):
        ref.setup(N, batch, ndens, thin, thick, pos_flat, flux_flat, nsrc_ref)
        best, best0 = 1e30, 1e30
        for rep in range(3):
            t0 = time.perf_counter()
            ref.sweep(R, dr, ndens, xh, phi, nsrc_ref, N, dlogtau, numtau)
            best = min(best, time.perf_counter() - t0)
            t0 = time.perf_counter()
            ref.sweep(R, dr, ndens, xh, phi, 0, N, dlogtau, numtau)
            best0 = min(best0, time.perf_counter() - t0)
        ref.close()
        k = max(best - best0, 1e-9)
        out[f"batch{batch}"] = {"e2e_ms": 1e3 * best, "e2e_updates_per_s": nsrc_ref * cells / best,
                                "kernel_ms": 1e3 * k, "kernel_updates_per_s": nsrc_ref * cells / k}
    return out


# ---------------------------------------------------------------------------------------------------------------------
# secondary figures
# ---------------------------------------------------------------------------------------------------------------------

def eor_inputs(N, nsrc, seed=244):
    su = _load("_bench_sourceutils", "utils", "sourceutils.py")
    rng = np.random.default_rng(seed)
    srcpos = su.generate_test_sources(N, nsrc, seed=seed)
    flux = 10 ** rng.normal(5.0, 0.5, size=nsrc)  # ~1e53 photons/s: a few cells ionised per source and step
    g = rng.normal(size=(N, N, N))
    ndens = 1.87e-7 * (1.0 + 9.0) ** 3 * np.exp(0.5 * g - 0.125)
    xh = np.full((N, N, N), 2e-4)
    temp = np.full((N, N, N), 1e4)
    dr = 244.0 / 0.7 * MPC / N / (1.0 + 9.0)
    R = 15.0 * N * 0.7 / 244.0
    return srcpos, flux, ndens, xh, temp, dr, R


def eor_step(p, thin, thick, dlogtau, N=250, nsrc=100000):
    """One full evolve3D time step of the c2ray_244paper configuration restated synthetically (the real
    density / halo files are not in the reference checkout): 250^3, 10^5 seeded sources, R_max = 15 cMpc =
    10.76 cells (test/paper_eor_simulation/parameters.yml:84), log-normal density, dt = 10 Myr.  Wall clock of
    the whole call with pageable numpy inputs: uploads, every ray-tracing + chemistry iteration until convergence,
    downloads.  Parity: (a) one convergence iteration (ray tracing of a 10^3-source subset + chemistry pass) per cell
    against the CPU oracle loop (pyc2ray/evolve.py:168-240); (b) the full step with the cells outside the R sphere swept
    as well (the reference's cell set) against the sphere-only step evolve3D uses."""
    import oracle
    from pyc2ray_b200.lib import _cabi, libasora
    srcpos, flux, ndens, xh, temp, dr, R = eor_inputs(N, nsrc)
    dt = 1e7 * 3.15576e7
    p.device_init(N, 96)
    try:
        p.photo_table_to_device(thin, thick)
        best, niter, mean_x, walls, loops = None, 0, 0.0, [], []
        for rep in range(4):   # the first call also builds the sweep plans and faults in the staging buffers
            t0 = time.perf_counter()
            x, phi = p.evolve3D(dt, dr, flux, srcpos, True, 1000, 64, 1e-2, temp, ndens, xh, thin, thick,
                                -20.0, dlogtau, R, 1e-4, SIG, *CHEM, logfile=None, quiet=True)
            wall = time.perf_counter() - t0
            best = wall if best is None else min(best, wall)
            walls.append(round(1e3 * wall, 1))
            loops.append(round(1e3 * p.evolve3D.last_loop_seconds, 1))
            niter, mean_x = p.evolve3D.last_niter, float(x.mean())
        # (b) the same step on the reference's full cell set (sphere-only off): evolve3D switches sphere-only on itself,
        # so the loop is driven here through the same C ABI calls
        su = _load("_bench_sourceutils", "utils", "sourceutils.py")
        pos_flat, flux_flat = su.format_sources(srcpos, flux)
        L, check, dptr = _cabi.L, _cabi.check, _cabi.dptr
        libasora.source_data_to_device(pos_flat, flux_flat, nsrc)
        for buf, arr in ((_cabi.BUF_NDENS, ndens), (_cabi.BUF_TEMP, temp), (_cabi.BUF_XH, xh)):
            check(L.asora_buffer_upload(buf, dptr(np.ascontiguousarray(arr.ravel()))))
        for b in (_cabi.BUF_XH_AV, _cabi.BUF_XH_INTERMED):
            check(L.asora_buffer_copy(b, _cabi.BUF_XH))
        flag, s1, s0 = ctypes.c_int(0), ctypes.c_double(0), ctypes.c_double(0)
        prev1 = prev0 = 2.0 * N ** 3
        conv = min(int(1e-4 * N ** 3), (nsrc - 1) / 3)
        it_full = 0
        while True:
            it_full += 1
            check(L.asora_raytrace_device(R, SIG, dr, 0, nsrc, -20.0, dlogtau, thin.size, 1))
            check(L.asora_global_pass_device(dt, *CHEM, ctypes.byref(flag), ctypes.byref(s1), ctypes.byref(s0)))
            r1 = abs((s1.value - prev1) / s1.value) if s1.value > 0 else 1.0
            r0 = abs((s0.value - prev0) / s0.value) if s0.value > 0 else 1.0
            if flag.value < conv or (r1 < 1e-4 and r0 < 1e-4) or it_full > 200:
                break
            prev1, prev0 = s1.value, s0.value
        x_full = np.empty(N ** 3)
        check(L.asora_buffer_download(_cabi.BUF_XH_INTERMED, dptr(x_full)))
        full_vs_sphere = {"iterations_full_cell_set": it_full, "mean_xh_full_cell_set": float(x_full.mean()),
                          "max_abs_diff_xh": float(np.abs(x_full - x.ravel()).max()),
                          "rel_diff_mean_xh": abs(float(x_full.mean()) - mean_x) / mean_x}
        # (a) one convergence iteration per cell against the oracle: 10^3-source subset, sphere-only on (what evolve3D runs)
        ns1 = 1000
        libasora.source_data_to_device(np.ascontiguousarray(pos_flat[:3 * ns1]), np.ascontiguousarray(flux_flat[:ns1]), ns1)
        check(L.asora_buffer_upload(_cabi.BUF_XH_AV, dptr(np.ascontiguousarray(xh.ravel()))))
        check(L.asora_buffer_copy(_cabi.BUF_XH_INTERMED, _cabi.BUF_XH))
        check(L.asora_set_sphere_only(1))
        check(L.asora_raytrace_device(R, SIG, dr, 0, ns1, -20.0, dlogtau, thin.size, 1))
        check(L.asora_set_sphere_only(0))
        phi1 = np.empty(N ** 3)
        check(L.asora_buffer_download(_cabi.BUF_PHI_ION, dptr(phi1)))
        check(L.asora_global_pass_device(dt, *CHEM, ctypes.byref(flag), ctypes.byref(s1), ctypes.byref(s0)))
        xav1, xint1 = np.empty(N ** 3), np.empty(N ** 3)
        check(L.asora_buffer_download(_cabi.BUF_XH_AV, dptr(xav1)))
        check(L.asora_buffer_download(_cabi.BUF_XH_INTERMED, dptr(xint1)))
    finally:
        p.device_close()
    phi_o, _, _ = oracle.asora_do_all_sources(R, SIG, dr, ndens.ravel(), xh.ravel(), pos_flat[:3 * ns1], flux_flat[:ns1], N,
                                              thin, thick, -20.0, dlogtau, thin.size, nthreads=host_threads())
    flat = lambda a: np.ascontiguousarray(a.ravel())
    # the chemistry pass of the oracle is fed the rates the GPU produced: its fixed point stops on a threshold
    # (chemistry.f90:182-189), so rates that differ in the 13th digit could otherwise stop a borderline cell one
    # iteration apart, which is not a property of the chemistry kernel
    xav_o, xint_o = flat(xh).copy(), flat(xh).copy()
    flag_o = oracle.global_pass(dt, flat(ndens), flat(temp), flat(xh), xav_o, xint_o, phi1, *CHEM)
    rel_av = np.abs(xav1 - xav_o) / np.abs(xav_o)
    one_iter = {"sources": ns1, "phi_max_rel": max_rel(phi1, phi_o), "xh_av_max_rel": float(rel_av.max()),
                "xh_av_cells_above_1e-10": int((rel_av > 1e-10).sum()),
                "xh_intermed_max_rel": max_rel(xint1, xint_o, 1e-300), "conv_flag": [int(flag.value), int(flag_o)],
                "vs": "oracle/ C port (raytracing.cu; chemistry.f90 on the GPU's rates)",
                "tolerance": {"phi": PARITY_TOL, "xh_intermed": 1e-9, "xh_av": 1e-6, "xh_av_cells_above_1e-10": "<= 1e-5 of the cells"},
                "note": "the time-averaged fraction comes out of a fixed point that stops on a threshold (chemistry.f90:182-189): in "
                        "a few cells per million the last-ulp difference between CUDA's and glibc's exp() stops it one iteration "
                        "apart, which moves xh_av by the (small) size of that last step; everywhere else the agreement is 1e-12"}
    ok = (one_iter["phi_max_rel"] <= PARITY_TOL and one_iter["xh_av_max_rel"] <= 1e-6 and
          one_iter["xh_av_cells_above_1e-10"] <= 1e-5 * N ** 3 and
          one_iter["xh_intermed_max_rel"] <= 1e-9 and flag.value == flag_o and full_vs_sphere["max_abs_diff_xh"] <= 1e-9)
    return {"ms": 1e3 * best, "ms_all_calls": walls, "ms_convergence_loop": loops, "iterations": niter, "mean_xh_after": mean_x,
            "config": f"synthetic c2ray_244paper step: {N}^3, {nsrc} sources, R={R:.2f} cells, dt=10 Myr, "
                      "evolve3D with pageable numpy inputs incl. all host<->device copies",
            "parity": {"ok": bool(ok), "one_iteration_vs_oracle": one_iter, "full_cell_set_vs_sphere_only": full_vs_sphere}}


def chemistry_pass(p, N, passes=20):
    """Secondary figure: the ionisation-ODE kernel (chemistry.f90:13-316 as global_pass_kernel) on device-resident
    grids, wall clock per asora_global_pass_device call (each returns conv_flag, i.e. ends with a device->host read).
    56 algorithmic bytes per cell and pass."""
    from pyc2ray_b200.lib import _cabi
    from pyc2ray_b200.lib._cabi import L, check, dptr
    rng = np.random.default_rng(3)
    n3 = N ** 3
    p.device_init(N, 8)
    try:
        for buf, arr in ((_cabi.BUF_NDENS, 1e-3 * np.exp(0.5 * rng.normal(size=n3))), (_cabi.BUF_TEMP, np.full(n3, 1e4)),
                         (_cabi.BUF_XH, np.full(n3, 2e-4)), (_cabi.BUF_PHI_ION, 10 ** rng.uniform(-16, -12, size=n3))):
            check(L.asora_buffer_upload(buf, dptr(np.ascontiguousarray(arr))))
        for b in (_cabi.BUF_XH_AV, _cabi.BUF_XH_INTERMED):
            check(L.asora_buffer_copy(b, _cabi.BUF_XH))
        f, a, b2 = ctypes.c_int(0), ctypes.c_double(0), ctypes.c_double(0)
        times = []
        for _ in range(passes):
            t0 = time.perf_counter()
            check(L.asora_global_pass_device(3.15576e13, *CHEM, ctypes.byref(f), ctypes.byref(a), ctypes.byref(b2)))
            times.append(time.perf_counter() - t0)
    finally:
        p.device_close()
    steady = float(np.median(times[passes // 2:]))
    return {"ms_first_pass": 1e3 * times[0], "ms_per_pass": 1e3 * steady, "cells": n3, "bytes_per_cell": 56,
            "achieved_gbs": 56 * n3 / steady / 1e9,
            "note": "first pass: temperature-factor fill and several fixed-point iterations per cell; later passes are "
                    "near convergence"}


def small_evolve_case(N, ns, R, seed=5):
    """Inputs of the multi-GPU parity check: log-normal density, bubbles, ns sources (every rank builds the same)."""
    su = _load("_bench_sourceutils", "utils", "sourceutils.py")
    rng = np.random.default_rng(seed)
    srcpos = su.generate_test_sources(N, ns, seed=seed)
    flux = 10 ** rng.normal(6.0, 0.5, size=ns)
    ndens = 1e-3 * np.exp(rng.normal(size=(N, N, N)) * 0.8 - 0.32)
    xh = np.full((N, N, N), 2e-4)
    temp = np.full((N, N, N), 1e4)
    return srcpos, flux, ndens, xh, temp, 3e20, R


def multi_gpu_parity(p, thin, thick, dlogtau, world):
    """evolve3D_dist with every decomposition against single-GPU evolve3D on the same inputs, on every rank (the content
    of tests/test_gpu_multi.py, which a 1-GPU test box cannot run).  Mesh 64, R = 3.5 so that slabs exist up to 8 ranks."""
    import torch
    import torch.distributed as dist
    N, ns, R = 64, 96, 3.5
    srcpos, flux, ndens, xh, temp, dr, R = small_evolve_case(N, ns, R)
    out = {"mesh": N, "sources": ns, "R_cells": R, "vs": "single-GPU evolve3D on every rank, max over ranks"}
    p.device_init(N, 8)
    try:
        p.photo_table_to_device(thin, thick)
        args = (3.15576e13, dr, flux, srcpos)
        tail = (temp, ndens, xh, thin, thick, -20.0, dlogtau, R, 1e-4, SIG) + CHEM
        x1, phi1 = p.evolve3D(*args, True, 1000, 64, 1e-2, *tail, logfile=None, quiet=True)
        it1 = p.evolve3D.last_niter
        for dec in ("list", "rsag", "slab", "auto"):
            try:
                xd, phid = p.evolve3D_dist(*args, *tail, logfile=None, quiet=True, decomposition=dec)
            except ValueError as e:  # a decomposition that does not exist for this rank count
                out[dec] = {"skipped": str(e)}
                continue
            t = torch.tensor([max_rel(xd, x1, 1e-300), max_rel(phid, phi1), float(p.evolve3D.last_niter != it1)],
                             dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[dec] = {"xh_max_rel": float(t[0]), "phi_max_rel": float(t[1]), "same_iterations": bool(t[2] == 0)}
    finally:
        p.device_close()
    out["ok"] = all(("skipped" in v) or (v["xh_max_rel"] <= 1e-10 and v["phi_max_rel"] <= 1e-9 and v["same_iterations"])
                    for k, v in out.items() if isinstance(v, dict))
    return out


def strong_512(p, thin, thick, dlogtau, rank, world, N=512, nsrc=100000):
    """BASELINE config 5: 512^3, 10^5 sources, R = 10.76, one evolve3D time step split over the ranks
    (evolve3D_dist, decomposition "auto") against the same step on one GPU (rank 0 alone)."""
    import torch
    import torch.distributed as dist
    srcpos, flux, ndens, xh, temp, dr, _ = eor_inputs(N, nsrc, seed=512)
    R = 10.76  # cells, as in the 250^3 step (BASELINE config 5)
    args = (1e7 * 3.15576e7, dr, flux, srcpos)
    tail = (temp, ndens, xh, thin, thick, -20.0, dlogtau, R, 1e-4, SIG) + CHEM
    out = {"mesh": N, "sources": nsrc, "R_cells": R}
    p.device_init(N, 64)
    try:
        p.photo_table_to_device(thin, thick)
        t1 = None
        if rank == 0:  # the one-GPU step
            for rep in range(2):
                t0 = time.perf_counter()
                x1, _ = p.evolve3D(*args, True, 1000, 64, 1e-2, *tail, logfile=None, quiet=True)
                t1 = min(t1 or 1e30, time.perf_counter() - t0)
            out["iterations"] = p.evolve3D.last_niter
            loop1 = p.evolve3D.last_loop_seconds
        dist.barrier()
        def timed(**kw):
            best, res = None, None
            for rep in range(2):
                dist.barrier()
                t0 = time.perf_counter()
                res, _ = p.evolve3D_dist(*args, *tail, logfile=None, quiet=True, decomposition="auto", **kw)
                torch.cuda.synchronize()
                el = torch.tensor([time.perf_counter() - t0, p.evolve3D.last_loop_seconds], dtype=torch.float64, device="cuda")
                dist.all_reduce(el, op=dist.ReduceOp.MAX)
                if best is None or float(el[0]) < best[0]:
                    best = (float(el[0]), float(el[1]), {k: round(1e3 * v, 2) for k, v in p.evolve3D.last_phase_seconds.items()})
            return best, res
        (tn, loopn, phases), xd = timed(io_rank=0)  # one host copy of the grids in, one out (rank 0), GPU-to-GPU broadcast of the inputs
        (tn_all, _, _), _ = timed()                 # the reference's semantics: every rank passes and receives all grids
        if rank == 0:
            out.update({"ms_1gpu": 1e3 * t1, "ms": 1e3 * tn, "speedup": t1 / tn, "efficiency_vs_1gpu": t1 / tn / world,
                        "ms_every_rank_copies": 1e3 * tn_all, "xh_max_rel_vs_1gpu": max_rel(xd, x1, 1e-300),
                        "device_loop": {"ms_1gpu": 1e3 * loop1, "ms": 1e3 * loopn, "speedup": loop1 / loopn,
                                        "efficiency_vs_1gpu": loop1 / loopn / world, "phases_ms_rank0": phases,
                                        "what": "the convergence loop alone (sweeps, exchanges, chemistry, 3 scalars per iteration to the "
                                                "host), grids resident on the devices"},
                        "what": "wall clock of the whole evolve3D(_dist) call incl. the host<->device copies of five 1.07 GB "
                                "grids; ms: io_rank=0 (rank 0 reads and returns the grids, inputs broadcast over NVLink); "
                                "ms_every_rank_copies: every rank uploads and downloads all five grids (reference semantics)",
                        "limited_by": "the host copies (5 x 1.07 GB through pageable memory at ~40 GB/s = 0.13 s) and the serial "
                                      "host-side preparation do not shrink with the rank count; the device loop does"})
    finally:
        p.device_close()
    return out


def run_sweep_table(args):
    """--sweep: throughput against radius and source count, the axes of the reference's raytracing_benchmark."""
    env = dict(os.environ, ASORA_SWEEP_MESH=str(args.mesh))
    return subprocess.call([sys.executable, os.path.join(ROOT, "scripts", "radius_sweep.py")], env=env)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--R", type=float, default=30.0)
    ap.add_argument("--nsrc", type=int, default=10000)
    ap.add_argument("--mesh", type=int, default=256, help="256 (BASELINE) or 250 (the paper's own mesh)")
    ap.add_argument("--cpu-sample", type=int, default=2048, help="most sources per CPU-baseline step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-eor", action="store_true", help="skip the secondary EoR-step timing")
    ap.add_argument("--no-refgpu", action="store_true", help="skip the reference-kernel timing")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling and multi-GPU parity legs")
    ap.add_argument("--sweep", action="store_true", help="print the radius x source-count table instead of the bench line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.sweep:
        return sys.exit(run_sweep_table(args))

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local) if world > 1 else "single rank: not bound"
    if world > 1:
        with quiet_stdout():  # NCCL prints its version line to stdout when the first communicator is created
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.all_reduce(torch.zeros(1, device="cuda"))
            torch.cuda.synchronize()

    import pyc2ray_b200 as p
    from pyc2ray_b200.lib import _cabi, libasora
    from pyc2ray_b200.lib._cabi import L, check, dptr
    from pyc2ray_b200.parallel import device_tensor, shard_bounds

    N, R, K, W = args.mesh, args.R, args.steps, args.warmup
    dr = 3 * MPC / N
    thin, thick, dlogtau, numtau = tables()
    srcpos, flux, pos_flat, flux_flat, ndens, xh = workload(N, args.nsrc, seed=100 + rank)
    cells = int(L.asora_cells_per_source(N, R))
    units_per_step = args.nsrc * cells
    failures = []

    p.device_init(N, 64)
    stream = torch.cuda.Stream()
    check(L.asora_set_stream(ctypes.c_void_p(stream.cuda_stream)))
    p.photo_table_to_device(thin, thick)
    libasora.source_data_to_device(pos_flat, flux_flat, args.nsrc)
    libasora.density_to_device(ndens, N)
    check(L.asora_buffer_upload(_cabi.BUF_XH_AV, dptr(xh)))
    phi_t = device_tensor(L.asora_device_buffer(_cabi.BUF_PHI_ION), N ** 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    kernel_ms = []

    def step_device(count=args.nsrc, begin=0):
        check(L.asora_raytrace_device(R, SIG, dr, begin, count, -20.0, dlogtau, numtau, 1))
        if world > 1:
            dist.all_reduce(phi_t, op=dist.ReduceOp.SUM)

    sampler = ClockSampler(local)
    allreduce_check = strong = None
    with torch.cuda.stream(stream):
        # ---- device-resident throughput -----------------------------------------------------------
        for _ in range(W):
            step_device()
        barrier()
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches = 0
        e0.record(stream)
        for _ in range(K):
            step_device()
            nl = ctypes.c_int(0)
            L.asora_last_sweep_stats(None, ctypes.byref(nl), None, None, None, None)
            launches += nl.value
        e1.record(stream)
        barrier()
        dev_ms = reduce_max(e0.elapsed_time(e1))
        clocks = sampler.stop() if rank == 0 else None

        # ---- sweep kernel alone (roofline), CUDA events on the launching stream --------------------
        sweep_ms = []
        for _ in range(K):
            check(L.asora_raytrace_device(R, SIG, dr, 0, args.nsrc, -20.0, dlogtau, numtau, 1))
            check(L.asora_sync())
            ms, kms = ctypes.c_float(0.0), ctypes.c_float(0.0)
            L.asora_last_sweep_stats(None, None, None, None, None, ctypes.byref(ms))
            check(L.asora_last_sweep_kernel_ms(ctypes.byref(kms)))
            sweep_ms.append(ms.value)      # opacity pre-pass + zeroing + sweep kernel + division pass
            kernel_ms.append(kms.value)    # the sweep kernel alone
        variant = ctypes.c_int(0)
        levels = ctypes.c_int(0)
        qmax = ctypes.c_int(0)
        L.asora_last_sweep_stats(ctypes.byref(variant), None, None, ctypes.byref(qmax), ctypes.byref(levels), None)

        # ---- N > 1: the all-reduced rates against the per-rank partials, summed on rank 0 (32-plane slab) -------
        if world > 1:
            nslab = 32 * N * N
            check(L.asora_raytrace_device(R, SIG, dr, 0, args.nsrc, -20.0, dlogtau, numtau, 1))
            torch.cuda.synchronize()
            partial = phi_t[:nslab].clone()
            gathered = [torch.empty_like(partial) for _ in range(world)] if rank == 0 else None
            dist.gather(partial, gathered, dst=0)
            dist.all_reduce(phi_t, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
            if rank == 0:
                total = torch.zeros_like(partial)
                for t in gathered:      # rank order, on the device, fp64
                    total += t
                a, b = phi_t[:nslab].cpu().numpy(), total.cpu().numpy()
                allreduce_check = {"max_rel": max_rel(a, b), "cells": int(nslab), "planes": 32,
                                   "what": "NCCL all-reduce of phi_ion vs the per-rank partial grids gathered to rank 0 and "
                                           "summed there in rank order"}
                if allreduce_check["max_rel"] > 1e-10:
                    failures.append("allreduce_check")

        # ---- strong scaling (i): the fixed 10^4-source problem of rank 0 split over the ranks ------------------------
        if world > 1 and not args.no_strong:
            _, _, pos0, flux0, _, _ = workload(N, args.nsrc, seed=100)
            b0, b1 = shard_bounds(args.nsrc, rank, world)
            t1 = None
            if rank == 0:  # the whole list on one GPU
                libasora.source_data_to_device(pos0, flux0, args.nsrc)
                for rep in range(3):
                    s0_, s1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s0_.record(stream)
                    check(L.asora_raytrace_device(R, SIG, dr, 0, args.nsrc, -20.0, dlogtau, numtau, 1))
                    s1_.record(stream)
                    torch.cuda.synchronize()
                    t1 = min(t1 or 1e30, s0_.elapsed_time(s1_))
            libasora.source_data_to_device(np.ascontiguousarray(pos0[3 * b0:3 * b1]), np.ascontiguousarray(flux0[b0:b1]), b1 - b0)
            tn = None
            for rep in range(4):
                barrier()
                s0_, s1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0_.record(stream)
                step_device(count=b1 - b0)
                s1_.record(stream)
                barrier()
                el = reduce_max(s0_.elapsed_time(s1_))
                if rep > 0:
                    tn = min(tn or 1e30, el)
            # the same problem sharded by position (parallel.slab_edges / SlabHalo, what evolve3D_dist's "auto" uses when the
            # slabs are wide enough): a rank's rates stay within its planes +- a halo, so two halo exchanges with the
            # neighbours replace the N^3 all-reduce; every rank ends with the complete rates of its OWN planes
            slab = None
            from pyc2ray_b200.parallel import slab_edges, SlabHalo
            edges, hh = slab_edges(pos0[0::3], N, world, R)
            if edges is not None:
                halo = SlabHalo(edges, hh, N, rank, world, peer=os.environ.get("ASORA_PEER_HALO", "1") != "0",
                                stream_ordered=True)   # asora_set_stream above: sweeps, collectives and halo kernels share a stream
                o, cnt = halo.own_cells()
                want = phi_t[o:o + cnt].clone()          # the all-reduced rates of the list-order run, my planes
                mine = (pos0[0::3] >= halo.lo) & (pos0[0::3] < halo.hi)
                nmine = int(mine.sum())
                pos_m = np.ascontiguousarray(pos0.reshape(-1, 3)[mine].ravel())
                libasora.source_data_to_device(pos_m, np.ascontiguousarray(flux0[mine]), nmine)
                # planes this rank's sweeps can read: the full cell set reaches q_max planes beyond its sources (rates: hh)
                reach = min(qmax.value, N // 2)
                # (only when that is well under the whole grid: a restricted sweep gives up the z-face grid copies)
                if (halo.hi - halo.lo) + 2 * reach <= 0.6 * N:
                    check(L.asora_set_active_slab((halo.lo - reach) % N, (halo.hi - halo.lo) + 2 * reach))
                ts = None
                for rep in range(4):
                    barrier()
                    s0_, s1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s0_.record(stream)
                    check(L.asora_raytrace_device(R, SIG, dr, 0, nmine, -20.0, dlogtau, numtau, 1))
                    halo.reduce_phi_(phi_t)
                    s1_.record(stream)
                    barrier()
                    el = reduce_max(s0_.elapsed_time(s1_))
                    if rep > 0:
                        ts = min(ts or 1e30, el)
                torch.cuda.synchronize()
                check(L.asora_set_active_slab(0, 0))
                peer_mode = halo.peer
                halo.close()
                got = phi_t[o:o + cnt]
                scale = float(want.abs().max().item())
                rel = ((got - want).abs() / torch.clamp(want.abs(), min=1e-12 * scale)).max().reshape(1)
                dist.all_reduce(rel, op=dist.ReduceOp.MAX)
                counts = torch.tensor([float(nmine)], device="cuda")
                dist.all_reduce(counts, op=dist.ReduceOp.MAX)
                slab = {"ms": ts, "halo_planes": hh, "max_sources_per_rank": int(counts.item()),
                        "halo_exchange": "neighbours' halo planes read over NVLink from their GPUs' memory (CUDA IPC), one kernel per "
                                         "neighbour" if peer_mode else "NCCL send/recv",
                        "max_rel_vs_allreduce": float(rel.item()),
                        "what": "sources sharded by x-plane ranges of equal source counts, sweep + two halo reductions of "
                                f"{hh} planes ({hh * N * N * 8 / 1e6:.0f} MB each) from the neighbouring ranks; every rank ends with "
                                "the complete rates of its own planes (the input of its share of the chemistry)"}
                if slab["max_rel_vs_allreduce"] > 1e-10:
                    failures.append("strong.fixed_256.slab")
            if rank == 0:
                strong = {"fixed_256": {"sources_total": args.nsrc, "ms_1gpu": t1, "ms": tn, "speedup": t1 / tn,
                                        "efficiency_vs_1gpu": t1 / tn / world,
                                        "what": f"the {args.nsrc} sources of rank 0's list split in list order over {world} ranks, "
                                                "sweep + all-reduce of phi_ion (134 MB), CUDA events, max over ranks",
                                        "limited_by": "the all-reduce and the fixed passes (opacity pre-pass, zeroing, division: "
                                                      "~0.3 ms) do not shrink with the shard"}}
                if slab is not None:
                    slab["speedup"] = t1 / slab["ms"]
                    slab["efficiency_vs_1gpu"] = t1 / slab["ms"] / world
                    strong["fixed_256"]["slab"] = slab
                best_ms = min(tn, slab["ms"]) if slab is not None else tn
                strong["fixed_256"]["best"] = {"decomposition": "slab" if (slab is not None and slab["ms"] < tn) else "list",
                                               "ms": best_ms, "efficiency_vs_1gpu": t1 / best_ms / world}
            libasora.source_data_to_device(pos_flat, flux_flat, args.nsrc)

        # ---- parity: the first 8 sources of rank 0's inputs, default launch shape, against the reference ----------
        _, _, pos0, flux0, _, _ = (None, None, pos_flat, flux_flat, None, None) if rank == 0 else workload(N, 8, seed=100)
        npar = min(8, args.nsrc)
        libasora.source_data_to_device(np.ascontiguousarray(pos0[:3 * npar]), np.ascontiguousarray(flux0[:npar]), npar)
        check(L.asora_raytrace_device(R, SIG, dr, 0, npar, -20.0, dlogtau, numtau, 1))
        phi_par = np.empty(N ** 3)
        check(L.asora_buffer_download(_cabi.BUF_PHI_ION, dptr(phi_par)))
        vpar = ctypes.c_int(0)
        L.asora_last_sweep_stats(ctypes.byref(vpar), None, None, None, None, None)
        libasora.source_data_to_device(pos_flat, flux_flat, args.nsrc)
        torch.cuda.synchronize()

        # ---- secondary: the same rates with the cells outside the R sphere left out (what evolve3D uses) -----
        # phi_ion is bit-identical (tests/test_gpu_parts.py); fewer cells are swept, so this is NOT the headline:
        # it is reported in the reference paper's own unit, time per source and sphere cell.
        sphere_ms, sphere_updates = [], ctypes.c_int64(0)
        check(L.asora_set_sphere_only(1))
        for i in range(K + 1):
            check(L.asora_raytrace_device(R, SIG, dr, 0, args.nsrc, -20.0, dlogtau, numtau, 1))
            check(L.asora_sync())
            ms = ctypes.c_float(0.0)
            L.asora_last_sweep_stats(None, None, ctypes.byref(sphere_updates), None, None, ctypes.byref(ms))
            if i > 0:
                sphere_ms.append(ms.value)
        check(L.asora_set_sphere_only(0))

        # ---- secondary: deterministic accumulation (asora_set_deterministic): cost, and two runs compared bit for bit ----
        det = None
        if rank == 0:
            check(L.asora_set_deterministic(1))
            det_ms, grids = [], []
            for i in range(3):
                check(L.asora_raytrace_device(R, SIG, dr, 0, args.nsrc, -20.0, dlogtau, numtau, 1))
                check(L.asora_sync())
                ms = ctypes.c_float(0.0)
                L.asora_last_sweep_stats(None, None, None, None, None, ctypes.byref(ms))
                det_ms.append(ms.value)
                if i > 0:
                    g_ = np.empty(N ** 3)
                    check(L.asora_buffer_download(_cabi.BUF_PHI_ION, dptr(g_)))
                    grids.append(g_)
            check(L.asora_set_deterministic(0))
            check(L.asora_raytrace_device(R, SIG, dr, 0, args.nsrc, -20.0, dlogtau, numtau, 1))
            g0 = np.empty(N ** 3)
            check(L.asora_buffer_download(_cabi.BUF_PHI_ION, dptr(g0)))
            det = {"ms_per_step": float(min(det_ms[1:])), "default_ms_per_step": float(np.mean(sweep_ms)),
                   "bit_identical_runs": bool(np.array_equal(grids[0], grids[1])),
                   "max_rel_vs_default": max_rel(grids[0], g0, 1e-15),
                   "what": "rates accumulated as 128-bit fixed-point integers (two 64-bit integer REDs per rated cell) instead of "
                           "one fp64 RED: run-to-run and launch-shape independent phi_ion"}
            if not det["bit_identical_runs"]:
                failures.append("deterministic")

        # ---- end to end through the reference-facing call: pageable numpy buffers (headline), then pinned -----
        dummy = np.zeros(1)
        phi_np = np.zeros(N ** 3)

        def e2e_loop(xh_buf, phi_buf):
            def step():
                if world == 1:
                    libasora.do_all_sources(R, dummy, SIG, dr, dummy, xh_buf, phi_buf, args.nsrc, N, -20.0, dlogtau, numtau)
                else:
                    # the same call for a source-sharded job: ONE xh_av grid goes in (rank 0 uploads it, the other GPUs
                    # receive it over NVLink), the rates are all-reduced on the devices, ONE phi_ion grid comes out
                    libasora.do_all_sources(R, dummy, SIG, dr, dummy, xh_buf, phi_buf, args.nsrc, N, -20.0, dlogtau, numtau,
                                            group=True, download=(rank == 0), xh_from=0)
            for _ in range(max(1, W // 2)):
                step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(K):
                step()
            barrier()
            return reduce_max(time.perf_counter() - t0)

        e2e_s = e2e_loop(xh, phi_np)
        phi_checksum = float(phi_np.sum())
        xh_pin = torch.from_numpy(xh).pin_memory()
        phi_pin = torch.zeros(N ** 3, dtype=torch.float64).pin_memory()
        e2e_pinned_s = e2e_loop(xh_pin.numpy(), phi_pin.numpy())

    p.device_close()

    # ---- parity of the headline inputs (every rank; the CPU / reference-kernel side runs after our context is closed) ----
    if ReferenceKernel.available():
        ref = ReferenceKernel()
        ref.setup(N, 8, ndens, thin, thick, np.ascontiguousarray(pos0[:3 * npar]), np.ascontiguousarray(flux0[:npar]), npar)
        phi_ref = np.zeros(N ** 3)
        ref.sweep(R, dr, ndens, xh, phi_ref, npar, N, dlogtau, numtau)
        ref.close()
        vs = "oracle/_ref/libasora_ref.so (the reference's CUDA kernel, unmodified, sm_100)"
    else:
        import oracle
        phi_ref, _, _ = oracle.asora_do_all_sources(R, SIG, dr, ndens, xh, pos0[:3 * npar], flux0[:npar], N, thin, thick, -20.0,
                                                    dlogtau, numtau, nthreads=min(8, host_threads()))
        vs = "oracle/ C port of raytracing.cu"
    par_rel = reduce_max(max_rel(phi_par, phi_ref))
    parity = {"vs": vs, "max_rel": par_rel, "cells": int(np.count_nonzero(phi_ref)), "sources": npar, "tolerance": PARITY_TOL,
              "sweep_variant": vpar.value, "ranks_checked": world,
              "what": "phi_ion of the first sources of rank 0's bench inputs, default launch shape, per cell"}
    if not (par_rel <= PARITY_TOL):
        failures.append("parity")

    # ---- secondary legs on rank 0 / all ranks -------------------------------------------------------------------------
    eor = chem = refgpu = mgp = None
    if rank == 0 and world == 1 and not args.no_eor:
        eor = eor_step(p, thin, thick, dlogtau)
        if not eor["parity"]["ok"]:
            failures.append("eor_step.parity")
        chem = chemistry_pass(p, N)
    if rank == 0 and world == 1 and not args.no_refgpu:
        refgpu = reference_gpu_timing(N, R, dr, ndens, xh, thin, thick, dlogtau, numtau, cells)
    if world > 1 and not args.no_strong:
        mgp = multi_gpu_parity(p, thin, thick, dlogtau, world)
        if rank == 0 and not mgp["ok"]:
            failures.append("multi_gpu_parity")
        s512 = strong_512(p, thin, thick, dlogtau, rank, world)
        if rank == 0:
            strong["evolve_512"] = s512

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        k_ms = float(np.mean(kernel_ms))
        achieved = units_per_step * BYTES_PER_UPDATE / (k_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "sweep_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(f"R{R:g}")  # bytes per launch, ncu --set full (profiles/README.md)
        value = world * units_per_step * K / (dev_ms * 1e-3)
        kname = {1: "sweep_smem_kernel", 2: "sweep_grid_kernel", 3: "sweep_octant_kernel", 4: "sweep_cluster_kernel"}
        line = {
            "metric": "ASORA source-cell updates/s", "value": value, "unit": "updates/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"raytracing_benchmark (test/paper_tests/raytracing_benchmark/run_test.py) at {N}^3: "
                                   f"uniform ndens=1e-3, xh=2e-4, Teff=1e5 tables NumTau=20000, {args.nsrc} seeded random "
                                   f"sources per GPU, R={R:g} cells, {cells} visited cells per source",
                       "mesh": N, "sources_per_gpu": args.nsrc, "R_cells": R, "q_max": qmax.value,
                       "levels": levels.value, "sweep_variant": variant.value,
                       "l2": "inputs larger than L2 (3 x 134 MB grids vs 126 MB L2); no explicit flush",
                       "parallelism": f"source-sharded x{world}, one NCCL all-reduce of phi_ion ({8 * N ** 3 // 10 ** 6} MB) per step",
                       "host_binding": numa},
            "e2e": {"value": world * units_per_step * K / e2e_s, "unit": "updates/s",
                    "h2d_bytes_per_step": 8 * N ** 3, "d2h_bytes_per_step": 8 * N ** 3, "ms_per_step": 1e3 * e2e_s / K,
                    "host_buffers": "pageable numpy arrays, as libasora.do_all_sources receives them from pyc2ray "
                                    "(python_module.cu:21-68)" + ("; the job's one xh_av grid is uploaded by rank 0 and broadcast GPU to GPU, "
                                                                  "its one phi_ion grid downloaded by rank 0" if world > 1 else "")},
            "e2e_pinned": {"value": world * units_per_step * K / e2e_pinned_s, "unit": "updates/s",
                           "ms_per_step": 1e3 * e2e_pinned_s / K, "host_buffers": "page-locked"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": kname.get(variant.value, "?"),
                         "kernel_ms": k_ms, "bytes_per_update": BYTES_PER_UPDATE, "peak_source": peak_src,
                         "kernel_ms_scope": "CUDA events on the launching stream right before and after the sweep kernel's launch",
                         "sweep_ms_with_companions": float(np.mean(sweep_ms)),
                         "traffic_unit": "bytes per launch (ncu dram__bytes_read+write)",
                         "note": "bound by latency, the L1 data pipe and instruction issue, not by HBM (L2 hit rate 82-92 %): "
                                 "see DESIGN.md and profiles/README.md"},
            "clocks": clocks,
            "phi_checksum": phi_checksum,
            "parity": parity,
        }
        sp_ms = float(np.mean(sphere_ms))
        r_eff = min(R, N * 0.5 * 3 ** 0.5)
        line["sphere_only"] = {
            "ms_per_step": sp_ms, "updates_swept_per_step": int(sphere_updates.value),
            "updates_per_s": sphere_updates.value / (sp_ms * 1e-3),
            "ns_per_source_and_sphere_cell": sp_ms * 1e6 / (args.nsrc * 4.0 / 3.0 * np.pi * r_eff ** 3),
            "headline_ns_per_source_and_sphere_cell": float(np.mean(sweep_ms)) * 1e6 / (args.nsrc * 4.0 / 3.0 * np.pi * r_eff ** 3),
            "note": "identical phi_ion, cells outside the R sphere not swept (asora_set_sphere_only); the unit is the "
                    "reference paper's 3t/(Ns 4 pi R^3), published as 3.156 ns on a P100"}
        if eor is not None:
            line["eor_step"] = eor
        if chem is not None:
            chem["frac_of_hbm_peak"] = chem["achieved_gbs"] / peak
            line["chemistry_pass"] = chem
        if refgpu is not None:
            line["reference_gpu"] = refgpu
        if det is not None:
            line["deterministic"] = det
        if allreduce_check is not None:
            line["allreduce_check"] = allreduce_check
        if strong is not None:
            line["strong"] = strong
        if mgp is not None:
            line["multi_gpu_parity"] = mgp
        if not args.no_cpu and world == 1:  # the CPU baseline is reported at N = 1 only
            threads = host_threads()
            cal = cpu_reference_rate(R, 4 * threads, threads, N)
            sample = int(max(threads, min(args.cpu_sample, 12.0 * cal["sources"] / cal["seconds"])))  # ~10 s of CPU work
            run = cpu_reference_rate(R, sample, threads, N)
            one = cpu_reference_rate(R, 16, 1, N)  # the reference itself is serial (raytracing.f90:177)
            line["cpu_baseline"] = {"value": run["rate"], "unit": "updates/s", "cores": threads, "kind": "port",
                                    "value_cells_swept": run["rate_cells_swept"],
                                    "cells_swept_per_source": run["cells_swept_per_source"], "cells_credited_per_source": cells,
                                    "sample": f"{sample} of the {args.nsrc} sources, {run['seconds']:.1f} s, CPU port of "
                                              "src/c2ray/raytracing.f90 with the benchmark's sub-box settings",
                                    "single_thread": {"value": one["rate"], "unit": "updates/s",
                                                      "sample": f"16 sources, {one['seconds']:.1f} s"}}
        if failures:
            line["failed_checks"] = failures
        print(json.dumps(line), flush=True)
    if world > 1:
        fl = torch.tensor([float(len(failures))], device="cuda")
        dist.all_reduce(fl, op=dist.ReduceOp.MAX)
        dist.destroy_process_group()
        if fl.item() > 0:
            sys.exit(3)
    elif failures:
        sys.exit(3)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- ASORA source-cell updates/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--R 30] [--nsrc 10000]

Workload (config.workload): the reference's ray-tracing benchmark
(test/paper_tests/raytracing_benchmark/run_test.py) at the BASELINE size: 256^3 mesh, uniform
ndens = 1e-3, xh = 2e-4, black-body Teff = 1e5 tables with NumTau = 20000, 10^4 seeded random sources
per GPU, ray-tracing radius R = 30 cells (the radius of the reference's published 3.156 ns asymptote).
A step = one do_all_sources pass over the rank's 10^4 sources; with N > 1 ranks every rank sweeps its
own 10^4 sources (weak scaling) and the rate grids are summed by one NCCL all-reduce inside the step.

One JSON line is printed by rank 0:
  value          source-cell updates/s, all ranks, inputs resident in HBM (CUDA events, max over ranks)
  e2e            same metric through the reference-facing call libasora.do_all_sources with pinned HOST
                 buffers (H2D of xh_av and D2H of phi_ion inside the timed region)
  roofline       sweep kernel: 32 algorithmic bytes per update / mean launch duration (CUDA events on the
                 launching stream) vs the measured HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline   the CPU port of the reference's Fortran ray tracer (oracle/, Fortran flavour with the
                 benchmark's sub-box settings) on all host cores, on a bounded sample of the same sources
`--impl reference` times only that CPU port (the Fortran itself cannot be built: no Fortran compiler).
"""
import argparse
import ctypes
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
N_MESH = 256
BYTES_PER_UPDATE = 32  # read ndens 8 + read xh_av 8 + read-modify-write phi_ion 16


def workload(N, nsrc, seed):
    from pyc2ray_b200.utils.sourceutils import format_sources, generate_test_sources
    srcpos = generate_test_sources(N, nsrc, seed=seed)
    flux = 10 ** np.random.default_rng(seed).normal(0, 0.5, size=nsrc)
    ndens = np.full(N ** 3, 1e-3)
    xh = np.full(N ** 3, 2e-4)
    pos_flat, flux_flat = format_sources(srcpos, flux)
    return srcpos, flux, pos_flat, flux_flat, ndens, xh


def tables():
    from pyc2ray_b200.radiation import blackbody_tables
    thin, thick, dlogtau = blackbody_tables(1e5, False, -20.0, 4.0, 20000)
    return thin, thick, dlogtau, 20000


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


def cpu_reference_rate(R, nsrc_sample, threads, N=N_MESH, seed=100):
    """Time the CPU port of src/c2ray/raytracing.f90 (oracle, Fortran flavour) exactly as the reference's
    benchmark drives it: sub-boxes of size R, max_subbox 1000, loss_fraction 1e-2
    (raytracing_benchmark/run_test.py:38,88; parameters.yml).  Returns (updates/s, seconds, threads)."""
    import oracle
    srcpos, flux, _, _, ndens, xh = workload(N, nsrc_sample, seed)
    thin, thick, dlogtau, numtau = tables()
    nd3 = ndens.reshape(N, N, N)
    xh3 = xh.reshape(N, N, N)
    t0 = time.perf_counter()
    oracle.fortran_do_all_sources(flux, srcpos, 1000, max(1, int(R)), SIG, 3 * MPC / N, nd3, xh3, 1e-2, thin, thick,
                                  -20.0, dlogtau, R, NumTau=numtau, use_subbox=True, nthreads=threads)
    dt = time.perf_counter() - t0
    units = nsrc_sample * oracle.cells_per_source(N, R)
    return units / dt, dt, threads


def eor_step(p, thin, thick, dlogtau, N=250, nsrc=100000):
    """One full evolve3D time step of the c2ray_244paper configuration restated synthetically (the real
    density / halo files are not in the reference checkout): 250^3, 10^5 seeded sources, R_max = 15 cMpc =
    10.76 cells (test/paper_eor_simulation/parameters.yml:84), log-normal density, dt = 10 Myr.  Wall clock of
    the whole call: uploads, every ray-tracing + chemistry iteration until convergence, downloads."""
    from pyc2ray_b200.utils.sourceutils import generate_test_sources
    rng = np.random.default_rng(244)
    srcpos = generate_test_sources(N, nsrc, seed=244)
    flux = 10 ** rng.normal(5.0, 0.5, size=nsrc)  # ~1e53 photons/s: a few cells ionised per source and step
    g = rng.normal(size=(N, N, N))
    ndens = 1.87e-7 * (1.0 + 9.0) ** 3 * np.exp(0.5 * g - 0.125)
    xh = np.full((N, N, N), 2e-4)
    temp = np.full((N, N, N), 1e4)
    dr = 244.0 / 0.7 * MPC / N / (1.0 + 9.0)
    R = 15.0 * N * 0.7 / 244.0
    chem = (2.59e-13, -0.7, 1.3e-8 * 0.83 / 13.598 ** 2, 13.598 / 8.617e-5, 7.1e-7)
    p.device_init(N, 96)
    try:
        p.photo_table_to_device(thin, thick)
        best, niter, mean_x = None, 0, 0.0
        for rep in range(2):
            t0 = time.perf_counter()
            x, phi = p.evolve3D(1e7 * 3.15576e7, dr, flux, srcpos, True, 1000, 64, 1e-2, temp, ndens, xh, thin, thick,
                                -20.0, dlogtau, R, 1e-4, SIG, *chem, logfile=None, quiet=True)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            niter, mean_x = p.evolve3D.last_niter, float(x.mean())
    finally:
        p.device_close()
    return {"ms": 1e3 * best, "iterations": niter, "mean_xh_after": mean_x,
            "config": f"synthetic c2ray_244paper step: {N}^3, {nsrc} sources, R={R:.2f} cells, dt=10 Myr, "
                      "evolve3D incl. host<->device copies"}


def chemistry_pass(p, N=N_MESH, passes=20):
    """Secondary figure: the ionisation-ODE kernel (chemistry.f90:13-316 as global_pass_kernel) on device-resident 256^3
    grids, wall clock per asora_global_pass_device call (each returns conv_flag, i.e. ends with a device->host read).
    56 algorithmic bytes per cell and pass."""
    import ctypes
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
            check(L.asora_global_pass_device(3.15576e13, 2.59e-13, -0.7, 1.3e-8 * 0.83 / 13.598 ** 2, 13.598 / 8.617e-5, 7.1e-7,
                                             ctypes.byref(f), ctypes.byref(a), ctypes.byref(b2)))
            times.append(time.perf_counter() - t0)
    finally:
        p.device_close()
    steady = float(np.median(times[passes // 2:]))
    return {"ms_first_pass": 1e3 * times[0], "ms_per_pass": 1e3 * steady, "cells": n3, "bytes_per_cell": 56,
            "achieved_gbs": 56 * n3 / steady / 1e9,
            "note": "first pass includes the temperature-factor fill and several fixed-point iterations per cell; later "
                    "passes are near convergence"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    threads = oracle.max_threads()
    sample = max(threads, args.cpu_sample)
    if args.warmup > 0:
        cpu_reference_rate(args.R, threads, threads)
    rates, secs = [], []
    for _ in range(args.steps):
        r, dt, _ = cpu_reference_rate(args.R, sample, threads)
        rates.append(r)
        secs.append(dt)
    total_units = args.steps * sample * oracle.cells_per_source(N_MESH, args.R)
    value = total_units / sum(secs)
    line = {
        "impl": "reference", "metric": "ASORA source-cell updates/s", "value": value, "unit": "updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"raytracing_benchmark 256^3 uniform ndens=1e-3 xh=2e-4 R={args.R:g}, CPU port of "
                               f"src/c2ray/raytracing.f90 (subboxsize=R, loss_fraction=1e-2), {sample} sources per step"},
        "cpu_baseline": {"value": value, "unit": "updates/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} of the 10^4 benchmark sources per step, {args.steps} steps"},
        "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--R", type=float, default=30.0)
    ap.add_argument("--nsrc", type=int, default=10000)
    ap.add_argument("--cpu-sample", type=int, default=2048, help="sources per CPU-baseline step (~10 s on 16 cores)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-eor", action="store_true", help="skip the secondary EoR-step timing")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import pyc2ray_b200 as p
    from pyc2ray_b200.lib import _cabi, libasora
    from pyc2ray_b200.lib._cabi import L, check, dptr
    from pyc2ray_b200.parallel import device_tensor

    N, R, K, W = N_MESH, args.R, args.steps, args.warmup
    dr = 3 * MPC / N
    thin, thick, dlogtau, numtau = tables()
    srcpos, flux, pos_flat, flux_flat, ndens, xh = workload(N, args.nsrc, seed=100 + rank)
    units_per_step = args.nsrc * int(L.asora_cells_per_source(N, R))

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

    def step_device():
        check(L.asora_raytrace_device(R, SIG, dr, 0, args.nsrc, -20.0, dlogtau, numtau, 1))
        if world > 1:
            dist.all_reduce(phi_t, op=dist.ReduceOp.SUM)

    sampler = ClockSampler(local)
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

        # ---- end to end through the reference-facing call, pinned host buffers ---------------------
        xh_host = torch.from_numpy(xh).pin_memory()
        phi_host = torch.zeros(N ** 3, dtype=torch.float64).pin_memory()
        xh_np, phi_np = xh_host.numpy(), phi_host.numpy()
        dummy = np.zeros(1)

        def step_e2e():
            if world == 1:
                libasora.do_all_sources(R, dummy, SIG, dr, dummy, xh_np, phi_np, args.nsrc, N, -20.0, dlogtau, numtau)
            else:
                check(L.asora_buffer_upload(_cabi.BUF_XH_AV, dptr(xh_np)))
                step_device()
                check(L.asora_buffer_download(_cabi.BUF_PHI_ION, dptr(phi_np)))

        for _ in range(max(1, W // 2)):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            step_e2e()
        barrier()
        e2e_s = reduce_max(time.perf_counter() - t0)
        phi_checksum = float(phi_np.sum())

    p.device_close()

    # ---- secondary: one full EoR time step (BASELINE metric "EoR step time"), rank 0, single GPU ----------
    eor = chem = None
    if rank == 0 and world == 1 and not args.no_eor:
        eor = eor_step(p, thin, thick, dlogtau)
        chem = chemistry_pass(p)

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
        line = {
            "metric": "ASORA source-cell updates/s", "value": value, "unit": "updates/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"raytracing_benchmark (test/paper_tests/raytracing_benchmark/run_test.py) at 256^3: "
                                   f"uniform ndens=1e-3, xh=2e-4, Teff=1e5 tables NumTau=20000, {args.nsrc} seeded random "
                                   f"sources per GPU, R={R:g} cells, {units_per_step // args.nsrc} visited cells per source",
                       "mesh": N, "sources_per_gpu": args.nsrc, "R_cells": R, "q_max": qmax.value,
                       "levels": levels.value, "sweep_variant": variant.value,
                       "l2": "inputs larger than L2 (3 x 134 MB grids vs 126 MB L2); no explicit flush",
                       "parallelism": f"source-sharded x{world}, one NCCL all-reduce of phi_ion (134 MB) per step"},
            "e2e": {"value": world * units_per_step * K / e2e_s, "unit": "updates/s",
                    "h2d_bytes_per_step": 8 * N ** 3, "d2h_bytes_per_step": 8 * N ** 3, "ms_per_step": 1e3 * e2e_s / K},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "sweep_smem_kernel" if variant.value == 1 else "sweep_grid_kernel",
                         "kernel_ms": k_ms, "bytes_per_update": BYTES_PER_UPDATE, "peak_source": peak_src,
                         "kernel_ms_scope": "CUDA events on the launching stream right before and after the sweep kernel's launch",
                         "sweep_ms_with_companions": float(np.mean(sweep_ms)),
                         "traffic_unit": "bytes per launch (ncu dram__bytes_read+write)",
                         "note": "bound by instruction issue and the L1 data pipe, not by HBM (L2 hit rate 92 %): see DESIGN.md and profiles/README.md"},
            "clocks": clocks,
            "phi_checksum": phi_checksum,
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
        if not args.no_cpu and world == 1:  # the CPU baseline is reported at N = 1 only
            import oracle
            threads = oracle.max_threads()
            sample = max(threads, args.cpu_sample)
            rate, secs, _ = cpu_reference_rate(R, sample, threads)
            rate1, secs1, _ = cpu_reference_rate(R, 16, 1)  # the reference itself is serial (raytracing.f90:177)
            line["cpu_baseline"] = {"value": rate, "unit": "updates/s", "cores": threads, "kind": "port",
                                    "sample": f"{sample} of the {args.nsrc} sources, {secs:.1f} s, CPU port of "
                                              "src/c2ray/raytracing.f90 with the benchmark's sub-box settings",
                                    "single_thread": {"value": rate1, "unit": "updates/s",
                                                      "sample": f"16 sources, {secs1:.1f} s"}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

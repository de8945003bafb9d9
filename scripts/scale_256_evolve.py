"""evolve3D_dist on N GPUs, 256^3, 10^4 sources, R = 30 (slabs impossible: list-order sharding), timing the three
exchanges: all-reduce + full-grid chemistry ("list") against reduce-scatter / chemistry on N^3/ranks cells / all-gather
("rsag").

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/scale_256_evolve.py
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("ASORA_QUIET", "1")
import torch
import torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import pyc2ray_b200 as p

N, NS, R = 256, 10000, 30.0
rng = np.random.default_rng(256)
ndens = 1e-3 * np.exp(rng.normal(size=(N, N, N)) * 0.5 - 0.125)
xh = np.full((N, N, N), 2e-4); temp = np.full((N, N, N), 1e4)
srcpos = p.generate_test_sources(N, NS, seed=256)
flux = 10 ** np.random.default_rng(1).normal(3.0, 0.5, size=NS)
thin, thick, dlogtau = p.blackbody_tables(1e5, False, -20.0, 4.0, 20000)
chem = (2.59e-13, -0.7, 1.3e-8 * 0.83 / 13.598 ** 2, 13.598 / 8.617e-5, 7.1e-7)
dr = 3 * 3.086e24 / N
p.device_init(N, 16)
p.photo_table_to_device(thin, thick)
res = {}
for mode in ("list", "rsag"):
    best = 1e9
    for rep in range(2):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        x, phi = p.evolve3D_dist(1e6 * 3.15576e7, dr, flux, srcpos, temp, ndens, xh, thin, thick, -20.0, dlogtau, R, 1e-4, 6.3e-18,
                                 *chem, logfile=None, quiet=True, decomposition=mode)
        torch.cuda.synchronize(); dist.barrier(); best = min(best, time.perf_counter() - t0)
    res[mode] = (best, p.evolve3D.last_niter, x)
if rank == 0:
    d = np.max(np.abs(res["list"][2] - res["rsag"][2]) / np.maximum(np.abs(res["list"][2]), 1e-300))
    for mode in ("list", "rsag"):
        print(f"256^3, {NS} sources, R={R}, {world} GPU(s), {mode}: {res[mode][0]*1e3:.1f} ms per evolve3D_dist call, "
              f"{res[mode][1]} iterations, {res[mode][0]*1e3/res[mode][1]:.2f} ms per iteration", flush=True)
    print(f"max relative difference of xh between the two exchanges: {d:.2e}")
p.device_close()
dist.destroy_process_group()

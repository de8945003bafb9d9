"""evolve3D_dist at 512^3 / 10^5 sources / R = 10.76 under torchrun: list-order sharding + all-reduce versus slab
decomposition + halo exchanges.  Prints per-iteration ray-tracing (incl. exchange) and chemistry times of rank 0."""
import io, os, re, sys, time, contextlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("ASORA_QUIET", "1")
import torch, torch.distributed as dist
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import pyc2ray_b200 as p
N, NS = int(os.environ.get("N_MESH", 512)), 100000
rng = np.random.default_rng(512)
ndens = 1.87e-4 * np.exp(rng.normal(size=(N, N, N)) * 0.5 - 0.125)
xh = np.full((N, N, N), 2e-4); temp = np.full((N, N, N), 1e4)
srcpos = p.generate_test_sources(N, NS, seed=512)
flux = 10 ** np.random.default_rng(1).normal(5.0, 0.5, size=NS)
thin, thick, dlogtau = p.blackbody_tables(1e5, False, -20.0, 4.0, 20000)
dr = 244.0 / 0.7 * 3.086e24 / N / 10.0; R = 10.76
chem = (2.59e-13, -0.7, 1.3e-8 * 0.83 / 13.598 ** 2, 13.598 / 8.617e-5, 7.1e-7)
p.device_init(N, 16); p.photo_table_to_device(thin, thick)
res = {}
for mode in ("list", "slab", "list", "slab"):
    log = f"/tmp/evolve_{mode}_{rank}.log"
    open(log, "w").close()
    dist.barrier(); t0 = time.perf_counter()
    x, phi = p.evolve3D_dist(1e7 * 3.15576e7, dr, flux, srcpos, temp, ndens, xh, thin, thick, -20.0, dlogtau, R, 1e-4, 6.3e-18,
                             *chem, logfile=log, quiet=True, decomposition=mode)
    dist.barrier(); wall = time.perf_counter() - t0
    if rank == 0:
        it = re.findall(r"Raytracing took ([0-9.]+) ms, chemistry ([0-9.]+) ms", open(log).read())
        rt = np.array([float(a) for a, b in it]); ch = np.array([float(b) for a, b in it])  # milliseconds
        print(f"{mode}: {world} GPUs, {len(it)} iterations, wall {wall:.3f} s; per iteration: ray tracing+exchange median "
              f"{1e3*np.median(rt):.2f} ms, chemistry+exchange median {1e3*np.median(ch):.2f} ms; mean x {x.mean():.6e}", flush=True)
        res[mode] = x
if rank == 0:
    print("max |x_slab - x_list| =", float(np.abs(res["slab"] - res["list"]).max()))
p.device_close(); dist.destroy_process_group()

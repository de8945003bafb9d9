"""Per-phase wall clock of the slab-decomposed 512^3 / 10^5-source evolve step (bench.py: strong.evolve_512).
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/evolve_512_phases.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ASORA_QUIET"] = "1"
import torch, torch.distributed as dist
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
with bench.quiet_stdout():
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.all_reduce(torch.zeros(1, device="cuda"))
import pyc2ray_b200 as p
thin, thick, dlogtau, numtau = bench.tables()
N, nsrc = 512, 100000
srcpos, flux, ndens, xh, temp, dr, _ = bench.eor_inputs(N, nsrc, seed=512)
R = 10.76
p.device_init(N, 64); p.photo_table_to_device(thin, thick)
for rep in range(3):
    dist.barrier(); t0 = time.perf_counter()
    p.evolve3D_dist(1e7 * 3.15576e7, dr, flux, srcpos, temp, ndens, xh, thin, thick, -20.0, dlogtau, R, 1e-4, bench.SIG, *bench.CHEM,
                    logfile=None, quiet=True, decomposition="auto", io_rank=0)
    torch.cuda.synchronize(); wall = time.perf_counter() - t0
    ph = p.evolve3D.last_phase_seconds
    print(f"rank {rank}: call {1e3*wall:.1f} ms, loop {1e3*p.evolve3D.last_loop_seconds:.1f} ms, {p.evolve3D.last_niter} iterations; "
          + ", ".join(f"{k} {1e3*v:.1f}" for k, v in ph.items()), flush=True)
p.device_close()
dist.destroy_process_group()

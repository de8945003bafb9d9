"""512^3, 10^5 sources, source-sharded over the ranks of a torchrun launch (strong scaling), one NCCL all-reduce
of phi_ion per sweep -- the fifth BASELINE.json config.  Prints one line per radius on rank 0.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/scale_512.py
"""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("ASORA_QUIET", "1")
import torch
import torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import pyc2ray_b200 as p
from pyc2ray_b200.lib import _cabi, libasora
from pyc2ray_b200.lib._cabi import L, check, dptr
from pyc2ray_b200.parallel import device_tensor, shard_bounds

N, NS = 512, 100000
rng = np.random.default_rng(512)
ndens = 1e-3 * np.exp(rng.normal(size=N ** 3) * 0.5 - 0.125)
xh = np.full(N ** 3, 2e-4)
srcpos = p.generate_test_sources(N, NS, seed=512)
flux = 10 ** np.random.default_rng(1).normal(0, 0.5, size=NS)
a, b = shard_bounds(NS, rank, world)
pos_flat, flux_flat = p.format_sources(srcpos[:, a:b], flux[a:b])
thin, thick, dlogtau = p.blackbody_tables(1e5, False, -20.0, 4.0, 20000)
p.device_init(N, 16)
stream = torch.cuda.Stream()
check(L.asora_set_stream(ctypes.c_void_p(stream.cuda_stream)))
p.photo_table_to_device(thin, thick)
libasora.source_data_to_device(pos_flat, flux_flat, b - a)
libasora.density_to_device(ndens, N)
check(L.asora_buffer_upload(_cabi.BUF_XH_AV, dptr(xh)))
phi_t = device_tensor(L.asora_device_buffer(_cabi.BUF_PHI_ION), N ** 3)
dr = 244.0 / 0.7 * 3.086e24 / N / 10.0
with torch.cuda.stream(stream):
    for R in (10.76, 30.0):
        units = NS * int(L.asora_cells_per_source(N, R))
        times = []
        for rep in range(4):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            check(L.asora_raytrace_device(R, 6.3e-18, dr, 0, b - a, -20.0, dlogtau, 20000, 1))
            if world > 1:
                dist.all_reduce(phi_t)
            e1.record(stream)
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rep > 0:
                times.append(float(t.item()))
        if rank == 0:
            ms = min(times)
            print(f"512^3, {NS} sources, R={R}: {world} GPU(s), {ms:.2f} ms per sweep+allreduce, "
                  f"{units / ms / 1e6:.1f} G updates/s, checksum {float(phi_t.sum().item()):.6e}", flush=True)
p.device_close()
if world > 1:
    dist.destroy_process_group()
